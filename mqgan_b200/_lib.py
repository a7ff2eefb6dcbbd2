"""ctypes binding of libmqgan_b200.so (the C ABI declared in include/mqgan_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call
fails, this module raises.  ``build()`` compiles it in-tree with nvcc for
sm_100a (works without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MQ_LIB") or os.path.join(_HERE, "libmqgan_b200.so")   # MQ_LIB: a bench-only probe build
CSRC = os.path.join(_HERE, "csrc")

MQ_MAX_TAPS = 16
MQ_MAX_SEGS = 6


class MqError(RuntimeError):
    pass


class ConvParams(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p),
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("in_ld", C.c_int),
        ("wpack", C.c_void_p),
        ("cout", C.c_int), ("cout_pad", C.c_int), ("bn", C.c_int),
        ("taps", C.c_int), ("nseg", C.c_int), ("kchunks", C.c_int),
        ("tap_dh", C.c_int * MQ_MAX_TAPS),
        ("tap_dw", C.c_int * MQ_MAX_TAPS),
        ("a_coff", C.c_int * MQ_MAX_SEGS),
        ("bh", C.c_int), ("bw", C.c_int), ("msub", C.c_int),
        ("bias", C.c_void_p),
        ("row_mask", C.c_void_p),
        ("mask_pre", C.c_int), ("mask_post", C.c_int),
        ("act", C.c_int), ("fast_tanh", C.c_int),
        ("beta", C.c_float), ("gamma", C.c_float),
        ("res_mode", C.c_int),
        ("res", C.c_void_p),
        ("res_is_bf16", C.c_int), ("res_ld", C.c_int), ("res_coff", C.c_int),
        ("out_f32", C.c_void_p), ("f32_ld", C.c_int), ("f32_coff", C.c_int),
        ("out_bf16", C.c_void_p), ("bf16_ld", C.c_int), ("bf16_coff", C.c_int),
        ("out_split", C.c_void_p), ("split_ld", C.c_int), ("split_seg", C.c_int),
        ("in2", C.c_void_p), ("in2_ld", C.c_int), ("up_taps", C.c_int), ("kchunks2", C.c_int),
        ("tap_dh_odd", C.c_int * MQ_MAX_TAPS),
        ("op_f16", C.c_int), ("split_kind", C.c_int),
        ("acc_scale", C.c_float),
        ("halo", C.c_int),
        ("pair", C.c_int),
        ("out_pool", C.c_void_p), ("pool_ld", C.c_int),
    ]


class Cb2dParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_is_bf16", C.c_int),
        ("B", C.c_int), ("T", C.c_int), ("C", C.c_int),
        ("dw", C.c_void_p), ("pw", C.c_void_p),
        ("bout", C.c_float),
        ("row_mask", C.c_void_p),
        ("fast_tanh", C.c_int),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("out_split", C.c_void_p),
        ("table", C.c_void_p), ("table_n", C.c_int), ("table_off", C.c_int), ("table_inv_h", C.c_float),
        ("split_kind", C.c_int),
    ]


class CbamApplyParams(C.Structure):
    _fields_ = [
        ("o", C.c_void_p), ("gate", C.c_void_p), ("res", C.c_void_p),
        ("row_mask", C.c_void_p),
        ("B", C.c_int), ("T", C.c_int), ("C", C.c_int),
        ("sam_w", C.c_void_p),
        ("beta", C.c_float), ("gamma", C.c_float),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("out_split", C.c_void_p),
        ("split_kind", C.c_int),
    ]


class VqParams(C.Structure):
    _fields_ = [
        ("z", C.c_void_p), ("n", C.c_int64), ("d", C.c_int),
        ("cb_img", C.c_void_p), ("c2", C.c_void_p), ("codebook", C.c_void_p),
        ("k", C.c_int), ("k_pad", C.c_int), ("mode", C.c_int), ("acc_scale", C.c_float),
        ("idx", C.c_void_p), ("codes_out", C.c_void_p), ("dist_out", C.c_void_p),
        ("fold", C.c_int), ("zconst", C.c_float),
    ]


class MelspecParams(C.Structure):
    _fields_ = [
        ("wav", C.c_void_p), ("wav_ld", C.c_int64), ("lengths", C.c_void_p), ("B", C.c_int),
        ("n_fft", C.c_int), ("hop", C.c_int), ("n_mels", C.c_int), ("n_freqs", C.c_int),
        ("window", C.c_void_p), ("twiddle", C.c_void_p),
        ("fb_start", C.c_void_p), ("fb_count", C.c_void_p), ("fb_off", C.c_void_p), ("fb_w", C.c_void_p),
        ("clip", C.c_float), ("out", C.c_void_p), ("out_frames", C.c_int64),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("dy", C.c_void_p), ("dy_ld", C.c_int),
        ("x", C.c_void_p), ("x_ld", C.c_int),
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("cout", C.c_int), ("cin", C.c_int),
        ("taps", C.c_int),
        ("tap_dh", C.c_int * MQ_MAX_TAPS),
        ("tap_dw", C.c_int * MQ_MAX_TAPS),
        ("bh", C.c_int), ("bw", C.c_int),
        ("split", C.c_int),
        ("dw", C.c_void_p),
    ]


class FsqParams(C.Structure):
    _fields_ = [
        ("D", C.c_int),
        ("half_l", C.c_float * 8), ("shift", C.c_float * 8), ("offset", C.c_float * 8),
        ("half_w", C.c_int * 8), ("basis", C.c_int * 8), ("levels", C.c_int * 8),
    ]


# name -> (restype, argtypes); must list every symbol include/mqgan_b200.h declares
SIGNATURES = {
    "mq_version": (C.c_int, []),
    "mq_last_error": (C.c_char_p, []),
    "mq_device_check": (C.c_int, []),
    "mq_sm_count": (C.c_int, []),
    "mq_tmem_read_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "mq_conv_gemm": (C.c_int, [C.POINTER(ConvParams), C.c_void_p]),
    "mq_split_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "mq_convblock2d": (C.c_int, [C.POINTER(Cb2dParams), C.c_void_p]),
    "mq_cam_chunks": (C.c_int, [C.c_int]),
    "mq_cam_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mq_cam_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_cbam_apply": (C.c_int, [C.POINTER(CbamApplyParams), C.c_void_p]),
    "mq_qin_fsq": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                             C.POINTER(FsqParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_fsq_quantize": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(FsqParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_vq_nearest": (C.c_int, [C.POINTER(VqParams), C.c_void_p]),
    "mq_conv_wgrad_split": (C.c_int, [C.POINTER(WgradParams)]),
    "mq_conv_wgrad": (C.c_int, [C.POINTER(WgradParams), C.c_void_p]),
    "mq_cb2d_point_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mq_cb2d_grad_blocks": (C.c_int, [C.c_int64, C.c_int]),
    "mq_cb2d_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_act_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    "mq_act_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_act_bias_blocks": (C.c_int, [C.c_int64, C.c_int]),
    "mq_leaky_mask_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                                        C.c_void_p, C.c_void_p]),
    "mq_leaky_mask_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                         C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_npy_probe": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "mq_npy_read_f32": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    "mq_npy_write_f32": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int64]),
    "mq_log_mel": (C.c_int, [C.POINTER(MelspecParams), C.c_void_p]),
    "mq_code_gather": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_refiner_masks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_zero_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "mq_avgpool_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mq_upcat_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]),
    "mq_avgpool_mask_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mq_upcat_mask_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]),
    "mq_refiner_stem_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_refiner_stem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mq_refiner_tail": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mq_sequence_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
launch_count = 0     # kernels launched through this binding (bench.py reports it)


def build(verbose: bool = False) -> str:
    """Compile libmqgan_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise MqError("building libmqgan_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MqError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU/PyTorch fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C mqgan_b200/csrc`.")
        l = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = l
        return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mq_last_error().decode(errors="replace")
        raise MqError(f"{what} failed (rc={rc}): {msg}")


class LaunchProfiler:
    """Times every launch with CUDA events on the launching (current) stream.
    ``bench.py`` uses it for the per-kernel roofline table; off by default."""

    def __init__(self):
        self.records = []          # (entry point, meta dict, start event, end event)

    def begin(self, name, meta):
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._cur = (name, meta or {}, ev)

    def end(self):
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        name, meta, ev0 = self._cur
        self.records.append((name, meta, ev0, ev))

    def summary(self):
        """[(name, tag, launches, total ms, total algorithmic flops)] after a device sync."""
        import torch
        torch.cuda.synchronize()
        agg = {}
        for name, meta, e0, e1 in self.records:
            key = (name, meta.get("tag", ""))
            a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += float(meta.get("flops", 0.0))
            a[3] += float(meta.get("mma_flops", 0.0))
        return [(k[0], k[1], v[0], v[1], v[2], v[3]) for k, v in agg.items()]


profiler = None


NVTX = os.environ.get("MQ_NVTX", "0") == "1"     # NVTX range per library launch, named "<entry point> <layer tag>" (nsys / ncu --nvtx)


def call(name: str, *args, meta=None) -> None:
    """Invoke a launching entry point and raise on a non-zero status."""
    global launch_count
    if profiler is not None:
        profiler.begin(name, meta)
    if NVTX:
        import torch
        torch.cuda.nvtx.range_push(name + (" " + meta["tag"] if meta and meta.get("tag") else ""))
    try:
        rc = getattr(lib(), name)(*args)
    finally:
        if NVTX:
            torch.cuda.nvtx.range_pop()
    if rc != 0:
        check(rc, name)
    launch_count += 1
    if profiler is not None:
        profiler.end()
