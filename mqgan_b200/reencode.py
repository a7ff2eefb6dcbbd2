"""Host side of the re-encode CLIs (reference: reencode_spectrograms.py:8-88 and
reencode_spectrograms_from_checkpoint.py:9-108).

Same observable behaviour: walk ``input_dir`` for ``*.npy`` in ``os.walk`` order,
form consecutive ``batch_size`` chunks, zero-pad each chunk to its longest
utterance, ``encode`` -> ``decode`` with the lengths, trim, and save float32
``(T, n_mels)`` arrays under the mirrored relative path; a failing batch is
reported and skipped.

Added (additive, never changes results): the unit of sharding is the reference's
batch (SURVEY §8e - batch composition influences the encoder through padding), so
with ``world`` workers batch ``i`` goes to worker ``i % world`` and every worker
writes a disjoint set of files; there is no data-path collective.  Reading and
writing happen on background threads so disk I/O overlaps the GPU.
"""
from __future__ import annotations

import contextlib
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def list_npy_files(input_dir: str) -> List[str]:
    """Recursive *.npy listing in the reference's order (reencode_spectrograms.py:30-34)."""
    out = []
    for root, _, files in os.walk(input_dir):
        for file in files:
            if file.endswith(".npy"):
                out.append(os.path.join(root, file))
    return out


def make_batches(files: Sequence[str], batch_size: int) -> List[List[str]]:
    """Consecutive chunks (reencode_spectrograms.py:43)."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    return [list(files[i:i + batch_size]) for i in range(0, len(files), batch_size)]


def shard_indices(n_batches: int, rank: int, world: int) -> List[int]:
    """Batch indices owned by ``rank``: round-robin, disjoint, covering."""
    if not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    return list(range(rank, n_batches, world))


IO_THREADS = int(os.environ.get("MQ_IO_THREADS", "4"))    # file reads / writes inside one batch run concurrently
_io_pool: Optional[ThreadPoolExecutor] = None


def _pool() -> ThreadPoolExecutor:
    global _io_pool
    if _io_pool is None:
        _io_pool = ThreadPoolExecutor(max_workers=max(1, IO_THREADS), thread_name_prefix="mq_io")
    return _io_pool


def sort_batches_by_length(files: Sequence[str]) -> List[str]:
    """Optional (``--sort_by_length``): order files by frame count so that batches are nearly rectangular
    (less padding).  NOT the reference's batch composition: batch composition influences the encoder
    through padding, so this can change indices; off by default."""
    def n_frames(p):
        if NATIVE_IO:
            sh = _native_probe(p)              # header only, GIL-free
            if sh is not None:
                return sh[0]
        return int(np.load(p, mmap_mode="r").shape[0])
    lens = list(_pool().map(n_frames, files))
    return [f for _, f in sorted(zip(lens, files), key=lambda t: (t[0], t[1]))]


NATIVE_IO = os.environ.get("MQ_NATIVE_IO", "1") != "0"     # .npy parsing / writing in the library (GIL-free), numpy otherwise


class _PinnedPool:
    """Reusable page-locked staging buffers: cudaHostAlloc per batch costs more than the copy it speeds up."""

    def __init__(self):
        self._free: List[torch.Tensor] = []
        self._lock = threading.Lock()

    def get(self, numel: int) -> torch.Tensor:
        with self._lock:
            for i, t in enumerate(self._free):
                if t.numel() >= numel:
                    return self._free.pop(i)
        n = max(int(numel * 1.25), 1 << 20)
        return torch.empty(n, dtype=torch.float32, pin_memory=torch.cuda.is_available())

    def put(self, t: Optional[torch.Tensor]) -> None:
        if t is not None:
            with self._lock:
                if len(self._free) < 8:
                    self._free.append(t)


_pinned = _PinnedPool()


def _native_probe(path: str):
    """(rows, cols) of a float .npy the library can read, None if it needs numpy."""
    from . import _lib
    import ctypes as C
    rows, cols = C.c_int64(), C.c_int64()
    rc = _lib.lib().mq_npy_probe(os.fsencode(path), C.byref(rows), C.byref(cols), None, None)
    if rc == 4:
        return None
    if rc != 0:
        raise OSError(_lib.lib().mq_last_error().decode(errors="replace"))
    return int(rows.value), int(cols.value)


def load_and_pad(paths: Sequence[str], pooled: bool = False):
    """Load each file, zero-pad to the longest, stack, float32 (reencode_spectrograms.py:49-62).  Returns
    (batch (B, Tmax, M), lengths) - plus the pooled staging buffer to hand back to ``_pinned.put`` when ``pooled``."""
    def done(batch, lengths, handle=None):
        return (batch, lengths, handle) if pooled else (batch, lengths)

    threaded = len(paths) > 1 and IO_THREADS > 1
    if NATIVE_IO:
        from . import _lib
        import ctypes as C
        shapes = list(_pool().map(_native_probe, paths)) if threaded else [_native_probe(p) for p in paths]
        if all(sh is not None for sh in shapes):
            lengths = [sh[0] for sh in shapes]
            max_len, n_mels = max(lengths), shapes[0][1]
            for p, sh in zip(paths, shapes):
                if sh[1] != n_mels:
                    raise ValueError(f"{p}: {sh[1]} mel channels, expected {n_mels}")
            flat = _pinned.get(len(paths) * max_len * n_mels) if pooled else torch.empty(len(paths) * max_len * n_mels)
            batch = flat[: len(paths) * max_len * n_mels].view(len(paths), max_len, n_mels)
            base, pitch = batch.data_ptr(), max_len * n_mels * 4
            lib = _lib.lib()

            def read_one(i):
                rc = lib.mq_npy_read_f32(os.fsencode(paths[i]), base + i * pitch, max_len, n_mels, None)
                if rc != 0:
                    raise OSError(lib.mq_last_error().decode(errors="replace"))

            if threaded:
                list(_pool().map(read_one, range(len(paths))))
            else:
                for i in range(len(paths)):
                    read_one(i)
            return done(batch, lengths, flat if pooled else None)
    specs = list(_pool().map(np.load, paths)) if threaded else [np.load(p) for p in paths]
    lengths = [int(s.shape[0]) for s in specs]
    max_len = max(lengths)
    n_mels = specs[0].shape[1]
    batch = np.zeros((len(specs), max_len, n_mels), dtype=np.float32)
    for i, s in enumerate(specs):
        if s.shape[1] != n_mels:
            raise ValueError(f"{paths[i]}: {s.shape[1]} mel channels, expected {n_mels}")
        batch[i, : s.shape[0]] = s
    return done(torch.from_numpy(batch), lengths)


def save_outputs(reencoded: torch.Tensor, lengths: Sequence[int], paths: Sequence[str], input_dir: str,
                 output_dir: str) -> None:
    """Trim to the original length and save under the mirrored path (:69-81)."""
    t = reencoded if not reencoded.is_cuda else reencoded.cpu()
    native = NATIVE_IO and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 3
    arr = None if native else t.numpy()
    if native:
        from . import _lib
        lib = _lib.lib()
        base, pitch, n_mels = t.data_ptr(), t.shape[1] * t.shape[2] * 4, t.shape[2]

    def save_one(i):
        out_path = os.path.join(output_dir, os.path.relpath(paths[i], input_dir))
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        if native:
            rc = lib.mq_npy_write_f32(os.fsencode(out_path), base + i * pitch, int(lengths[i]), n_mels)
            if rc != 0:
                raise OSError(lib.mq_last_error().decode(errors="replace"))
        else:
            np.save(out_path, np.ascontiguousarray(arr[i, : lengths[i], :], dtype=np.float32))

    if len(paths) > 1 and IO_THREADS > 1:
        list(_pool().map(save_one, range(len(paths))))
    else:
        for i in range(len(paths)):
            save_one(i)


def _cap_host_threads(world: int) -> None:
    """The host side of this path is file I/O and launches; torch's intra-op pool (as wide as the machine by default)
    only adds spinning threads.  Measured with 4 workers on a 32-core box (tools/cli_probe4.sh): 0.73 M frames/s with
    the default pools, 2.60 M with them capped at 4 threads, 3.05 M with one I/O thread per worker as well.
    MQ_WORKER_THREADS=0 / MQ_IO_THREADS=<n> override."""
    global IO_THREADS, _io_pool
    nthr = int(os.environ.get("MQ_WORKER_THREADS", str(max(1, min(4, (os.cpu_count() or 1) // max(1, world))))))
    if nthr > 0:
        torch.set_num_threads(nthr)
    if "MQ_IO_THREADS" not in os.environ and world >= 4 and _io_pool is None:
        IO_THREADS = 1


def reencode_tree(run_batch: Callable[[torch.Tensor, List[int]], torch.Tensor], input_dir: str, output_dir: str,
                  batch_size: int, rank: int = 0, world: int = 1, progress: bool = True,
                  prefetch: int = 2, sort_by_length: bool = False,
                  files: Optional[Sequence[str]] = None) -> Tuple[int, int]:
    """Process this worker's share of the tree.  ``run_batch(batch (B,T,M) float32 CPU, lengths)``
    returns the re-encoded (B,T,M) tensor (any device).  Returns (files done, batches failed)."""
    _cap_host_threads(world)
    listed = files is not None              # run_multi_gpu lists (and sorts) once in the parent and hands the order down
    if not listed:
        files = list_npy_files(input_dir)
    if not files:
        print("Warning: No .npy files were found.")
        return 0, 0
    if rank == 0:
        print(f"Found {len(files)} spectrogram files to process.")
    if sort_by_length and not listed:
        files = sort_batches_by_length(files)
    batches = make_batches(files, batch_size)
    mine = shard_indices(len(batches), rank, world)

    in_q: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))
    out_q: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))

    def reader():
        for bi in mine:
            paths = batches[bi]
            try:
                batch, lengths, handle = load_and_pad(paths, pooled=True)
                if handle is None and torch.cuda.is_available():
                    batch = batch.pin_memory()
                in_q.put((paths, batch, lengths, None, handle))
            except Exception as e:  # noqa: BLE001 - mirror the reference's catch-all (:83-85)
                in_q.put((paths, None, None, e, None))
        in_q.put(None)

    failed = [0]

    def writer():
        while True:
            item = out_q.get()
            if item is None:
                return
            paths, out, lengths, handle = item
            try:
                save_outputs(out, lengths, paths, input_dir, output_dir)
            except Exception as e:  # noqa: BLE001
                failed[0] += 1
                print(f"\nCould not save batch starting with {paths[0]}. Error: {e}")
            finally:
                _pinned.put(handle)

    rt = threading.Thread(target=reader, daemon=True)
    wt = threading.Thread(target=writer, daemon=True)
    rt.start()
    wt.start()
    pbar = None
    if progress and rank == 0:
        try:
            from tqdm import tqdm
            pbar = tqdm(total=len(mine), desc="Re-encoding Spectrograms")
        except Exception:
            pbar = None
    cuda = torch.cuda.is_available()
    dev_index = torch.cuda.current_device() if cuda else None
    # MQ_COMPUTE_THREADS=2 runs two compute threads, each on its own CUDA stream, so that one enqueues the next batch
    # while the other waits for its result (streams and the current device are per thread in torch; the engine keeps no
    # per-call state).  Measured on one B200 (tools/cli_bench_multi.py, 4096 files): 874 k frames/s against 858 k with
    # the serial loop, 631 k with three threads - not worth a default, the limiter is elsewhere.
    n_compute = max(1, int(os.environ.get("MQ_COMPUTE_THREADS", "1"))) if cuda else 1
    done = [0]
    lock = threading.Lock()

    def compute():
        stream = None
        if cuda:
            torch.cuda.set_device(dev_index)
            stream = torch.cuda.Stream() if n_compute > 1 else None
        while True:
            item = in_q.get()
            if item is None:
                in_q.put(None)                                  # let the sibling thread see the end marker too
                return
            paths, batch, lengths, err, in_handle = item
            try:
                if err is not None:
                    raise err
                ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
                with ctx:
                    out = run_batch(batch, lengths)
                    out_handle = None
                    if out.is_cuda:
                        if out.dtype == torch.float32:
                            out_handle = _pinned.get(out.numel())
                            host = out_handle[: out.numel()].view(out.shape)
                        else:
                            host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
                        host.copy_(out, non_blocking=True)
                        torch.cuda.current_stream().synchronize()  # also: the input staging buffer has been consumed
                        out = host
                out_q.put((paths, out, lengths, out_handle))
                with lock:
                    done[0] += len(paths)
            except Exception as e:  # noqa: BLE001
                with lock:
                    failed[0] += 1
                print(f"\nCould not process batch starting with {paths[0]}. Error: {e}")
            finally:
                _pinned.put(in_handle)
                if pbar is not None:
                    with lock:
                        pbar.update(1)

    if n_compute == 1:
        compute()
    else:
        cts = [threading.Thread(target=compute, daemon=True) for _ in range(n_compute)]
        for t in cts:
            t.start()
        for t in cts:
            t.join()
    if pbar is not None:
        pbar.close()
    out_q.put(None)
    wt.join()
    return done[0], failed[0]


def dist_env() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment, (0, 1, 0) otherwise."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def finish_distributed(done: int, failed: int, backend: Optional[str] = None) -> Tuple[int, int]:
    """Sum the per-worker counters (the only exchange on this path) when launched under torchrun."""
    rank, world, _ = dist_env()
    if world == 1:
        return done, failed
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
        created = True
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([done, failed], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    if created:
        dist.destroy_process_group()
    return int(t[0]), int(t[1])


def _worker(rank: int, world: int, make_model: Callable[[str], object], input_dir: str, output_dir: str,
            batch_size: int, ret, sort_by_length: bool = False, files: Optional[Sequence[str]] = None) -> None:
    import time
    torch.cuda.set_device(rank)
    t_load = time.time()
    model = make_model(f"cuda:{rank}")

    def run(batch, lengths):
        idx = model.encode(batch, lengths=lengths)
        return model.decode(idx, lengths=lengths)

    t0 = time.time()
    done, failed = reencode_tree(run, input_dir, output_dir, batch_size, rank, world, progress=(rank == 0),
                                 sort_by_length=sort_by_length, files=files)
    torch.cuda.synchronize()
    # (done, failed, model load seconds, processing start / end as epoch seconds): tools/cli_bench_multi.py reads the span
    ret[rank] = (done, failed, t0 - t_load, t0, time.time())


def run_multi_gpu(make_model: Callable[[str], object], input_dir: str, output_dir: str, batch_size: int,
                  gpus: int, sort_by_length: bool = False) -> Tuple[int, int]:
    """One worker process per GPU (``--gpus N``); each builds its own model copy."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    # the tree is listed - and, for --sort_by_length, its headers probed - once here, not once per worker (with eight
    # workers the per-worker probe of all files took longer than the re-encoding itself)
    files = list_npy_files(input_dir)
    if sort_by_length and files:
        files = sort_batches_by_length(files)
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, gpus, make_model, input_dir, output_dir, batch_size, ret,
                                                   sort_by_length, files))
                 for r in range(gpus)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        done = sum(v[0] for v in ret.values())
        failed = sum(v[1] for v in ret.values()) + sum(1 for p in procs if p.exitcode != 0)
        if os.environ.get("MQ_CLI_TIMING") == "1" and len(ret) > 0:
            import json
            vals = list(ret.values())
            print("MQ_CLI_TIMING " + json.dumps({
                "workers": gpus, "load_s_max": max(v[2] for v in vals),
                "process_span_s": max(v[4] for v in vals) - min(v[3] for v in vals),
                "process_s_per_worker": [v[4] - v[3] for v in vals]}), flush=True)
    return done, failed
