"""One re-encode step of a bench workload between cudaProfilerStart / cudaProfilerStop, for ncu:

    python tools/ncu_step.py [--workload W] [--batch B]                                 # plain run first (must exit 0)
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/ncu_step.py
    ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_pair_kernel -c 15 \
        -o gpurun_out/prof_pair python tools/ncu_step.py --batch 32

Weights are packed and one warm-up step runs before the profiled region, so every captured launch is a steady-state
launch of the step bench.py times.  Prints the launch tags in order (the n-th conv_pair_kernel launch is the n-th
"pair" line), so a capture can be mapped to its layer.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from mqgan_b200 import _lib, spec as S
from mqgan_b200.synth import synth_mels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD, choices=sorted(bench.WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="utterances (default: the workload's)")
    args = ap.parse_args()
    cfg_name, B, T, prec, dprec = bench.WORKLOADS[args.workload]
    B = args.batch or B
    cfg = getattr(S, cfg_name)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model, _ = bench.build_model(cfg, prec, dev, decoder_precision=dprec)
    eng = model.engine()
    mel = synth_mels(B, T, cfg.mel_channels, seed=1).to(dev)
    idx = eng.encode(mel, None)
    eng.decode(idx, None)
    torch.cuda.synchronize()

    class Tags:
        records = []

        def begin(self, name, meta):
            self.records.append((name, (meta or {}).get("tag", "")))

        def end(self):
            pass

    _lib.profiler = Tags()
    torch.cuda.profiler.start()
    idx = eng.encode(mel, None)
    out = eng.decode(idx, None)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    recs = _lib.profiler.records
    _lib.profiler = None
    print(f"workload {args.workload} batch {B}: {len(recs)} library launches in the profiled step; checksum {float(out.double().sum()):.6e}")
    for i, (name, tag) in enumerate(recs):
        if name == "mq_conv_gemm":
            print(f"conv launch {i}: {tag}")


if __name__ == "__main__":
    main()
