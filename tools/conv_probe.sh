#!/bin/bash
# Bottleneck probes of the conv kernels on the narrow refiner layers (needs `make -C mqgan_b200/csrc probes`):
# MQ_CONV_DEBUG bits: 1 no epilogue math/stores, 2 no MMA, 4 no TMA, 8 no global stores.
OUT=${1:-gpurun_out/conv_probe_r02.log}
: > $OUT
export MQ_LIB=$PWD/mqgan_b200/libmqgan_b200_probes.so
for pool in 0 1; do
for d in 0 1 2 3 4 6 8; do
  echo "== MQ_CONV_DEBUG=$d pool=$pool" >> $OUT
  CONV_BENCH_POOL=$pool MQ_CONV_DEBUG=$d python tools/conv_bench.py 32 "pre.conv2,down0,up2" 2>&1 | grep -E "pair|halo" >> $OUT
done
done
