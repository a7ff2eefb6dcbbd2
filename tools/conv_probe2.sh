#!/bin/bash
OUT=${1:-gpurun_out/conv_probe2_r02.log}
: > $OUT
export MQ_LIB=$PWD/mqgan_b200/libmqgan_b200_probes.so
for pool in 0 1; do
for d in 0 6 14 22 70 30 94 ; do
  echo "== MQ_CONV_DEBUG=$d pool=$pool" >> $OUT
  CONV_BENCH_POOL=$pool MQ_CONV_DEBUG=$d python tools/conv_bench.py 32 "pre.conv2,down0.conv1" 2>&1 | grep -E "pair  msub=(2|4)" >> $OUT
done
done
