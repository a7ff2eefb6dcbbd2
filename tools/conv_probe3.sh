#!/bin/bash
# A/B of the epilogue modes on the fused up-concat layers (usage: bash tools/conv_probe3.sh OUT.log [envs...])
out=${1:-gpurun_out/conv_probe3.log}
: > $out
for envs in "" "MQ_STAGE_OUT=0" "MQ_STAGE_OUT=1"; do
  echo "== env: ${envs:-default}" >> $out
  env $envs python tools/conv_bench.py 32 "up0.conv1,up1.conv1,up2.conv1" 2>&1 | grep pair >> $out
done
