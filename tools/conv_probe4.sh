#!/bin/bash
# A/B of one vs two MMA-issuing warps on the pair kernel (usage: bash tools/conv_probe4.sh OUT.log)
out=${1:-gpurun_out/conv_probe4.log}
: > $out
for envs in "MQ_MMA_ISSUERS=1" "MQ_MMA_ISSUERS=2" "MQ_MMA_ISSUERS=1" "MQ_MMA_ISSUERS=2"; do
  echo "== env: ${envs}" >> $out
  env $envs python tools/conv_bench.py 32 "pre.conv2,down0,down1.conv1,up1.conv1,up2.conv1" 2>&1 | grep "pair" | grep -v "msub=1 " >> $out
done
