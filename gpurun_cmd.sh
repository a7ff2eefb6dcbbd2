D=$(mktemp -d)
CFG=$(python tools/train_cli_smoke.py $D | tail -1)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 train.py --config $CFG > gpurun_out/train_cli_2gpu.log 2>&1
echo "rc=$?"; grep -E "Epoch \[|Checkpoint|Training finished|Error|error" gpurun_out/train_cli_2gpu.log | head -12
ls $D/run | head; wc -l $D/run/train_log.jsonl; tail -1 $D/run/train_log.jsonl | cut -c1-300
