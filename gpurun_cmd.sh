timeout 900 python -m pytest tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/train_parity.py gpurun_out/train_parity_r01b.json > gpurun_out/train_parity.log 2>&1; tail -2 gpurun_out/train_parity.log
timeout 900 python tools/train_bench.py --steps 5 --warmup 3 --layers gpurun_out/train_layers_r01b.md > gpurun_out/train_bench_r01b.log 2>&1; tail -1 gpurun_out/train_bench_r01b.log | cut -c1-400
