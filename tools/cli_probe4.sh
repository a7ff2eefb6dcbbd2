export MQ_BENCH_TREE=/dev/shm/mq_tree
python tools/cli_bench_multi.py 8192 4 > gpurun_out/cli4_default.log 2>&1
MQ_WORKER_THREADS=0 python tools/cli_bench_multi.py 8192 4 > gpurun_out/cli4_nothreadcap.log 2>&1
MQ_IO_THREADS=1 python tools/cli_bench_multi.py 8192 4 > gpurun_out/cli4_io1.log 2>&1
CUDA_DEVICE_SCHEDULE=blocking python tools/cli_bench_multi.py 8192 4 > gpurun_out/cli4_blocking.log 2>&1
MQ_NATIVE_IO=0 python tools/cli_bench_multi.py 8192 4 > gpurun_out/cli4_numpyio.log 2>&1
python tools/cli_bench_multi.py 8192 1 > gpurun_out/cli4_one.log 2>&1
for f in default nothreadcap io1 blocking numpyio one; do echo "== $f"; cut -c1-520 gpurun_out/cli4_$f.log; done
