N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r10_bench_n$N.json 2> gpurun_out/r10_bench_n$N.err
tail -c 300 gpurun_out/r10_bench_n$N.err
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -2
