N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r8_bench_n$N.json 2> gpurun_out/r8_bench_n$N.err
tail -c 300 gpurun_out/r8_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r8_benchref_n$N.json 2> gpurun_out/r8_benchref_n$N.err
tail -c 200 gpurun_out/r8_benchref_n$N.json
