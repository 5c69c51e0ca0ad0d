python -m pytest tests -m gpu -x -q > gpurun_out/r10_gpu_all.log 2>&1; tail -2 gpurun_out/r10_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r10_bench_n1.json 2> gpurun_out/r10_bench_n1.err; tail -2 gpurun_out/r10_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r10_benchref_n1.json 2> gpurun_out/r10_benchref_n1.err; tail -c 300 gpurun_out/r10_benchref_n1.json
