python -m pytest tests -m gpu -x -q > gpurun_out/r7_gpu_all.log 2>&1; tail -2 gpurun_out/r7_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py) > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err; tail -4 gpurun_out/r7_bench.err
(time python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/r7_benchref.json 2> gpurun_out/r7_benchref.err; tail -4 gpurun_out/r7_benchref.err
