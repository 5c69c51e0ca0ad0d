python -m pytest tests -m gpu -x -q > gpurun_out/r9_gpu_all.log 2>&1; tail -3 gpurun_out/r9_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py) > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; tail -4 gpurun_out/r9_bench.err
