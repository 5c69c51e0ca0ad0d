python -m pytest tests -m gpu -x -q > gpurun_out/r11_gpu_all.log 2>&1; tail -2 gpurun_out/r11_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r11_bench_n1.json 2> gpurun_out/r11_bench_n1.err; tail -2 gpurun_out/r11_bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r11_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:(round(v['value'],1), round(v['ms_per_step'],4)) for k,v in d['suite'].items()})"
