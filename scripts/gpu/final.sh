python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ALFIB_FACTOR_TIMING=1 timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.log; grep -E "timing.*npatch=4913|setup" gpurun_out/bench_final.log | tail -3; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['setup_s'], d['cpu_baseline']['value'], d['continuation']['time_s'], d['continuation']['iteration_parity'])"
