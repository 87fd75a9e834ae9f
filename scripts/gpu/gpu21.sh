# condensed inverses: whole GPU suite, smoke, full-size bench
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_r1_condensed.json 2> gpurun_out/bench_r1_condensed.log; grep -E "setup|built" gpurun_out/bench_r1_condensed.log | tail -4; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_condensed.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline'], d['setup_s'], d['cpu_baseline']['value'], d['continuation']['time_s'], d['continuation']['iteration_parity']); print(d['breakdown_ms']); print(d['config'])"
