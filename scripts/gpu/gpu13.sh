python -m pytest tests/test_gpu_edges.py -q 2>&1 | tail -12
timeout 1500 python bench.py > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.log; tail -3 gpurun_out/bench_r1_d.log; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_d.json')); print(d['ms_per_step'], d['value'], d['gpu_launches']); print(json.dumps(d['continuation'])[:900]); print(json.dumps(d['cpu_baseline'])[:1200])"
