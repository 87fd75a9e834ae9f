set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -40
