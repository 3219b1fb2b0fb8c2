timeout 1500 python -m pytest tests/test_gpu_scale.py -x -q --durations=10 2>&1 | tail -25
