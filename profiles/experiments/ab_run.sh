timeout 900 python -m pytest tests/test_gpu_tree_vanilla.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python profiles/experiments/lake_time.py
