timeout 900 python -m pytest tests/test_gpu_tree_cta.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python profiles/experiments/c3_time.py 2>&1 | tail -8
