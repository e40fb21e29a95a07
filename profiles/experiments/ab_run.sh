timeout 900 python -m pytest tests/test_gpu_tree_cta.py -x -q -m gpu -k tree_parallel -s 2>&1 | tail -12
