timeout 600 python -m pytest tests/test_gpu_tree_vanilla.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python profiles/experiments/r2_sweep.py 1:0 256:0 4096:0 8192:0 16384:0 32768:0 65536:0
