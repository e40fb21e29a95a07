timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_episodes.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python profiles/experiments/roll_time.py
