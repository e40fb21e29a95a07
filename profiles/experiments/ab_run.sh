timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warning | tail -8
