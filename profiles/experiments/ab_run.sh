ISLA_LIB_PATH=profiles/experiments/lib_pt.so timeout 600 python profiles/experiments/r2_sweep.py 8192:0 2>&1 | grep -v "^trees" | sort | uniq -c | sort -rn | head -0
ISLA_LIB_PATH=profiles/experiments/lib_pt.so timeout 600 python profiles/experiments/r2_sweep.py 8192:0 2>&1 | tail -8
ISLA_LIB_PATH=profiles/experiments/lib_pt.so timeout 600 python profiles/experiments/r2_sweep.py 65536:0 2>&1 | tail -5
