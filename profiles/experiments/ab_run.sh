timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -3 gpurun_out/bench_r2b.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_r2b.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["e2e"]["value"])
x=d["extras"]
for k in x:
    if k.startswith("c") or k.startswith("agent"): print(k, json.dumps(x[k])[:400])
print(d.get("cpu_baseline"))
P
