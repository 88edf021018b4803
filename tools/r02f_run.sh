#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
tail -3 gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02f_bench_n1.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "host", d["host_enqueue_ms_per_step"], "launches", d["gpu_launches"])
for r in d["roofline_all"][:6]: print(r["timer"], r["launches_timed"], round(r["ms_per_launch"],3), round(r["ms_total"],2), round(r["frac"],3))
print(d["per_step"]["device_resident"])
PY
