#!/bin/bash
for i in 1 2 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_bench_$i.json 2> gpurun_out/r02w_bench_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02w_bench_$i.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "also", round(d["also"]["value"],1), round(d["also"]["e2e_value"],1))
print("  dev", d["per_step"]["device_resident"]["device_ms"]); print("  e2e", d["per_step"]["e2e"]["device_ms"])
PY
done
nproc; uptime
