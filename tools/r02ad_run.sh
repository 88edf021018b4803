#!/bin/bash
mkdir -p gpurun_out
for v in a b; do
cp piano_a2s_b200/libpa2s_$v.so piano_a2s_b200/libpa2s.so
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02ad_bench.json").read().strip().splitlines()[-1])
k=d["config"]["kernel_ms"]
print("$v", round(d["value"],1), round(d["ms_per_step"],2), {n:v for n,v in k.items() if "conv" in n and ("fwd" in n or "dgrad" in n)})
PY
done
