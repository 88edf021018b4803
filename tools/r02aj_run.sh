#!/bin/bash
mkdir -p gpurun_out
for bg in 2 4; do
PA2S_GRU_BG=$bg timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02aj_bench.json 2> gpurun_out/r02aj_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02aj_bench.json").read().strip().splitlines()[-1])
k=d["config"]["kernel_ms"]
print("bg=$bg", round(d["value"],1), round(d["ms_per_step"],2), {n:v for n,v in k.items() if "gru" in n})
PY
done
