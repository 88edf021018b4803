#!/bin/bash
# encoder recurrences with 2 and 4 clips per cluster (PA2S_GRU_BG): measured 2.53 vs 1.69 ms per layer
mkdir -p gpurun_out
for bg in 2 4; do
PA2S_GRU_BG=$bg timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/gru_bg_bench.json 2> gpurun_out/gru_bg_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/gru_bg_bench.json").read().strip().splitlines()[-1])
k=d["config"]["kernel_ms"]
print("bg=$bg", round(d["value"],1), round(d["ms_per_step"],2), {n:v for n,v in k.items() if "gru" in n})
PY
done
