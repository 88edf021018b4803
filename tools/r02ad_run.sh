#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02ad_bench.json").read().strip().splitlines()[-1])
k=d["config"]["kernel_ms"]
print(round(d["value"],1), round(d["ms_per_step"],2), {n:v for n,v in k.items() if "conv" in n and ("fwd" in n or "grad" in n)})
PY
