#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02x_tests.log
tail -4 gpurun_out/r02x_tests.log
timeout 300 python bench.py --workload infer --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_infer.json 2> gpurun_out/r02x_infer.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02x_infer.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["kernel_ms"])
PY
