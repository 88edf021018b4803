#!/bin/bash
# development loop on the GPU box: tensor-core kernel tests, the whole GPU suite, one short bench line with the conv / decoder timers
# usage: gpurun -- bash tools/dev_check.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/dev_tc.log 2>&1; echo "tc tests exit $?" >> gpurun_out/dev_tc.log; tail -3 gpurun_out/dev_tc.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/dev_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/dev_tests.log; tail -3 gpurun_out/dev_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/dev_bench.json").read().strip().splitlines()[-1])
k=d["config"]["kernel_ms"]
print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), {n:v for n,v in k.items() if "wgrad" in n or "dgrad" in n or "_fwd" in n})
PY
