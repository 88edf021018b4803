#!/bin/bash
# final 1-GPU records of round 2: smoke, GPU tests, bench lines of every configuration + the CPU arms
TAG=${1:-r02z}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],2), "cpu", (d.get("cpu_baseline") or {}).get("value"), "also", (d.get("also") or {}).get("value"))
except Exception as e: print("$name FAILED", e)
PY
}
run n1 --steps 10 --warmup 3
run bf16_n1 --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline
run finetune_n1 --steps 10 --warmup 3 --workload finetune --no-cpu-baseline
run infer_n1 --steps 4 --warmup 3 --workload infer
run reference_cpu --impl reference --steps 2 --warmup 1
run reference_cpu_infer --impl reference --workload infer --steps 1 --warmup 1
run reference_cpu_16clips --impl reference --cpu-clips 16 --steps 1 --warmup 0
