#!/bin/bash
mkdir -p gpurun_out
python tools/prof_decm.py > gpurun_out/r02r_prof_decm.txt 2>&1; grep -E "^(fwd|bwd)" gpurun_out/r02r_prof_decm.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02r_tests.log
tail -3 gpurun_out/r02r_tests.log
for i in 1 2 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02r_bench_$i.json 2> gpurun_out/r02r_bench_$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02r_bench_$i.json"))
print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "host", d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_idle_device"), "mallocs", d.get("cuda_mallocs_in_timed_region"))
print(d["per_step"]["device_resident"]["device_ms"])
PY
done
