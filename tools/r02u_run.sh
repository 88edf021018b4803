#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
PA2S_BENCH_DEBUG=1 timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02u_bench_$i.json 2> gpurun_out/r02u_bench_$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02u_bench_$i.json"))
print(round(d["value"],1), d["per_step"]["device_resident"]["device_ms"])
PY
grep "dbg step" gpurun_out/r02u_bench_$i.err | head -12
done
