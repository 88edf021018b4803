#!/bin/bash
# the last record of a round: ncu launch list of one timed step + the default headline bench line
TAG=${1:-last}
mkdir -p gpurun_out
export PA2S_PROFILE_RANGE=1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
unset PA2S_PROFILE_RANGE
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 600 gpurun_out/${TAG}_bench_n1.json | head -c 300; echo
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["value"], "also", (d.get("also") or {}).get("value"))
PY
