#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encoder" > gpurun_out/r02aa_enc.log 2>&1; echo "enc tests exit $?" >> gpurun_out/r02aa_enc.log; tail -3 gpurun_out/r02aa_enc.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02aa_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02aa_tests.log; tail -3 gpurun_out/r02aa_tests.log
for mode in async async_cols; do
PA2S_GRU_EXCHANGE=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-steps 0 > gpurun_out/r02aa_bench_$mode.json 2> gpurun_out/r02aa_bench_$mode.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02aa_bench_$mode.json").read().strip().splitlines()[-1])
k=d["config"]["kernel_ms"]
print("$mode", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "gru fwd/bwd", k["encoder_gru_fwd"], k["encoder_gru_bwd"])
PY
done
