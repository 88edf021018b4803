#!/bin/bash
# round 2, first GPU call: full GPU test suite (incl. full-size parity), bench lines of every configuration, same-box GPU baseline
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
python bench.py --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_bf16_n1.json 2> gpurun_out/r02a_bench_bf16_n1.err
python bench.py --workload finetune --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_finetune_n1.json 2> gpurun_out/r02a_bench_finetune_n1.err
python bench.py --workload infer --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_infer_n1.json 2> gpurun_out/r02a_bench_infer_n1.err
timeout 600 python bench.py --impl reference-gpu --steps 2 --warmup 1 > gpurun_out/r02a_bench_refgpu_train.json 2> gpurun_out/r02a_bench_refgpu_train.err
timeout 600 python bench.py --impl reference-gpu --workload infer --batch 32 --steps 1 --warmup 1 > gpurun_out/r02a_bench_refgpu_infer.json 2> gpurun_out/r02a_bench_refgpu_infer.err
for f in gpurun_out/r02a_bench_*.json; do echo "== $f"; head -c 600 $f; echo; done
