#!/bin/bash
# round 2, 2-GPU call: NCCL / DDP parity tests, N=2 bench lines (bf16x3 + bf16, rank_consistent), same-box GPU baseline
mkdir -p gpurun_out; ls -la oracle/_ref oracle/_ref/data_processing > gpurun_out/r02b_ls.txt 2>&1
python -m pytest tests/test_gpu_nccl2.py -m gpu -x -q -s > gpurun_out/r02b_nccl2_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02b_nccl2_tests.log
tail -15 gpurun_out/r02b_nccl2_tests.log
for prec in bf16x3 bf16; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --precision $prec > gpurun_out/r02b_bench_${prec}_n2.json 2> gpurun_out/r02b_bench_${prec}_n2.err
done
timeout 900 python bench.py --impl reference-gpu --steps 2 --warmup 1 > gpurun_out/r02b_bench_refgpu_train.json 2> gpurun_out/r02b_bench_refgpu_train.err
timeout 900 python bench.py --impl reference-gpu --workload infer --batch 32 --steps 1 --warmup 1 > gpurun_out/r02b_bench_refgpu_infer.json 2> gpurun_out/r02b_bench_refgpu_infer.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_cpu.json 2> gpurun_out/r02b_bench_reference_cpu.err
for f in gpurun_out/r02b_bench_*.json; do echo "== $f"; head -c 400 $f; echo; done
