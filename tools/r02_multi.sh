#!/bin/bash
# multi-GPU bench lines of every configuration at N GPUs of one box: usage bash tools/r02_multi.sh N tag [tests]
N=$1; TAG=$2
mkdir -p gpurun_out
if [ "$3" == "tests" ]; then
python -m pytest tests/test_gpu_nccl2.py -m gpu -q -s > gpurun_out/${TAG}_nccl2_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_nccl2_tests.log
tail -4 gpurun_out/${TAG}_nccl2_tests.log
fi
run() {  # name args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.out 2> gpurun_out/${TAG}_${name}_n$N.err
  tail -1 gpurun_out/${TAG}_${name}_n$N.out > gpurun_out/${TAG}_${name}_n$N.json; rm -f gpurun_out/${TAG}_${name}_n$N.out
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${name}_n$N.json"))
    print("$name N=$N", round(d["value"],1), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "rank_consistent", d.get("rank_consistent"), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$name FAILED", e)
PY
}
# 4th argument: "short" = bf16x3 + bf16 only, "headline" = bf16x3 + greedy decode only, default all four
run bf16x3 --steps 8 --warmup 3 --also-steps 0
[ "$4" == "headline" ] || run bf16 --steps 8 --warmup 3 --precision bf16
[ "$4" == "short" ] || [ "$4" == "headline" ] || run finetune --steps 8 --warmup 3 --workload finetune
[ "$4" == "short" ] || run infer --steps 4 --warmup 3 --workload infer
