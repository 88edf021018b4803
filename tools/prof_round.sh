# One profiling round on the GPU box: ncu launch list of a timed bench step + `--set full` captures of the hot kernels.
# The .ncu-rep files stay on the box (gpurun_out/ is limited to 64 MiB): every capture is exported with `--page raw --csv`.
# usage: bash tools/prof_round.sh r02p      (outputs under gpurun_out/<tag>_*)
TAG=${1:-r02p}
mkdir -p gpurun_out /tmp/ncu
export PA2S_PROFILE_RANGE=1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also-steps 0"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
cap() {   # name regex count
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"$2" -c $3 -f -o /tmp/ncu/${TAG}_$1 $B > gpurun_out/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/ncu/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  tail -3 gpurun_out/${TAG}_$1.log > gpurun_out/${TAG}_$1.tail; rm -f gpurun_out/${TAG}_$1.log
}
cap conv "conv_tma3_kernel|conv_wgrad_tma_kernel|planes|conv1_" 17
cap decm_fwd decm_fwd_kernel 2
cap decm_bwd decm_bwd_kernel 2
cap decm_deferred decm_attn_deferred_kernel 2
cap gru "gru_seq_fwd_kernel|gru_seq_bwd2_kernel" 4
[ "$2" == "nogemm" ] || cap gemm "tc_gemm_tma_kernel" 45
ls -la gpurun_out | grep ${TAG}; du -sh gpurun_out
