# One profiling round on the GPU box: ncu launch list of a timed bench step + `--set full` captures of the hot kernels.
# usage: bash tools/prof_round.sh r01d      (outputs under gpurun_out/<tag>_*)
TAG=${1:-r01d}
mkdir -p gpurun_out
export PA2S_PROFILE_RANGE=1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tma_kernel|conv_wgrad_tma_kernel|planes" -c 15 -f -o gpurun_out/${TAG}_conv $B > gpurun_out/${TAG}_conv.log 2>&1
echo "conv rc=$?"
for k in dec_persist_fwd_kernel dec_persist_bwd_kernel gru_seq_fwd_kernel gru_seq_bwd_kernel; do
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${TAG}_$k $B > gpurun_out/${TAG}_$k.log 2>&1
echo "$k rc=$?"
done
ls -la gpurun_out
