# Source-level (SASS) ncu capture of ONE launch of a kernel inside a timed bench step: where do the warps stall?
# usage: bash tools/prof_source.sh <tag> <kernel regex> [skip]
TAG=${1:-src}; K=${2:-conv_tma_kernel}; SKIP=${3:-0}
mkdir -p gpurun_out /tmp/ncu
export PA2S_PROFILE_RANGE=1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also-steps 0"
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"$K" -s $SKIP -c 1 -f -o /tmp/ncu/${TAG} $B > gpurun_out/${TAG}.log 2>&1
echo "rc=$?"
ncu -i /tmp/ncu/${TAG}.ncu-rep --page source --csv > /tmp/ncu/${TAG}_source.csv 2>/dev/null
ncu -i /tmp/ncu/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
gzip -c /tmp/ncu/${TAG}_source.csv > gpurun_out/${TAG}_source.csv.gz
ls -la gpurun_out/${TAG}*
