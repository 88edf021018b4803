#!/bin/bash
# kernel timeline of one training step with the bar-level chain as one node per bar (default) and as individual autograd ops
# usage: gpurun -- bash tools/timeline_compare.sh   (profiles/r02af_timeline.txt is the first of the two outputs)
mkdir -p gpurun_out
for bc in 1 0; do
PA2S_BAR_CHAIN=$bc PA2S_TIMELINE=1 PA2S_TIMELINE_WINDOW="0,100" timeout 300 python tools/trace_step.py --top 12 > gpurun_out/cmp_timeline_bc$bc.txt 2>&1
grep "step span" gpurun_out/cmp_timeline_bc$bc.txt
done
