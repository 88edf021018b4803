#!/bin/bash
mkdir -p gpurun_out
for bc in 1 0; do
PA2S_BAR_CHAIN=$bc PA2S_TIMELINE=1 PA2S_TIMELINE_WINDOW="0,100" timeout 300 python tools/trace_step.py --top 12 > gpurun_out/r02af_timeline_bc$bc.txt 2>&1
grep "step span" gpurun_out/r02af_timeline_bc$bc.txt
done
