#!/bin/bash
# Round evidence, run on the GPU box AFTER a plain bench.py run has exited 0:
#   1. launch list of the bench command (durations only, cold-cache and serialised)
#   2. one `--set full` capture of every kernel the command launches (first launch after the warm-up)
# usage: bash profiles/scripts/capture.sh <tag>     -> gpurun_out/launches_<tag>.csv, prof_<tag>.ncu-rep
cd "$(dirname "$0")/../.." || exit 1
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_l_$tag.log 2>&1
# skip the three warm-up steps (6 kernels each + mesh_prepare on the first): one timed step + the fp64 probe
ncu --set full --clock-control none --import-source on --launch-skip 19 --launch-count 7 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
