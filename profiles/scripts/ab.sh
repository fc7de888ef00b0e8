#!/bin/bash
# A/B harness: bench every library variant under voxel_game_b200/build/var_*.so on the same box.
# Usage (on the GPU box): bash profiles/scripts/ab.sh [steps]
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
cp voxel_game_b200/libvoxel_cuda.so /tmp/keep.so
for v in voxel_game_b200/build/var_*.so; do
  cp "$v" voxel_game_b200/libvoxel_cuda.so
  for rep in 1 2; do
    python bench.py --steps ${1:-30} --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        k = d['kernel_ms_per_step']
        print('$v', 'e2e %.0f (%.2f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'ms/step %.4f' % d['ms_per_step'], ' '.join('%s=%.4f' % (a.replace('_ms', ''), b) for a, b in k.items()), 'clk', d['clocks']['sm_mhz'])
"
  done
done | tee gpurun_out/ab.log
cp /tmp/keep.so voxel_game_b200/libvoxel_cuda.so
