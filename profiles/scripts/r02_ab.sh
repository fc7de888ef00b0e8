# usage (GPU box): bash profiles/scripts/r02_ab.sh  -> gpurun_out/r02_ab.log ; every voxel_game_b200/build/var_*.so, twice
for rep in 1 2; do for v in voxel_game_b200/build/var_*.so; do VX_LIB_PATH=$v python profiles/scripts/r02_ab_kernels.py 30 2>&1 | tail -1; done; done | tee -a gpurun_out/r02_ab.log
