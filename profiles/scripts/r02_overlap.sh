cp voxel_game_b200/libvoxel_cuda.so /tmp/keep.so
for v in voxel_game_b200/build/var_*.so; do VX_LIB=$v python profiles/scripts/r02_overlap.py 30; done 2>&1 | grep contexts | tee gpurun_out/r02_overlap.log
cp /tmp/keep.so voxel_game_b200/libvoxel_cuda.so
