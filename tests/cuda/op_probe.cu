// op_probe.cu -- TEST INFRASTRUCTURE: the Fbm loops of csrc/vx_noise.cuh compiled on their own, so
// that tests/test_op_counts.py can count the f64 instructions ptxas emits per simplex octave (the
// numerators F2_OPS / F3_OPS of bench.py's FP64 rooflines) from the SASS instead of by hand.
// Compiled with the library's own flags (-fmad=false, sm_100a); never linked into the product.
#include "../../voxel_game_b200/csrc/vx_noise.cuh"

using namespace vx;

// one 2-D octave per trip of the only loop of this kernel
extern "C" __global__ void __launch_bounds__(256) probe_fbm2(const __grid_constant__ NoiseTables nt, double* io) {
    __shared__ __align__(256) NoiseShared s;
    load_noise_shared<FBM_TABLE<FBM_ALT>>(&s, nt, threadIdx.x, 256);
    __syncthreads();
    io[threadIdx.x] = fbm2<FBM_HEIGHT1>(nt, &s, io[threadIdx.x], io[threadIdx.x + 256]);
}

// one 3-D octave per trip
extern "C" __global__ void __launch_bounds__(256) probe_fbm3(const __grid_constant__ NoiseTables nt, double* io) {
    __shared__ __align__(256) NoiseShared s;
    load_noise_shared<FBM_TABLE<FBM_ALT>>(&s, nt, threadIdx.x, 256);
    __syncthreads();
    io[threadIdx.x] = fbm3<FBM_ALT>(nt, &s, io[threadIdx.x], io[threadIdx.x + 256], io[threadIdx.x + 512]);
}
