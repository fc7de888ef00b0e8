// issue_model.cu -- does an integer instruction hide in the shadow of an FP64 instruction on B200?
// Each FP64 warp instruction occupies the 16-lane FP64 pipe of an SM sub-partition for 2 cycles.
// This probe times, per loop step, F independent DADD/DMUL (no FMA) plus I independent integer
// LOP3/IADD, and reports cycles per step and sub-partition, to tell "max(F*2, F+I)" from "F*2 + I".
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o issue_model issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int F, int I>
__global__ void __launch_bounds__(256) probe(double* sink, unsigned* isink, int iters) {
    double a[8];
    unsigned u[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + threadIdx.x * 1e-9 + k * 0.125;
#pragma unroll
    for (int k = 0; k < 16; ++k) u[k] = threadIdx.x * 2654435761u + k;
    const double m = 0.9999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < F; ++k) a[k % 8] = (k & 8) ? a[k % 8] + c : a[k % 8] * m;
#pragma unroll
        for (int k = 0; k < I; ++k) u[k % 16] = (u[k % 16] ^ (u[(k + 1) % 16] >> 3)) + 0x9e3779b9u;  // 2 ops: LOP3(shift folded? no) + IADD
    }
    double r = 0;
    unsigned q = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += a[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) q ^= u[k];
    if (r == 12345.678) sink[0] = r;
    if (q == 0x12345u) isink[0] = q;
}

template <int F, int I>
void run(double* sink, unsigned* isink, int sms, double mhz) {
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<F, I><<<blocks, 256>>>(sink, isink, iters);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        probe<F, I><<<blocks, 256>>>(sink, isink, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // warps per sub-partition = blocks * 8 / (sms * 4) = 16; cycles per step per sub-partition:
    const double cyc = best * 1e-3 * mhz * 1e6 / ((double)iters * 16.0);
    // one integer step is 3 instructions (SHF, LOP3, IADD)
    printf("F=%2d fp64 + %2d int instr per step: %.3f ms = %.1f cycles per warp-step;  max(2F, F+I)=%d  2F+I=%d\n", F,
           3 * I, best, cyc, (2 * F > F + 3 * I ? 2 * F : F + 3 * I), 2 * F + 3 * I);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s  SMs=%d  clock=%.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    double* sink;
    unsigned* isink;
    cudaMalloc(&sink, 8);
    cudaMalloc(&isink, 4);
    run<16, 0>(sink, isink, p.multiProcessorCount, mhz);
    run<0, 16>(sink, isink, p.multiProcessorCount, mhz);
    run<16, 8>(sink, isink, p.multiProcessorCount, mhz);
    run<16, 16>(sink, isink, p.multiProcessorCount, mhz);
    run<8, 16>(sink, isink, p.multiProcessorCount, mhz);
    return 0;
}
