// Probe: does setmaxnreg.inc work for 16 of 20 warps when the pool has exactly the registers asked for?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int EPI, int CTRL>
__global__ void __launch_bounds__(640, 1) k(const float* __restrict__ in, float* __restrict__ out, int n) {
    const int warp = threadIdx.x >> 5;
    if (warp >= 16) {
        if (CTRL < 96) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL));
        if (warp == 16) out[threadIdx.x] = in[threadIdx.x] * 2.f;
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI));
        float v[80];
#pragma unroll
        for (int i = 0; i < 80; ++i) v[i] = in[threadIdx.x + i * 640];
        float s = 0.f;
        for (int it = 0; it < n; ++it) {
#pragma unroll
            for (int i = 0; i < 80; ++i) { v[i] = v[i] * v[(i + 1) % 80] + 1.0f; }
        }
#pragma unroll
        for (int i = 0; i < 80; ++i) s += v[i];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}
template <int EPI, int CTRL> void run(const float* in, float* out) {
    k<EPI, CTRL><<<148, 640>>>(in, out, 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("EPI %d CTRL %d: %s\n", EPI, CTRL, cudaGetErrorString(e));
}
int main(int argc, char** argv) {
    float *in, *out;
    cudaMalloc(&in, 640 * 120 * 4); cudaMalloc(&out, 640 * 4);
    cudaMemset(in, 0, 640 * 120 * 4);
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) run<104, 64>(in, out);
    if (which == 1) run<104, 80>(in, out);
    if (which == 2) run<112, 40>(in, out);
    if (which == 3) run<104, 88>(in, out);
    if (which == 4) run<120, 24>(in, out);
    if (which == 5) run<112, 56>(in, out);
    fflush(stdout);
    return 0;
}
