// Microbenchmark: how fast can one B200 generate Philox-4x32-10 blocks and nothing else?  The floor under the fused
// kernel's generator (18 IMAD.WIDE + 19 LOP3 per sample once the high counter word is launch-uniform).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/philox_floor scripts/philox_floor.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../jetmontecarlo_b200/csrc/jmc_philox.cuh"
using namespace jmc;

template <int U, int EXTRA>
__global__ void __launch_bounds__(256) gen(uint64_t n, uint64_t seed, uint32_t* out) {
    PhiloxKeys keys(seed);
    keys.pin();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    uint32_t acc = 0;
    float facc = 0.f;
    for (uint64_t i = tid; i + (uint64_t)(U - 1) * stride < n; i += (uint64_t)U * stride) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
            uint32_t r[4];
            philox4x32_10((uint32_t)i + k * stride, 0u, 0u, keys, r);
            acc ^= r[0] ^ r[1] ^ r[2] ^ r[3];
            if (EXTRA) {   // the conversions + sampling FMAs of the fused kernel
                facc += fmaf(__uint2float_rn(r[0]), 1e-9f, 1.f) * fmaf(__uint2float_rn(r[1]), 1e-9f, 1.f)
                        + fmaf(__uint2float_rn(r[2]), 1e-9f, 1.f) * fmaf(__uint2float_rn(r[3]), 1e-9f, 1.f);
            }
        }
    }
    if (acc == 0x12345678u || facc == 1.2345f) out[tid] = acc;
}

template <int U, int EXTRA>
void run(const char* name, int blocks_per_sm) {
    uint32_t* out;
    cudaMalloc(&out, 1 << 24);
    const uint64_t n = 1000000000ull;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * blocks_per_sm;
    for (int w = 0; w < 2; ++w) gen<U, EXTRA><<<grid, 256>>>(n, 1234, out);
    cudaEventRecord(a);
    for (int w = 0; w < 5; ++w) gen<U, EXTRA><<<grid, 256>>>(n, 1234 + w, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    printf("%-28s U=%d blocks/SM=%d  %.3f ms per 1e9 blocks  %.3g blocks/s  %.1f SMSP-cycles per warp-block @1.965GHz\n", name, U,
           blocks_per_sm, ms, n / (ms * 1e-3), ms * 1e-3 * 1.965e9 / (n / 32.0 / (148 * 4)));
    cudaFree(out);
}

int main() {
    run<1, 0>("philox only", 4); run<1, 0>("philox only", 8);
    run<2, 0>("philox only", 4); run<2, 0>("philox only", 8);
    run<4, 0>("philox only", 2); run<4, 0>("philox only", 4); run<4, 0>("philox only", 8);
    run<2, 1>("philox + 4 cvt + 4 fma", 4); run<4, 1>("philox + 4 cvt + 4 fma", 4); run<4, 1>("philox + 4 cvt + 4 fma", 8);
    return 0;
}
