// Microbenchmarks of the instruction mixes in the walk kernel's hot loop (issue / pipe ceilings on B200).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../monte_carlompi_b200/csrc/mc3d_device.cuh"
using namespace mc3d;

// the round-1 generator (ten rounds, one block per event), kept here for the log this file produced
static __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t *__restrict__ rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, rk[2 * (r % PHILOX_ROUNDS)], rk[2 * (r % PHILOX_ROUNDS) + 1]);
    return make_uint4(c0, c1, c2, c3);
}

struct Keys { uint32_t rk[20]; };

template <int MODE>
__global__ void __launch_bounds__(256) bench(const __grid_constant__ Keys K, uint32_t iters, uint32_t *out)
{
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    float f0 = 1.0f + tid * 1e-9f, f1 = 0.5f, f2 = 0.25f;
    for (uint32_t i = 0; i < iters; ++i) {
        if (MODE == 0) {          // Philox4x32-10 only
            uint4 w = philox4x32_10(i, 0u, tid, 0u, K.rk);
            acc ^= w.x ^ w.y ^ w.z ^ w.w;
        } else if (MODE == 1) {   // 45 dependent-ish FFMA (2 chains)
#pragma unroll
            for (int k = 0; k < 22; ++k) { f0 = fmaf(f0, f1, f2); f1 = fmaf(f1, f2, f0); }
        } else if (MODE == 2) {   // 20 IMAD.WIDE only (2 chains)
            uint32_t a = i ^ tid, b = tid + 7u;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                uint64_t p = (uint64_t)a * 0xD2511F53u, q = (uint64_t)b * 0xCD9E8D57u;
                a = (uint32_t)(p >> 32) ^ (uint32_t)q; b = (uint32_t)(q >> 32) ^ (uint32_t)p;
            }
            acc ^= a ^ b;
        } else if (MODE == 3) {   // 40 LOP3 (2 chains)
            uint32_t a = i ^ tid, b = tid + 7u;
#pragma unroll
            for (int k = 0; k < 20; ++k) { a = (a ^ b ^ K.rk[k]) + 0u; b = (b & a) ^ K.rk[19 - k] ^ (a >> 1); }
            acc ^= a ^ b;
        } else if (MODE == 4) {   // Philox + 40 FFMA interleaved (independent)
            uint4 w = philox4x32_10(i, 0u, tid, 0u, K.rk);
#pragma unroll
            for (int k = 0; k < 20; ++k) { f0 = fmaf(f0, f1, f2); f1 = fmaf(f1, f2, f0); }
            acc ^= w.x ^ w.y ^ w.z ^ w.w;
        } else if (MODE == 5) {   // 6 MUFU + 10 FFMA
            float x = f0 + i;
            float a = __sinf(x), b = __cosf(x), c = __log2f(x), d = rsqrtf(x), e = __frcp_rn(x), g = sqrtf(x);
            f0 = fmaf(a, b, c); f1 = fmaf(d, e, g); f2 = fmaf(f0, f1, f2);
        }
    }
    out[tid] = acc + __float_as_uint(f0 + f1 + f2);
}

template <int MODE>
void run(const char *name, double instr_per_iter, int blocks_per_sm)
{
    Keys K;
    for (int i = 0; i < 20; ++i) K.rk[i] = 0x9E3779B9u * (i + 1);
    int sms = 148;
    const uint32_t iters = 20000;
    uint32_t *out;
    cudaMalloc(&out, sizeof(uint32_t) * sms * blocks_per_sm * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<sms * blocks_per_sm, 256>>>(K, 100, out);
    cudaEventRecord(e0);
    bench<MODE><<<sms * blocks_per_sm, 256>>>(K, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_iters = (double)sms * blocks_per_sm * 256 * iters;
    const double per_s = lane_iters / (ms * 1e-3);
    const double cyc_per_warp_iter_per_smsp = 1.965e9 / (per_s / 32 / (sms * 4));
    printf("%-34s warps/SMSP %2d: %.3e lane-iters/s, %.1f cycles per warp-iteration per SMSP (%.0f instr -> IPC %.2f)\n", name,
           blocks_per_sm * 2, per_s, cyc_per_warp_iter_per_smsp, instr_per_iter, instr_per_iter / cyc_per_warp_iter_per_smsp);
    cudaFree(out);
}

int main()
{
    for (int bps : {2, 4, 6, 8}) {
        run<0>("philox4x32-10", 45, bps);
        run<1>("44 FFMA", 46, bps);
        run<2>("20 IMAD.WIDE + 20 LOP3", 42, bps);
        run<3>("~60 LOP3/SHF/IADD", 62, bps);
        run<4>("philox + 40 FFMA", 87, bps);
        run<5>("6 MUFU + FFMA", 20, bps);
    }
    return 0;
}
