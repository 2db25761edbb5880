// Latency of one photon's dependent chain, one warp alone on an SM sub-partition (the regime of a launch's tail):
// cycles per event of (A) apply_event alone with prepared inputs, (B) a bare rsqrt chain, (C) a bare FFMA chain,
// (D) apply_event + the attention predicate with select masking as the latency-oriented loops use it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../monte_carlompi_b200/csrc -o chain chain.cu
#include <cstdio>
#include <cstring>
#include "walk_device.cuh"
using namespace mc3d;

template <int MODE>
__global__ void chain(uint32_t iters, float seed, float *out, long long *cycles)
{
    Lane L;
    L.z = -50.0f; L.ux = 0.26f + seed; L.uy = 0.01f; L.uz = -0.9659f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 1; L.blk = 0; L.plo = threadIdx.x; L.phi = 0; L.row_addr = 0; L.key = 0; L.imp = false;
    Prepared e[4];
    for (int k = 0; k < 4; ++k) { e[k].ct = 0.95f - 0.01f * k + seed; e[k].st2 = 1.0f - e[k].ct * e[k].ct; e[k].cp = 0.6f + 0.05f * k; e[k].sp = sqrtf(1.0f - e[k].cp * e[k].cp); e[k].dtau = 0.5f + seed; e[k].key = k; }
    float x = 1.5f + seed;
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) apply_event(L, e[k]);
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) x = rsqrt_fast(x) + 1.0f;
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) x = fmaf(x, 0.999f, 0.5f);
        } else {
            bool go = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                Lane N = L; N.i += 1u; apply_event(N, e[k]);
                const bool att = needs_attention(-1e30f, N, 0xfffffff0u);
                L.z = go ? N.z : L.z; L.ux = go ? N.ux : L.ux; L.uy = go ? N.uy : L.uy; L.uz = go ? N.uz : L.uz;
                L.path_lo = go ? N.path_lo : L.path_lo; L.key = go ? N.key : L.key; L.i = go ? N.i : L.i;
                go = go && !att;
            }
            if (L.z < -1e20f) L.z = 0.f;
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = L.z + L.ux + L.uy + L.uz + L.path_lo + x + L.key;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main()
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const uint32_t iters = 20000;
    const char *names[4] = {"apply_event alone", "rsqrt.approx + fadd", "ffma", "apply_event + predicate + select masking"};
    for (int m = 0; m < 4; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) chain<0><<<1, 32>>>(iters, 0.f, out, cyc);
            if (m == 1) chain<1><<<1, 32>>>(iters, 0.f, out, cyc);
            if (m == 2) chain<2><<<1, 32>>>(iters, 0.f, out, cyc);
            if (m == 3) chain<3><<<1, 32>>>(iters, 0.f, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %7.1f cycles per event (one warp, 4 events per iteration)\n", names[m], (double)h / (iters * 4.0));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
