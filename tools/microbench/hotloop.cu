// Pure hot-loop throughput: the walk kernel's group() (four events on three Philox4x32-7 blocks) with termination
// disabled, all lanes busy, no refill.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../monte_carlompi_b200/csrc -o hotloop hotloop.cu
#include <cstdio>
#include <cstring>
#include "walk_device.cuh"
using namespace mc3d;

template <int MODE>
__global__ void __launch_bounds__(256, 4) hot(const __grid_constant__ WalkParams P, uint32_t iters, float *out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    for (int k = threadIdx.x; k < P.n_rows * (int)(sizeof(DevRow) / 4); k += blockDim.x)
        reinterpret_cast<uint32_t *>(rows)[k] = reinterpret_cast<const uint32_t *>(P.rows)[k];
    __syncthreads();
    const uint32_t rows_addr = shared_address(rows);
    Lane L;
    const uint32_t tid = blockIdx.x * 256 + threadIdx.x;
    L.z = -50.0f; L.ux = 0.26f; L.uy = 0.f; L.uz = -0.9659f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 1; L.blk = 0; L.plo = tid; L.row_addr = rows_addr + (tid % P.n_rows) * (uint32_t)sizeof(DevRow); L.key = 0; L.imp = false;
    L.pk = philox_walk_constants(L.plo, P.rk);
    uint32_t stops = 0;
    for (uint32_t it = 0; it < iters; ++it) {          // one group = four events per iteration
        if (MODE == 0 || MODE == 3 || MODE == 4) {
            const bool alive = MODE == 0 ? group<false, false>(P, rows, rows_addr, L) : MODE == 3 ? group<false, true>(P, rows, rows_addr, L) : group_latency<false>(P, rows, rows_addr, L);
            stops += alive ? 0u : 1u;
            if (L.z > -10.0f) L.z -= 40.0f;     // keep the photon deep inside: never exits
        } else if (MODE == 1) {                   // Philox only: three blocks
            const uint4 a = philox_walk(L.blk, 0u, L.pk, P.rk), b = philox_walk(L.blk + 1u, 0u, L.pk, P.rk), c = philox_walk(L.blk + 2u, 0u, L.pk, P.rk);
            L.blk += 3u; stops += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w;
        } else {                                  // math only: draws from a cheap counter hash
            const HotRow H = load_hot_row(L.row_addr);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t x = (L.i + 1u) * 0x9E3779B9u ^ L.plo; const uint32_t w0 = x * 0x85EBCA6Bu, w1 = x ^ (x >> 15), w2 = w0 ^ (w1 << 3);
                L.i += 1u;
                scatter_and_move(L, H, w0, w1, w2);
                if (L.z > -10.0f) L.z -= 40.0f;
                stops += (L.z < P.neg_tau_tot) ? 1u : 0u;
            }
        }
    }
    out[tid] = L.z + L.ux + L.path_lo + stops;
}

int main()
{
    WalkParams P; memset(&P, 0, sizeof P);
    for (int i = 0; i < 2 * PHILOX_ROUNDS; ++i) P.rk[i] = 0x9E3779B9u * (i + 3);
    P.neg_tau_tot = -1e30f; P.tau_tot = 1e30f; P.n_rows = 53; P.refill_threshold = 4;
    DevRow h[53]; memset(h, 0, sizeof h);
    for (int r = 0; r < 53; ++r) { double g = 0.89; h[r].one_m_g = 1 - g; h[r].one_m_g2 = 1 - g * g; h[r].d_scale = (float)(2 * g / 4294967296.0); h[r].d_off = (float)(1 - g); h[r].omr_scale = -(float)(1 / 4294967296.0); h[r].omr_off = 1.0f; h[r].t_hot = 0xffffffffu; h[r].ti_hot = 0xffffffffu; h[r].t16 = 0x10000; h[r].ti16 = 0x10000; h[r].inv_ext = 1e-3f; }
    DevRow *drows; cudaMalloc(&drows, sizeof h); cudaMemcpy(drows, h, sizeof h, cudaMemcpyHostToDevice);
    P.rows = drows;
    float *out; cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    const uint32_t iters = 5000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[5] = {"group() full hot loop", "3 x philox_walk only", "4 x scatter_and_move only", "group() eager blocks", "group_latency()"};
    // latency: ONE warp per scheduler (148 blocks x 128 threads), the regime of a launch's tail
    for (int mode : {0, 3, 4}) {
        auto launch = [&](uint32_t n) {
            if (mode == 0) hot<0><<<148, 128, sizeof h>>>(P, n, out);
            else if (mode == 3) hot<3><<<148, 128, sizeof h>>>(P, n, out);
            else hot<4><<<148, 128, sizeof h>>>(P, n, out);
        };
        launch(100);
        cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-26s 1 warp/SMSP (latency): %.1f cycles per event\n", names[mode], ms * 1e-3 * 1.965e9 / (iters * 4.0));
    }
    for (int bps = 1; bps <= 4; ++bps) for (int mode = 0; mode < 4; ++mode) {
        auto launch = [&](uint32_t n) {
            if (mode == 0) hot<0><<<148 * bps, 256, sizeof h>>>(P, n, out);
            else if (mode == 1) hot<1><<<148 * bps, 256, sizeof h>>>(P, n, out);
            else if (mode == 2) hot<2><<<148 * bps, 256, sizeof h>>>(P, n, out);
            else hot<3><<<148 * bps, 256, sizeof h>>>(P, n, out);
        };
        launch(100);
        cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double lane_events = 148.0 * bps * 256 * iters * 4.0, per_s = lane_events / (ms * 1e-3);
        printf("%-26s warps/SMSP %d: %.3e events/s  %.1f cycles per warp-event per SMSP\n", names[mode], bps * 2, per_s, 1.965e9 / (per_s / 32 / 592));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
