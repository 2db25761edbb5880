// Pure hot-loop throughput: the walk kernel's event() with termination disabled, all lanes busy, no refill.
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
    for (int k = threadIdx.x; k < P.n_rows * (int)(sizeof(DevRow) / 4); k += 256)
        reinterpret_cast<uint32_t *>(rows)[k] = reinterpret_cast<const uint32_t *>(P.rows)[k];
    __syncthreads();
    const uint32_t rows_addr = shared_address(rows);
    Lane L;
    const uint32_t tid = blockIdx.x * 256 + threadIdx.x;
    L.z = -50.0f; L.ux = 0.26f; L.uy = 0.f; L.uz = -0.9659f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 1; L.plo = tid; L.row_addr = rows_addr + (tid % P.n_rows) * 48; L.w3 = 0; L.imp = false;
    L.pk = philox_event_constants(L.plo, P.rk);
    uint32_t stops = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        if (MODE == 0) {
            const bool alive = event<false>(P, rows, rows_addr, L);
            stops += alive ? 0u : 1u;
            if (L.z > -10.0f) L.z -= 40.0f;     // keep the photon deep inside: never exits
        } else if (MODE == 1) {                   // Philox only (hoisted form)
            const uint4 w = philox_event(L.i + 1u, 0u, L.pk, P.rk);
            L.i += 1u; stops += w.x ^ w.y ^ w.z ^ w.w;
        } else {                                  // math only: draws from a cheap counter hash
            uint4 w; uint32_t x = (L.i + 1u) * 0x9E3779B9u ^ L.plo; w.x = x * 0x85EBCA6Bu; w.y = x ^ (x >> 15); w.z = w.x ^ (w.y << 3); w.w = w.z + x;
            L.i += 1u;
            const HotRow H = load_hot_row(L.row_addr);
            scatter_and_move(L, H, w);
            if (L.z > -10.0f) L.z -= 40.0f;
            stops += (L.z < P.neg_tau_tot) ? 1u : 0u;
        }
    }
    out[tid] = L.z + L.ux + L.path_lo + stops;
}

int main()
{
    WalkParams P; memset(&P, 0, sizeof P);
    for (int i = 0; i < 20; ++i) P.rk[i] = 0x9E3779B9u * (i + 3);
    P.neg_tau_tot = -1e30f; P.tau_tot = 1e30f; P.n_rows = 53; P.refill_threshold = 4;
    DevRow h[53];
    for (int r = 0; r < 53; ++r) { double g = 0.89; h[r].one_m_g = 1 - g; h[r].one_m_g2 = 1 - g * g; h[r].d_scale = (float)(2 * g / 4294967296.0); h[r].flip = 0; h[r].t_hi = 0xffffffffu; h[r].d_off = (float)(1 - g); h[r].t_lo = 256; h[r].ti_hi = 0xffffffffu; h[r].ti_lo = 256; h[r].s_last = 0; h[r].s_any = 0; h[r].inv_ext = 1e-3f; }
    DevRow *drows; cudaMalloc(&drows, sizeof h); cudaMemcpy(drows, h, sizeof h, cudaMemcpyHostToDevice);
    P.rows = drows;
    float *out; cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    const uint32_t iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[3] = {"event() full hot loop", "philox_event only", "scatter_and_move only"};
    for (int bps = 1; bps <= 4; ++bps) for (int mode = 0; mode < 3; ++mode) {
        auto launch = [&](uint32_t n) {
            if (mode == 0) hot<0><<<148 * bps, 256, 53 * 48>>>(P, n, out);
            else if (mode == 1) hot<1><<<148 * bps, 256, 53 * 48>>>(P, n, out);
            else hot<2><<<148 * bps, 256, 53 * 48>>>(P, n, out);
        };
        launch(100);
        cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double lane_iters = 148.0 * bps * 256 * iters, per_s = lane_iters / (ms * 1e-3);
        printf("%-26s warps/SMSP %d: %.3e events/s  %.1f cycles per warp-iteration per SMSP\n", names[mode], bps * 2, per_s, 1.965e9 / (per_s / 32 / 592));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
