// init_kernel.cu -- the per-photon prologue, one thread per photon, fully coalesced.
//
//   * wavelength draw: wvls = np.around(np.random.normal(wvl0, scale), 2)          reference monte_carlo3D.py:1515-1520
//     (Box-Muller on the photon's own Philox block; the rounded value is an index into the SSP table)
//   * Lambertian_surface mode (monte_carlo3D.py:1228-1229, 1238-1250, 1385-1387): the snow is replaced by a
//     Lambertian reflector, every photon ends after one or two events, so it is finished right here
//   * first event: the three draws of initial_pdfs (monte_carlo3D.py:1035-1038), no deflection (1232-1237),
//     move, direct-transmission / Lambertian-bottom / first-extinction absorption tests (1399-1466)
//   (all of it in walk_device.cuh: first_event(), shared with the fused kernel)
//
// Photons that are still walking after event 1 are appended to the `fresh` list (warp-aggregated append: one
// atomicAdd per warp); the walk kernel's lanes pick them up from there.  Photons that end on event 1 store their
// raw record here.  Keeping this code out of the walk kernel is what lets the walk kernel run at 48 registers per
// thread (40 resident warps per SM) without spilling in its event loop.
#include <algorithm>

#include "walk_device.cuh"

namespace mc3d {

template <bool IMP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) init_kernel(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    for (int k = threadIdx.x; k < P.n_rows * (int)(sizeof(DevRow) / 4); k += BLOCK)
        reinterpret_cast<uint32_t *>(rows)[k] = reinterpret_cast<const uint32_t *>(P.rows)[k];
    __syncthreads();
    const uint32_t rows_addr = shared_address(rows);
    const uint32_t lane = threadIdx.x & 31u;
    // whole warps iterate together so that the ballot below is always executed by all 32 lanes
    for (uint32_t base = (blockIdx.x * BLOCK + (threadIdx.x & ~31u)); base < P.n_photon; base += gridDim.x * BLOCK) {
        const uint32_t pid = base + lane;
        bool survive = false;
        float dtau = 0.0f;
        uint32_t row = 0;
        if (pid < P.n_photon) {
            Lane L;
            uint32_t cond = first_event<IMP>(P, rows, rows_addr, (uint32_t)P.photon_begin + pid, L, row, dtau);
            if (cond == ALIVE && L.i != 1u) {
                // reflected off the Lambertian bottom on its first step and still alive after event 2: it no
                // longer has the "fresh photon" state, so it is walked to completion here (thin slabs only; those
                // normally take the fused kernel anyway)
                do {
                    if (!group<IMP, true>(P, rows, rows_addr, L)) cond = resolve<IMP>(P, rows[row], L);
                } while (cond == ALIVE);
            }
            if (cond == ALIVE) survive = true;
            else store_raw(P, pid, L.ux, L.uy, L.uz, L.path_hi + L.path_lo, L.i - 1u, cond, row);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, survive);
        uint32_t slot0 = 0;
        if (lane == 0 && m) slot0 = atomicAdd(P.n_fresh, (uint32_t)__popc(m));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (survive) {
            Fresh f;
            f.pid = pid; f.row = row; f.dtau = dtau; f.pad = 0u;
            *reinterpret_cast<uint4 *>(P.fresh + slot0 + __popc(m & ((1u << lane) - 1u))) = *reinterpret_cast<uint4 *>(&f);
        }
    }
}

cudaError_t launch_init(const WalkParams &P, bool impurity, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    const size_t smem = (size_t)P.n_rows * sizeof(DevRow);
    const long long want = ((long long)P.n_photon + BLOCK - 1) / BLOCK;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * 8));
    if (impurity) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(init_kernel<true, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        init_kernel<true, BLOCK><<<grid, BLOCK, smem, stream>>>(P);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(init_kernel<false, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        init_kernel<false, BLOCK><<<grid, BLOCK, smem, stream>>>(P);
    }
    return cudaGetLastError();
}

}  // namespace mc3d
