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
// Photons that are still walking after event 1 leave their state in the `fresh` list (entry = photon index; the
// others are marked dead); the walk kernel's lanes pick them up from there.  Photons that end on event 1 store their
// raw record here.  Keeping this code out of the walk kernel is what lets the walk kernel run at 48 registers per
// thread (40 resident warps per SM) without spilling in its event loop.
#include <algorithm>

#include "walk_device.cuh"

namespace mc3d {

template <bool IMP, bool SWEEP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) init_kernel(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    const DevCase *cases = staged_cases(P, smem_raw);
    stage_tables(P, smem_raw, BLOCK);
    __syncthreads();
    const uint32_t rows_addr = shared_address(rows);
    // Photon pid's entry of the `fresh` list is fresh[pid]: survivors of event 1 with their state, the others marked
    // FRESH_DEAD (the walk kernel's lanes skip those when they refill their ring).  No append counter: a list length is a
    // single address, and returning atomics on one address serialise in L2 at ~1 per ns -- one per warp took 0.27 ms
    // per 10^7 photons, one per block still stalled every block on its round trip
    // (profiles/r02_persistent_short_walk_launches.csv).  Coalesced 16-byte stores instead.
    for (uint32_t pid = blockIdx.x * BLOCK + threadIdx.x; pid < P.n_photon; pid += gridDim.x * BLOCK) {
        uint32_t redo = 0u, row = 0u;
        float dtau = 0.0f;
        Lane L;
        const uint32_t lcase = find_case<SWEEP>(P, cases, pid);
        const DevCase &C = SWEEP ? cases[lcase] : P.c;
        const uint64_t id = C.id0 + pid;
        const uint32_t cond = first_event<IMP>(P, C, (uint32_t)(id >> 32), rows, rows_addr, (uint32_t)id, L, row, dtau, redo);
        // cond == ALIVE with redo != 0: event 1 needed attention and the photon is still alive (reflected off a
        // Lambertian bottom on its first step, possibly through event 2 already; or a weakly absorbing row whose coarse
        // key asked for the fine test / a renormalisation): it is handed over in its state after the MOVE of event 1,
        // with the event's key and species in Fresh::redo, and the walk kernel's resolve pass redoes the chain of event
        // 1 (deterministic: same blocks, same result)
        Fresh f;
        f.pid = pid; f.row = cond == ALIVE ? row | (lcase << 12) : FRESH_DEAD; f.dtau = dtau; f.redo = redo;
        *reinterpret_cast<uint4 *>(P.fresh + pid) = *reinterpret_cast<uint4 *>(&f);
        if (cond != ALIVE) store_raw(P, pid, L.ux, L.uy, L.uz, __fadd_rn(L.path_hi, L.path_lo), L.i - 1u, cond, row, lcase);
    }
}

template <bool IMP, bool SWEEP>
static cudaError_t launch_init_variant(const WalkParams &P, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    const size_t smem = tables_bytes(P.n_rows, P.n_cases);
    const long long want = ((long long)P.n_photon + BLOCK - 1) / BLOCK;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * 8));
    auto kern = init_kernel<IMP, SWEEP, BLOCK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, BLOCK, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_init(const WalkParams &P, bool impurity, int sm_count, cudaStream_t stream)
{
    if (P.n_cases) return impurity ? launch_init_variant<true, true>(P, sm_count, stream) : launch_init_variant<false, true>(P, sm_count, stream);
    return impurity ? launch_init_variant<true, false>(P, sm_count, stream) : launch_init_variant<false, false>(P, sm_count, stream);
}

}  // namespace mc3d
