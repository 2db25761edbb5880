// fused_kernel.cu -- the short-walk path: one kernel, one thread per photon, nothing but the finished record touches
// HBM.
//
// Strongly absorbing grains (NIR wavelengths beyond ~1.5 um, large effective radii: config C3 of BASELINE.json) or
// optically thin slabs end a photon after a handful of events.  The persistent three-kernel path then spends its
// time moving per-photon state through HBM (16 B `fresh` written + read, 32 B raw result written + read, then the
// record: ~115 B per photon against 16-19 B of record) and on three launches.  Here the same device functions run
// back to back in registers: wavelength draw + first event (first_event, monte_carlo3D.py:1515-1520, 1035-1038,
// 1232-1237), groups of the walk stream with the termination chain resolved in place (group / resolve,
// monte_carlo3D.py:1212-1466), then angles, record and tallies (finalize_photon, monte_carlo3D.py:1468-1490,
// post_processing.py:73-76).  The photon's random stream depends only on (seed, photon id, event number), and the
// arithmetic is the same inlined code, so the records are bit-identical to the persistent path's (tested).
//
// A warp takes 32 consecutive photon ids; lanes that finish early idle until the longest walk of the 32 ends, which
// is cheap when walks are a few events long and is why the host only picks this kernel then (mc3d_api.cu:
// expected_events).  The 32 records of a warp are stored together (512 contiguous bytes in the packed form).
#include <algorithm>

#include "finalize_device.cuh"
#include "walk_device.cuh"

namespace mc3d {

template <bool IMP, bool SWEEP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) fused_kernel(const __grid_constant__ WalkParams P, const __grid_constant__ FinalizeParams F)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned int statics[6];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    const DevCase *cases = staged_cases(P, smem_raw);
    stage_tables(P, smem_raw, BLOCK);
    FinalizeBlock B;
    finalize_begin<BLOCK>(F, B, reinterpret_cast<unsigned int *>(smem_raw + tables_bytes(P.n_rows, P.n_cases)), statics);
    const uint32_t rows_addr = shared_address(rows);
    for (uint32_t base = blockIdx.x * BLOCK; base < P.n_photon; base += gridDim.x * BLOCK) {
        if (SWEEP) finalize_window<BLOCK>(F, B, base);
        const uint32_t pid = base + threadIdx.x;
        if (pid >= P.n_photon) continue;
        Lane L;
        uint32_t row = 0;
        float dtau = 0.0f;
        uint32_t redo;
        const uint32_t lcase = find_case<SWEEP>(P, cases, pid);
        const DevCase &C = SWEEP ? cases[lcase] : P.c;
        const uint64_t id = C.id0 + pid;
        const uint32_t phi = (uint32_t)(id >> 32);
        uint32_t cond = first_event<IMP>(P, C, phi, rows, rows_addr, (uint32_t)id, L, row, dtau, redo);
        while (cond == ALIVE) {
            if (!group<IMP, false, SWEEP>(P, rows, rows_addr, L)) cond = resolve<IMP>(P, C, phi, rows[row], L);
        }
        finalize_photon<SWEEP>(F, B, pid, L.ux, L.uy, L.uz, __fadd_rn(L.path_hi, L.path_lo), L.i - 1u, cond, row, row - C.row_begin,
                               lcase, rows[row].inv_ext);
    }
    finalize_flush<BLOCK>(F, B);
}

template <bool IMP, bool SWEEP>
static cudaError_t launch_fused_variant(const WalkParams &P, const FinalizeParams &F, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    FinalizeParams Q = F;
    const size_t tab_bytes = tables_bytes(P.n_rows, P.n_cases);
    // the tally block shares the SM with the tables: keep at least two blocks resident
    const size_t smem = tab_bytes + finalize_plan_smem(Q, tab_bytes < 64 * 1024 ? 96 * 1024 - tab_bytes : 0, 16 * 1024);
    auto kern = fused_kernel<IMP, SWEEP, BLOCK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int resident = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (resident < 1) return cudaErrorLaunchOutOfResources;
    const long long want = ((long long)P.n_photon + BLOCK - 1) / BLOCK;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * resident));
    kern<<<grid, BLOCK, smem, stream>>>(P, Q);
    return cudaGetLastError();
}

cudaError_t launch_fused(const WalkParams &P, const FinalizeParams &F, bool impurity, int sm_count, cudaStream_t stream)
{
    if (P.n_cases) return impurity ? launch_fused_variant<true, true>(P, F, sm_count, stream) : launch_fused_variant<false, true>(P, F, sm_count, stream);
    return impurity ? launch_fused_variant<true, false>(P, F, sm_count, stream) : launch_fused_variant<false, false>(P, F, sm_count, stream);
}

}  // namespace mc3d
