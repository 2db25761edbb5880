// finalize_kernel.cu -- turns the walk kernel's raw per-photon results into the reference's output columns and
// the outcome / BRF tallies.  One thread per photon, fully coalesced, HBM-trivial (32 B in, <= 19 B out).
//
// Reference lines reproduced here:
//   theta_n = arccos(muz_0), phi_n by quadrant (== atan2 wrapped to [0, 2 pi)), 0 for an unscattered photon
//                                                              monte_carloMPI/monte_carlo3D.py:1468-1485
//   path_length [m] = sum dtau / (ext_cff_mss rho_snw)        monte_carlo3D.py:1355-1356, 1372
//   outcome tallies by condition (calculate_albedo)           monte_carlo3D.py:1659-1671
//   BRF zenith histogram of reflected photons, np.histogram(theta, bins=n, range=(0, pi/2))
//                                                              post_processing.py:73-76
//   optionally split in azimuth (np.histogram2d over (0, pi/2) x (0, 2 pi)): the full-hemisphere BRF the reference
//   stores the data for (phi_n) but never bins
//   optional histograms of n_scat and of path_length * scale over caller-given ranges, np.histogram(x, bins=n)
//                                                              post_processing.py:162-223
// Tallies are integer counts per wavelength row (the wvn weights of the reference are applied on the host in
// fp64), so they are exact and independent of the order of accumulation and of the GPU count.
#include "finalize_device.cuh"

namespace mc3d {

template <bool SWEEP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) finalize_kernel(const __grid_constant__ FinalizeParams P)
{
    extern __shared__ __align__(16) unsigned int smem_u32[];
    __shared__ __align__(8) unsigned int statics[6];
    float *inv_ext = reinterpret_cast<float *>(smem_u32);   // [n_rows]: metres per unit of path_tau, by SSP row
    for (int k = threadIdx.x; k < P.n_rows; k += BLOCK) inv_ext[k] = P.rows[k].inv_ext;
    uint32_t *row_begin = smem_u32 + P.n_rows;               // [n_cases] (sweep launches): first row of each case
    if (SWEEP)
        for (int k = threadIdx.x; k < (int)P.n_cases; k += BLOCK) row_begin[k] = P.cases[k].row_begin;
    FinalizeBlock B;
    finalize_begin<BLOCK>(P, B, smem_u32 + ((P.n_rows + (SWEEP ? P.n_cases : 0u) + 1u) & ~1u), statics);   // 8-byte aligned
    for (uint32_t base = blockIdx.x * BLOCK; base < P.n_photon; base += gridDim.x * BLOCK) {
        if (SWEEP) finalize_window<BLOCK>(P, B, base);
        const uint32_t p = base + threadIdx.x;
        if (p >= P.n_photon) continue;
        const RawResult *src = P.raw + p;
        const float4 a = *reinterpret_cast<const float4 *>(src);
        uint32_t n_scat, meta, lcase = 0u;
        if (SWEEP) {
            const uint4 b = *reinterpret_cast<const uint4 *>(&src->n_scat);
            n_scat = b.x; meta = b.y; lcase = b.z;
        } else {
            const uint2 b = *reinterpret_cast<const uint2 *>(&src->n_scat);
            n_scat = b.x; meta = b.y;
        }
        const uint32_t cond = meta & 0xffu, row = meta >> 8;
        finalize_photon<SWEEP>(P, B, p, a.x, a.y, a.z, a.w, n_scat, cond, row, SWEEP ? row - row_begin[lcase] : row, lcase, inv_ext[row]);
    }
    finalize_flush<BLOCK>(P, B);
}

template <bool SWEEP>
static cudaError_t launch_finalize_variant(const FinalizeParams &P, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    FinalizeParams Q = P;
    const size_t smem = (((size_t)P.n_rows + (SWEEP ? P.n_cases : 0u) + 1) & ~(size_t)1) * sizeof(float) + finalize_plan_smem(Q, 96 * 1024, 64 * 1024);
    auto kern = finalize_kernel<SWEEP, BLOCK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long want = ((long long)P.n_photon + BLOCK - 1) / BLOCK;
    const int per_sm = smem > 112 * 1024 ? 1 : (smem > 56 * 1024 ? 2 : 4);
    int grid = (int)(want < (long long)sm_count * per_sm ? (want > 0 ? want : 1) : (long long)sm_count * per_sm);
    kern<<<grid, BLOCK, smem, stream>>>(Q);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeParams &P, int sm_count, cudaStream_t stream)
{
    return P.cases ? launch_finalize_variant<true>(P, sm_count, stream) : launch_finalize_variant<false>(P, sm_count, stream);
}

}  // namespace mc3d
