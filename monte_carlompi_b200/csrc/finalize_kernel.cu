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
#include "mc3d_device.cuh"

namespace mc3d {

// numpy/lib/_histograms_impl.py, uniform-bin fast path, applied to float64(theta_f32)
__device__ __forceinline__ int histogram_bin(double x, int n_bins, const double *__restrict__ edges)
{
    const double first = edges[0], last = edges[n_bins];
    if (!(x >= first && x <= last)) return -1;
    const double f = __dmul_rn(__ddiv_rn(__dsub_rn(x, first), __dsub_rn(last, first)), (double)n_bins);
    int idx = (int)f;
    if (idx == n_bins) idx -= 1;
    if (x < edges[idx]) idx -= 1;
    else if (x >= edges[idx + 1] && idx != n_bins - 1) idx += 1;
    return idx;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) finalize_kernel(const __grid_constant__ FinalizeParams P)
{
    extern __shared__ unsigned int smem_u32[];
    float *inv_ext = reinterpret_cast<float *>(smem_u32);   // [n_rows]: metres per unit of path_tau, by SSP row
    unsigned int *hist = smem_u32 + P.n_rows;               // [n_rows][N_COND + n_theta_bins * n_phi] when use_smem
    __shared__ unsigned int block_ext[4];                   // extrema of the block (minima complemented)
    for (int k = threadIdx.x; k < P.n_rows; k += BLOCK) inv_ext[k] = P.rows[k].inv_ext;
    __shared__ unsigned long long block_events;
    if (threadIdx.x < 4) block_ext[threadIdx.x] = 0u;
    if (threadIdx.x == 0) block_events = 0ull;
    const int n_phi = P.n_phi_bins > 1 ? P.n_phi_bins : 1;
    const int stride = N_COND + P.n_theta_bins * n_phi;
    const int hist_len = P.n_rows * stride;
    const bool tally = P.tally != nullptr;
    // the optional column histograms sit behind the tally block in shared memory
    const int xh_len = P.hist ? P.n_scat_bins + P.path_bins : 0;
    unsigned int *xh = hist + (tally && P.use_smem ? hist_len : 0);
    {
        const int len = (tally && P.use_smem ? hist_len : 0) + (P.hist_smem ? xh_len : 0);
        for (int k = threadIdx.x; k < len; k += BLOCK) hist[k] = 0u;
        __syncthreads();
    }
    unsigned long long events = 0ull;
    uint32_t ns_min = 0xffffffffu, ns_max = 0u, pl_min = 0xffffffffu, pl_max = 0u;
    for (uint32_t p = blockIdx.x * BLOCK + threadIdx.x; p < P.n_photon; p += gridDim.x * BLOCK) {
        const RawResult *src = P.raw + p;
        const float4 a = *reinterpret_cast<const float4 *>(src);
        const uint2 b = *reinterpret_cast<const uint2 *>(&src->n_scat);
        const uint32_t cond = b.y & 0xffu, row = b.y >> 8;
        const float theta = atan2f(sqrtf(fmaf(a.x, a.x, a.y * a.y)), a.z);
        float phi = 0.0f;
        if (b.x != 0u) {
            phi = atan2f(a.y, a.x);
            if (phi < 0.0f) phi += 6.283185307179586f;
        }
        // fp32 rounding can leave a path of (nearly) zero length slightly negative (Lambertian surface, immediate exit)
        const float path_m = fmaxf(a.w * inv_ext[row], 0.0f);
        if (P.packed) {
            P.packed[p] = make_uint4((min(b.x, 0x7fffffu) << 9) | row, __float_as_uint(theta) | (cond << 31),
                                     __float_as_uint(phi) | ((cond >> 1) << 31), __float_as_uint(path_m) | ((cond >> 2) << 31));
        } else {
            if (P.condition) P.condition[p] = (uint8_t)cond;
            if (P.wvl_row) P.wvl_row[p] = (int16_t)row;
            if (P.theta_n) P.theta_n[p] = theta;
            if (P.phi_n) P.phi_n[p] = phi;
            if (P.n_scat) P.n_scat[p] = b.x;
            if (P.path_length) P.path_length[p] = path_m;
        }
        events += (unsigned long long)b.x + 1ull;
        ns_min = min(ns_min, b.x);
        ns_max = max(ns_max, b.x);
        pl_min = min(pl_min, __float_as_uint(path_m));
        pl_max = max(pl_max, __float_as_uint(path_m));
        if (xh_len) {
            int hb[2] = {-1, -1};
            if (P.n_scat_bins > 0) hb[0] = histogram_bin((double)b.x, P.n_scat_bins, P.hist_edges);
            if (P.path_bins > 0) {
                hb[1] = histogram_bin(__dmul_rn((double)path_m, P.path_scale), P.path_bins, P.hist_edges + P.n_scat_bins + 1);
                if (hb[1] >= 0) hb[1] += P.n_scat_bins;
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (hb[k] < 0) continue;
                if (P.hist_smem) atomicAdd(&xh[hb[k]], 1u);
                else atomicAdd(&P.hist[hb[k]], 1ull);
            }
        }
        if (tally) {
            const int base = (int)row * stride;
            int bin = -1;
            if (cond == 1u && P.n_theta_bins > 0) {
                bin = histogram_bin((double)theta, P.n_theta_bins, P.edges);
                if (bin >= 0 && n_phi > 1) {   // np.histogram2d: a sample outside either range is dropped
                    const int pb = histogram_bin((double)phi, n_phi, P.edges + P.n_theta_bins + 1);
                    bin = pb >= 0 ? bin * n_phi + pb : -1;
                }
            }
            if (P.use_smem) {
                atomicAdd(&hist[base], 1u);
                atomicAdd(&hist[base + cond], 1u);
                if (bin >= 0) atomicAdd(&hist[base + N_COND + bin], 1u);
            } else {
                atomicAdd(&P.tally[base], 1ull);
                atomicAdd(&P.tally[base + cond], 1ull);
                if (bin >= 0) atomicAdd(&P.tally[base + N_COND + bin], 1ull);
            }
        }
    }
    // events: warp reduce, one shared-memory atomic per warp, one global atomic per block (below)
    for (int o = 16; o > 0; o >>= 1) events += __shfl_xor_sync(0xffffffffu, events, o);
    if ((threadIdx.x & 31) == 0 && events) atomicAdd(&block_events, events);
    if (P.extrema) {   // warp reduce -> one shared-memory atomic per warp -> one global atomic per block
        ns_min = __reduce_min_sync(0xffffffffu, ns_min);
        ns_max = __reduce_max_sync(0xffffffffu, ns_max);
        pl_min = __reduce_min_sync(0xffffffffu, pl_min);
        pl_max = __reduce_max_sync(0xffffffffu, pl_max);
        if ((threadIdx.x & 31) == 0 && ns_min <= ns_max) {
            atomicMax(&block_ext[0], ~ns_min);   // minima are stored complemented: the buffer starts as zeros
            atomicMax(&block_ext[1], ns_max);
            atomicMax(&block_ext[2], ~pl_min);
            atomicMax(&block_ext[3], pl_max);
        }
    }
    __syncthreads();
    if (P.extrema && threadIdx.x < 4 && block_ext[threadIdx.x] != 0u) atomicMax(&P.extrema[threadIdx.x], block_ext[threadIdx.x]);
    if (threadIdx.x == 32 % BLOCK && block_events) atomicAdd(P.n_events, block_events);
    if (tally && P.use_smem) {
        for (int k = threadIdx.x; k < hist_len; k += BLOCK) {
            const unsigned int v = hist[k];
            if (v) atomicAdd(&P.tally[k], (unsigned long long)v);
        }
    }
    if (xh_len && P.hist_smem) {
        for (int k = threadIdx.x; k < xh_len; k += BLOCK) {
            const unsigned int v = xh[k];
            if (v) atomicAdd(&P.hist[k], (unsigned long long)v);
        }
    }
}

cudaError_t launch_finalize(const FinalizeParams &P, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    FinalizeParams Q = P;
    const size_t hist_bytes = (size_t)P.n_rows * (N_COND + (size_t)P.n_theta_bins * (P.n_phi_bins > 1 ? P.n_phi_bins : 1)) * sizeof(unsigned int);
    Q.use_smem = (P.tally != nullptr && hist_bytes <= 96 * 1024) ? 1 : 0;
    const size_t xh_bytes = P.hist ? ((size_t)P.n_scat_bins + (size_t)P.path_bins) * sizeof(unsigned int) : 0;
    Q.hist_smem = (xh_bytes > 0 && xh_bytes <= 64 * 1024) ? 1 : 0;
    const size_t smem = (size_t)P.n_rows * sizeof(float) + (Q.use_smem ? hist_bytes : 0) + (Q.hist_smem ? xh_bytes : 0);
    auto kern = finalize_kernel<BLOCK>;
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

}  // namespace mc3d
