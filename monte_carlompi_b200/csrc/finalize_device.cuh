// finalize_device.cuh -- one photon's raw result -> the reference's output columns + tallies.  Shared by the
// finalize kernel (three-kernel path: reads the 32-byte raw records) and the fused short-walk kernel (the result is
// still in registers).  Reference lines are cited in finalize_kernel.cu.
#pragma once
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

// Block-level accumulation state: the shared-memory tally (when it fits), the optional column histograms behind it,
// and per-thread running sums that finalize_flush() reduces.
struct FinalizeBlock {
    unsigned int *hist;              // [win_rows][stride] in shared memory when P.use_smem: the tally rows [win_begin, +win_rows)
    int win_begin, win_rows;         // use_smem == 1: the whole table; == 2 (sweep launches whose concatenated table does
                                     // not fit): the rows of the case the block is working on, see finalize_window()
    unsigned int *xh;                // [n_scat_bins + path_bins] in shared memory when P.hist_smem
    unsigned int *block_ext;         // [4] shared: extrema of the block (minima complemented)
    unsigned long long *block_events;// shared
    int n_phi, stride, hist_len, xh_len;
    bool tally;
    unsigned long long *case_ev;     // [n_cases] shared (sweep launches that report events per case), or null
    // per thread
    unsigned long long run_events;   // events of the thread's current run of photons of one case (sweep launches)
    uint32_t run_case;
    unsigned long long events;
    uint32_t ns_min, ns_max, pl_min, pl_max;
};

// `hist_base`: shared memory for the tally / column histograms (sized by finalize_smem_words); `statics`: 6 words of
// shared memory for the extrema and the event count (8-byte aligned).  Ends with a __syncthreads().
template <int BLOCK>
__device__ __forceinline__ void finalize_begin(const FinalizeParams &P, FinalizeBlock &B, unsigned int *hist_base,
                                               unsigned int *statics)
{
    B.block_events = reinterpret_cast<unsigned long long *>(statics);
    B.block_ext = statics + 2;
    if (threadIdx.x < 4) B.block_ext[threadIdx.x] = 0u;
    if (threadIdx.x == 0) *B.block_events = 0ull;
    B.n_phi = P.n_phi_bins > 1 ? P.n_phi_bins : 1;
    B.stride = N_COND + P.n_theta_bins * B.n_phi;
    B.tally = P.tally != nullptr;
    B.win_begin = 0;
    B.win_rows = P.use_smem == 2 ? 0 : P.n_rows;
    B.hist_len = P.use_smem == 3 ? P.n_rows * N_COND : (P.use_smem == 2 ? P.win_rows : P.n_rows) * B.stride;
    B.xh_len = P.hist ? P.n_scat_bins + P.path_bins : 0;
    B.case_ev = nullptr;
    if (P.case_events && P.n_cases) {   // sweep launches: hist_base is 8-byte aligned
        B.case_ev = reinterpret_cast<unsigned long long *>(hist_base);
        hist_base += 2 * P.n_cases;
        for (int k = threadIdx.x; k < (int)P.n_cases; k += BLOCK) B.case_ev[k] = 0ull;
    }
    B.run_events = 0ull;
    B.run_case = 0u;
    B.hist = hist_base;
    B.xh = hist_base + (B.tally && P.use_smem ? B.hist_len : 0);
    const int len = (B.tally && P.use_smem ? B.hist_len : 0) + (P.hist_smem ? B.xh_len : 0);
    for (int k = threadIdx.x; k < len; k += BLOCK) hist_base[k] = 0u;
    B.events = 0ull;
    B.ns_min = 0xffffffffu; B.ns_max = 0u; B.pl_min = 0xffffffffu; B.pl_max = 0u;
    __syncthreads();
}

// Photon p ended with direction (ux, uy, uz), path `path_tau` (walk units; inv_ext converts to metres), n_scat
// scatterings, condition `cond`, SSP row `row` of the launch's table (the tally row); the record holds `rec_row`, the
// row within the photon's own case.  SWEEP: `lcase` is the case's index in the launch (events are also summed per case).
// Must be called by whole warps in step when SWEEP (one warp-wide vote).
template <bool SWEEP>
__device__ __forceinline__ void finalize_photon(const FinalizeParams &P, FinalizeBlock &B, uint32_t p, float ux, float uy,
                                                float uz, float path_tau, uint32_t n_scat, uint32_t cond, uint32_t row,
                                                uint32_t rec_row, uint32_t lcase, float inv_ext)
{
    // the angles cost more than everything else here: only computed for a record or a BRF bin
    const bool want_records = P.packed != nullptr || P.theta_n != nullptr || P.phi_n != nullptr;
    const bool binned = B.tally && cond == 1u && P.n_theta_bins > 0;
    float theta = 0.0f, phi = 0.0f;
    if (want_records || binned) theta = atan2f(sqrtf(fmaf(ux, ux, __fmul_rn(uy, uy))), uz);
    if ((want_records || (binned && B.n_phi > 1)) && n_scat != 0u) {
        phi = atan2f(uy, ux);
        if (phi < 0.0f) phi += 6.283185307179586f;
    }
    // fp32 rounding can leave a path of (nearly) zero length slightly negative (Lambertian surface, immediate exit)
    const float path_m = fmaxf(__fmul_rn(path_tau, inv_ext), 0.0f);
    if (P.packed) {
        P.packed[p] = make_uint4((min(n_scat, 0x7fffffu) << 9) | rec_row, __float_as_uint(theta) | (cond << 31),
                                 __float_as_uint(phi) | ((cond >> 1) << 31), __float_as_uint(path_m) | ((cond >> 2) << 31));
    } else {
        if (P.condition) P.condition[p] = (uint8_t)cond;
        if (P.wvl_row) P.wvl_row[p] = (int16_t)rec_row;
        if (P.theta_n) P.theta_n[p] = theta;
        if (P.phi_n) P.phi_n[p] = phi;
        if (P.n_scat) P.n_scat[p] = n_scat;
        if (P.path_length) P.path_length[p] = path_m;
    }
    B.events += (unsigned long long)n_scat + 1ull;
    if (SWEEP && B.case_ev) {   // a thread's photons ascend through the cases: one shared-memory atomic per run
        if (lcase != B.run_case) {
            if (B.run_events) atomicAdd(&B.case_ev[B.run_case], B.run_events);
            B.run_case = lcase;
            B.run_events = 0ull;
        }
        B.run_events += (unsigned long long)n_scat + 1ull;
    }
    B.ns_min = min(B.ns_min, n_scat);
    B.ns_max = max(B.ns_max, n_scat);
    B.pl_min = min(B.pl_min, __float_as_uint(path_m));
    B.pl_max = max(B.pl_max, __float_as_uint(path_m));
    if (B.xh_len) {
        int hb[2] = {-1, -1};
        if (P.n_scat_bins > 0) hb[0] = histogram_bin((double)n_scat, P.n_scat_bins, P.hist_edges);
        if (P.path_bins > 0) {
            hb[1] = histogram_bin(__dmul_rn((double)path_m, P.path_scale), P.path_bins, P.hist_edges + P.n_scat_bins + 1);
            if (hb[1] >= 0) hb[1] += P.n_scat_bins;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (hb[k] < 0) continue;
            if (P.hist_smem) atomicAdd(&B.xh[hb[k]], 1u);
            else atomicAdd(&P.hist[hb[k]], 1ull);
        }
    }
    if (B.tally) {
        const int base = (int)row * B.stride;
        int bin = -1;
        if (cond == 1u && P.n_theta_bins > 0) {
            bin = histogram_bin((double)theta, P.n_theta_bins, P.edges);
            if (bin >= 0 && B.n_phi > 1) {   // np.histogram2d: a sample outside either range is dropped
                const int pb = histogram_bin((double)phi, B.n_phi, P.edges + P.n_theta_bins + 1);
                bin = pb >= 0 ? bin * B.n_phi + pb : -1;
            }
        }
        const uint32_t wrow = row - (uint32_t)B.win_begin;
        if (P.use_smem == 3) {   // outcome counts of every row in shared memory, BRF bins (reflected photons only) global
            atomicAdd(&B.hist[(int)row * N_COND + cond], 1u);
            if (bin >= 0) atomicAdd(&P.tally[base + N_COND + bin], 1ull);
        } else if (P.use_smem && wrow < (uint32_t)B.win_rows) {   // [+ 0] (photons launched in this row) is formed from the condition counts at flush time
            const int wbase = (int)wrow * B.stride;
            atomicAdd(&B.hist[wbase + cond], 1u);
            if (bin >= 0) atomicAdd(&B.hist[wbase + N_COND + bin], 1u);
        } else {
            atomicAdd(&P.tally[base], 1ull);
            atomicAdd(&P.tally[base + cond], 1ull);
            if (bin >= 0) atomicAdd(&P.tally[base + N_COND + bin], 1ull);
        }
    }
}

// The block's shared-memory tally rows -> the global tally: one 64-bit atomic per non-zero bin.  Every thread of the
// block, after a __syncthreads() that ordered the photons' shared-memory atomics before it.
template <int BLOCK>
__device__ __forceinline__ void finalize_flush_tally(const FinalizeParams &P, const FinalizeBlock &B)
{
    if (P.use_smem == 3) {
        for (int k = threadIdx.x; k < P.n_rows * N_COND; k += BLOCK) {
            unsigned int v = B.hist[k];
            if (k % N_COND == 0)
                for (int c = 1; c < N_COND; ++c) v += B.hist[k + c];
            if (v) atomicAdd(&P.tally[(size_t)(k / N_COND) * B.stride + k % N_COND], (unsigned long long)v);
        }
        return;
    }
    const int len = B.win_rows * B.stride;
    unsigned long long *dst = P.tally + (size_t)B.win_begin * B.stride;
    for (int k = threadIdx.x; k < len; k += BLOCK) {
        unsigned int v = B.hist[k];
        if (k % B.stride == 0)
            for (int c = 1; c < N_COND; ++c) v += B.hist[k + c];
        if (v) atomicAdd(&dst[k], (unsigned long long)v);
    }
}

// Sweep launches whose concatenated table is too large for a shared-memory tally (use_smem == 2) keep the rows of ONE
// case there.  Photons are handed out in case order, so the 256 photons a block works on at a time almost always
// belong to the case of the first of them, `base_pid`; the few of a neighbouring case at a boundary go straight to the
// global tally.  Called by every thread of the block with the same (block-uniform) base_pid at the top of each
// iteration of the photon loop; when the block has moved on to a case with other rows it flushes and re-targets.
template <int BLOCK>
__device__ __forceinline__ void finalize_window(const FinalizeParams &P, FinalizeBlock &B, uint32_t base_pid)
{
    if (P.use_smem != 2 || !B.tally) return;
    uint32_t lo = 0u, hi = P.n_cases;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (P.cases[mid].pid_first <= base_pid) lo = mid;
        else hi = mid;
    }
    const int rb = (int)P.cases[lo].row_begin, nr = P.cases[lo].n_rows;
    if (rb == B.win_begin && nr == B.win_rows) return;
    __syncthreads();
    finalize_flush_tally<BLOCK>(P, B);
    __syncthreads();
    B.win_begin = rb;
    B.win_rows = nr <= P.win_rows ? nr : 0;
    for (int k = threadIdx.x; k < B.win_rows * B.stride; k += BLOCK) B.hist[k] = 0u;
    __syncthreads();
}

// Block reduction of the per-thread sums and flush of the shared-memory counts: one 64-bit global atomic per
// non-zero bin per block.  Must be reached by every thread of the block.
template <int BLOCK>
__device__ __forceinline__ void finalize_flush(const FinalizeParams &P, FinalizeBlock &B)
{
    if (B.case_ev && B.run_events) atomicAdd(&B.case_ev[B.run_case], B.run_events);
    unsigned long long events = B.events;
    for (int o = 16; o > 0; o >>= 1) events += __shfl_xor_sync(0xffffffffu, events, o);
    if ((threadIdx.x & 31) == 0 && events) atomicAdd(B.block_events, events);
    if (P.extrema) {   // warp reduce -> one shared-memory atomic per warp -> one global atomic per block
        const uint32_t ns_min = __reduce_min_sync(0xffffffffu, B.ns_min);
        const uint32_t ns_max = __reduce_max_sync(0xffffffffu, B.ns_max);
        const uint32_t pl_min = __reduce_min_sync(0xffffffffu, B.pl_min);
        const uint32_t pl_max = __reduce_max_sync(0xffffffffu, B.pl_max);
        if ((threadIdx.x & 31) == 0 && ns_min <= ns_max) {
            atomicMax(&B.block_ext[0], ~ns_min);   // minima are stored complemented: the buffer starts as zeros
            atomicMax(&B.block_ext[1], ns_max);
            atomicMax(&B.block_ext[2], ~pl_min);
            atomicMax(&B.block_ext[3], pl_max);
        }
    }
    __syncthreads();
    if (P.extrema && threadIdx.x < 4 && B.block_ext[threadIdx.x] != 0u) atomicMax(&P.extrema[threadIdx.x], B.block_ext[threadIdx.x]);
    if (threadIdx.x == 32 % BLOCK && *B.block_events) atomicAdd(P.n_events, *B.block_events);
    if (B.tally && P.use_smem) finalize_flush_tally<BLOCK>(P, B);
    if (B.case_ev) {
        for (int k = threadIdx.x; k < (int)P.n_cases; k += BLOCK) {
            const unsigned long long v = B.case_ev[k];
            if (v) atomicAdd(&P.case_events[P.case0 + k], v);
        }
    }
    if (B.xh_len && P.hist_smem) {
        for (int k = threadIdx.x; k < B.xh_len; k += BLOCK) {
            const unsigned int v = B.xh[k];
            if (v) atomicAdd(&P.hist[k], (unsigned long long)v);
        }
    }
}

// Shared-memory budget of the tally / column histograms: sets use_smem / hist_smem in Q and returns the bytes.
// `windowed`: the kernel's blocks move through the photons in step (finalize_window is usable); otherwise a table that
// does not fit keeps only its outcome counts in shared memory (use_smem == 3).
inline size_t finalize_plan_smem(FinalizeParams &Q, size_t tally_limit, size_t hist_limit, bool windowed = true)
{
    const size_t row_bytes = (N_COND + (size_t)Q.n_theta_bins * (Q.n_phi_bins > 1 ? Q.n_phi_bins : 1)) * sizeof(unsigned int);
    size_t hist_bytes = (size_t)Q.n_rows * row_bytes;
    Q.use_smem = (Q.tally != nullptr && hist_bytes <= tally_limit) ? 1 : 0;
    if (Q.tally != nullptr && !Q.use_smem && windowed && Q.cases && Q.win_rows > 0 && (size_t)Q.win_rows * row_bytes <= tally_limit) {
        Q.use_smem = 2;   // the rows of one case at a time (finalize_window)
        hist_bytes = (size_t)Q.win_rows * row_bytes;
    }
    if (Q.tally != nullptr && !Q.use_smem && !windowed && (size_t)Q.n_rows * N_COND * sizeof(unsigned int) <= tally_limit) {
        Q.use_smem = 3;
        hist_bytes = (size_t)Q.n_rows * N_COND * sizeof(unsigned int);
    }
    const size_t xh_bytes = Q.hist ? ((size_t)Q.n_scat_bins + (size_t)Q.path_bins) * sizeof(unsigned int) : 0;
    Q.hist_smem = (xh_bytes > 0 && xh_bytes <= hist_limit) ? 1 : 0;
    const size_t case_bytes = (Q.case_events && Q.n_cases) ? (size_t)Q.n_cases * sizeof(unsigned long long) : 0;
    return case_bytes + (Q.use_smem ? hist_bytes : 0) + (Q.hist_smem ? xh_bytes : 0);
}

}  // namespace mc3d
