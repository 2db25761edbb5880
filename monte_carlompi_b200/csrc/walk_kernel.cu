// walk_kernel.cu -- production-mode photon random walk (fp32, sm_100a).
//
// Replaces the Python loop reference monte_carloMPI/monte_carlo3D.py:1613-1616 and everything it calls per
// event: populate_pdfs (885-921, 1010-1025), initial_pdfs (1027-1044), Henyey_Greenstein2 (790-800) and the walk
// body monte_carlo3D (1111-1490), plus the per-photon wavelength draw (1515-1520).
//
// Execution model
//   * persistent warps; every lane walks one photon at a time.  The hot loop is one scattering event per
//     iteration: one Philox4x32-10 block -> (HG deflection, azimuth, free path, absorption variate), rotate,
//     move, and ONE rarely-taken branch for "left the top / hit the bottom / maybe absorbed".
//   * a warp claims photon ids 32 at a time with a single atomicAdd (warp-aggregated by construction) and
//     prepares them cooperatively with all 32 lanes active: wavelength draw, SSP row, and the first event
//     (which has no deflection, monte_carlo3D.py:1232-1237).  Survivors go to a per-warp shared-memory ring;
//     a lane whose photon terminated pops its next photon from the ring (a handful of instructions), so the
//     divergent part of a refill is tiny and walk-length divergence is bounded by the refill threshold.
//   * a finished photon leaves one 32-byte raw record (direction, path, n_scat, outcome); angles, records and
//     tallies are produced by the coalesced finalize kernel (finalize_kernel.cu).
//   * per-photon results depend only on (seed, photon id): bit-identical for any grid, block or GPU count.
#include "mc3d_device.cuh"

namespace mc3d {

constexpr int RING = 64;  // entries per warp; a refill adds at most 32 to fewer than 32 leftovers

// Fresh photons prepared by the whole warp, waiting for a lane (first event already taken).
struct WarpRing {
    uint32_t pid[RING];    // photon offset in this launch
    uint32_t row[RING];    // SSP row
    float dtau[RING];      // free path of the first event
};

// Walk state of the photon a lane is carrying.
struct Lane {
    float z, ux, uy, uz;
    float path_lo, path_hi;   // path in optical-depth units: path_hi + path_lo (flushed every 256 events)
    uint32_t i;               // events completed; 0 = the lane carries no photon
    uint32_t pid;             // photon offset in this launch
    uint32_t row;
    uint32_t plo, phi;        // global photon id (Philox counter words 2, 3)
    uint32_t w3;              // absorption word of the last event (for the deferred fine test)
    bool imp;                 // last event's extinction was by the impurity
    // row constants
    float one_m_g, one_m_g2, two_g;
    uint32_t flip, t_hi;
};

// ---- approximate special functions: one MUFU each (the XU pipe), flush-to-zero, independent of nvcc flags ----
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_fast(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_fast(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr float LN2 = 0.6931471805599453f;

// free path -ln(u), u = (w + 0.5) 2^-32   (monte_carlo3D.py:1014, 1036)
__device__ __forceinline__ float free_path(uint32_t w) { return -LN2 * lg2_fast(u32_to_unit(w)); }

// cos, sin of the azimuth 2 pi u, u = ((w >> 8) + 0.5) 2^-24   (monte_carlo3D.py:921, 1258-1259).
// Evaluated at 2 pi u - pi (inside the accurate range of sin/cos.approx) and negated.
__device__ __forceinline__ void azimuth(uint32_t w, float &cp, float &sp)
{
    const float a = fmaf(__uint2float_rn(w >> 8), 3.7450702829239286e-07f, -3.1415922790826485f);
    cp = -cos_fast(a);
    sp = -sin_fast(a);
}

__device__ __forceinline__ void store_raw(const WalkParams &P, uint32_t pid, float ux, float uy, float uz,
                                          float path, uint32_t n_scat, uint32_t cond, uint32_t row)
{
    RawResult *dst = P.raw + pid;
    *reinterpret_cast<float4 *>(dst) = make_float4(ux, uy, uz, path);
    *reinterpret_cast<uint2 *>(&dst->n_scat) = make_uint2(n_scat, cond | (row << 8));
}

// ice or impurity for event i (monte_carlo3D.py:1375-1383); only drawn when an impurity is present
__device__ __forceinline__ bool species_is_impurity(const WalkParams &P, const DevRow &R, uint32_t i, uint32_t plo,
                                                    uint32_t phi)
{
    if (!R.s_any) return false;
    const uint4 v = philox4x32_10(i >> 2, TAG_SPECIES, plo, phi, P.rk);
    const uint32_t sel = i & 3u;
    const uint32_t w = sel == 0 ? v.x : sel == 1 ? v.y : sel == 2 ? v.z : v.w;
    return w <= R.s_last;
}

constexpr uint32_t ALIVE = 0;

// The scattering part of event L.i+1 given its Philox block w: HG deflection, azimuth, rotation, move.
// monte_carlo3D.py:1252-1281 (deflection + rotation), 1352 (move), 1372 (path).  No termination logic.
__device__ __forceinline__ void scatter_and_move(Lane &L, const uint4 w)
{
    // Henyey-Greenstein inverse CDF (790-800) in a cancellation-free form:
    //   D = 1 - g + 2 g r,  s = (1 - g^2)/D,  1 - cos = (1 - g)(1 - r)(s + 1 - g)/D,  sin^2 = (1 - cos)(1 + cos)
    const float r = u32_to_unit(w.x ^ L.flip);
    const float invD = rcp_fast(fmaf(L.two_g, r, L.one_m_g));
    const float s = L.one_m_g2 * invD;
    const float omc = (L.one_m_g * invD) * ((1.0f - r) * (s + L.one_m_g));
    const float ct = 1.0f - omc;
    const float st = sqrt_fast(omc * (2.0f - omc));
    float cp, sp;
    azimuth(w.y, cp, sp);
    const float d2 = fmaf(L.ux, L.ux, L.uy * L.uy);
    float nx, ny, nz;
    if (d2 < 1e-24f) {   // travelling along +-z: the reference's muz_0 == +-1 branches (1262-1269)
        const float sg = L.uz > 0.0f ? 1.0f : -1.0f;
        nx = st * cp; ny = sg * st * sp; nz = sg * ct;
    } else {             // 1270-1281
        const float a = st * rsqrt_fast(d2);
        const float uzc = L.uz * cp;
        nx = fmaf(a, fmaf(L.ux, uzc, -L.uy * sp), L.ux * ct);
        ny = fmaf(a, fmaf(L.uy, uzc, L.ux * sp), L.uy * ct);
        nz = fmaf(-(d2 * a), cp, L.uz * ct);
    }
    L.ux = nx; L.uy = ny; L.uz = nz;
    const float dtau = free_path(w.z);
    L.z = fmaf(dtau, nz, L.z);
    L.path_lo += dtau;
    L.w3 = w.w;
}

// "Something may have happened": the photon left the slab, may have been absorbed (coarse 32-bit test), or is
// due for its periodic renormalisation.  Everything behind this predicate is resolved later, by resolve().
__device__ __forceinline__ bool needs_attention(const WalkParams &P, const Lane &L, uint32_t thi)
{
    return L.z > 0.0f || L.z < P.neg_tau_tot || L.w3 >= thi || (L.i & 255u) == 0u;
}

// Resolve the reference's termination chain monte_carlo3D.py:1390-1466, in its order, for the event L.i that
// just moved the photon to L.z along L.uz.  Executed by all lanes of a warp that need it at once (deferred), so it
// is off the hot path.  On a Lambertian-bottom reflection it also performs the NEXT event (1238-1250: cosine-law
// rejection sampling about +z) so that the hot loop never carries a bottom_reflection flag.
// Returns the condition (0 = keep walking; the lane's state is then ready for the next event).
template <bool IMP>
__device__ __forceinline__ uint32_t resolve(const WalkParams &P, const DevRow *rows, Lane &L)
{
    const DevRow &R = rows[L.row];
    uint32_t cond = ALIVE;
    if (L.z > 0.0f) {   // reflected (1390-1397); z - z_prev = dtau muz, so the overshoot path is z / muz
        L.path_lo -= __fdividef(L.z, L.uz);
        cond = 1u;
    } else if (L.z < P.neg_tau_tot) {   // 1399-1459
        L.path_lo -= __fdividef(L.z + P.tau_tot, L.uz);
        L.z = P.neg_tau_tot;
        cond = (L.i == 1u) ? 3u : 2u;
        if (P.lambert_bottom) {
            const uint4 b = philox4x32_10(L.i, TAG_LAMBERT, L.plo, L.phi, P.rk);
            if ((long long)b.x <= P.refl_thr) {
                // ---- reflected by the Lambertian bottom: event i+1 happens here ----
                L.i += 1u;
                const uint4 w = philox4x32_10(L.i, TAG_EVENT, L.plo, L.phi, P.rk);
                float ct, st;
                for (uint32_t j = 0;; ++j) {
                    const uint4 a = philox4x32_10(L.i, TAG_LAMBERT | ((1u + (j >> 1)) << 8), L.plo, L.phi, P.rk);
                    const float u_t = u32_to_unit((j & 1u) ? a.z : a.x);
                    const float r1 = u32_to_unit((j & 1u) ? a.w : a.y);
                    float s_, c_;
                    sincosf(1.5707963267948966f * u_t, &s_, &c_);
                    if (r1 < 2.0f * s_ * c_) { ct = c_; st = s_; break; }
                }
                float cp, sp;
                azimuth(w.y, cp, sp);
                L.ux = st * cp; L.uy = st * sp; L.uz = ct;   // muz_0 == 1 branch, 1262-1265
                const float dt2 = free_path(w.z);
                L.z = fmaf(dt2, ct, P.neg_tau_tot);
                L.path_lo += dt2;
                L.w3 = w.w;
                L.imp = IMP ? species_is_impurity(P, R, L.i, L.plo, L.phi) : false;
                cond = ALIVE;
                if (L.z > 0.0f) {
                    L.path_lo -= __fdividef(L.z, L.uz);
                    cond = 1u;
                }
            }
        }
    }
    if (cond == ALIVE) {   // 1461-1466: absorbed iff the 40-bit variate (w3 << 8 | low byte of w1) >= T40
        const uint32_t t_hi = L.imp ? R.ti_hi : R.t_hi, t_lo = L.imp ? R.ti_lo : R.t_lo;
        bool absorbed = L.w3 > t_hi;
        if (L.w3 == t_hi) {   // probability 2^-32: regenerate the event's block for the low byte
            const uint4 w = philox4x32_10(L.i, TAG_EVENT, L.plo, L.phi, P.rk);
            absorbed = (w.y & 0xffu) >= t_lo;
        }
        if (absorbed) cond = L.imp ? 5u : 4u;
    }
    if (cond == ALIVE && (L.i & 255u) == 0u) {   // keyed on the photon's own event count: scheduling independent
        const float rn = rsqrt_fast(fmaf(L.ux, L.ux, fmaf(L.uy, L.uy, L.uz * L.uz)));
        L.ux *= rn; L.uy *= rn; L.uz *= rn;
        L.path_hi += L.path_lo;
        L.path_lo = 0.0f;
    }
    return cond;
}

__device__ __forceinline__ void load_row_constants(Lane &L, const DevRow &R)
{
    L.one_m_g = R.one_m_g; L.one_m_g2 = R.one_m_g2; L.two_g = R.two_g; L.flip = R.flip; L.t_hi = R.t_hi;
}

// Claim 32 photon ids, draw their wavelengths and take the first step (which has no deflection); append the
// survivors to the warp's ring.  All 32 lanes execute this.  Returns 0 when the photon range is exhausted.
template <bool IMP>
__device__ __forceinline__ uint32_t prepare_batch(const WalkParams &P, const DevRow *rows, WarpRing &Q,
                                                  uint32_t &ring_tail, uint32_t lane)
{
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(P.counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= P.n_photon) return 0u;
    const uint32_t pid = base + lane;
    bool survive = false;
    float dtau = 0.0f;
    uint32_t row = 0;
    if (pid < P.n_photon) {
        const unsigned long long gid = P.photon_begin + pid;
        const uint32_t plo = (uint32_t)gid, phi = (uint32_t)(gid >> 32);
        // wavelength: np.around(np.random.normal(wvl0, scale), 2), monte_carlo3D.py:1515-1520 (Box-Muller)
        const uint4 wv = philox4x32_10(0u, TAG_WAVELENGTH, plo, phi, P.rk);
        const float zn = sqrtf(-2.0f * logf(u32_to_unit(wv.x))) * cospif(2.0f * u32_to_unit(wv.y));
        const int r = (int)rint(P.wvl0_x100 + P.sigma_x100 * (double)zn) - P.k_first;
        row = (uint32_t)max(0, min(P.n_rows - 1, r));
        const DevRow &R = rows[row];
        // first event: draws of initial_pdfs (1035-1038), no deflection (1232-1237)
        const uint4 w = philox4x32_10(1u, TAG_EVENT, plo, phi, P.rk);
        dtau = free_path(w.z);
        const float z1 = dtau * P.mu0z;
        const bool imp = IMP ? species_is_impurity(P, R, 1u, plo, phi) : false;
        survive = true;
        if (z1 < P.neg_tau_tot || w.w >= (imp ? R.ti_hi : R.t_hi)) {
            Lane L;
            L.z = z1; L.ux = P.mu0x; L.uy = 0.0f; L.uz = P.mu0z; L.i = 1u; L.path_lo = dtau; L.path_hi = 0.0f;
            L.pid = pid; L.row = row; L.plo = plo; L.phi = phi; L.w3 = w.w; L.imp = imp;
            load_row_constants(L, R);
            uint32_t cond = resolve<IMP>(P, rows, L);
            if (cond == ALIVE && L.i != 1u) {
                // reflected off the Lambertian bottom on its first step and still alive after event 2: it no
                // longer fits the ring's "fresh photon" format, so it is walked to completion here (thin slabs only)
                do {
                    L.i += 1u;
                    scatter_and_move(L, philox4x32_10(L.i, TAG_EVENT, plo, phi, P.rk));
                    L.imp = IMP ? species_is_impurity(P, R, L.i, plo, phi) : false;
                    if (needs_attention(P, L, L.imp ? R.ti_hi : L.t_hi)) cond = resolve<IMP>(P, rows, L);
                } while (cond == ALIVE);
            }
            if (cond != ALIVE) {
                store_raw(P, pid, L.ux, L.uy, L.uz, L.path_hi + L.path_lo, L.i - 1u, cond, row);
                survive = false;
            }
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, survive);
    if (survive) {
        const uint32_t slot = (ring_tail + __popc(m & ((1u << lane) - 1u))) & (RING - 1);
        Q.pid[slot] = pid;
        Q.row[slot] = row;
        Q.dtau[slot] = dtau;
    }
    ring_tail += __popc(m);
    __syncwarp();
    return 32u;
}

template <bool IMP, int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) walk_kernel(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    WarpRing *rings = reinterpret_cast<WarpRing *>(smem_raw + ((P.n_rows * sizeof(DevRow) + 15) & ~size_t(15)));
    for (int k = threadIdx.x; k < P.n_rows * (int)(sizeof(DevRow) / 4); k += BLOCK)
        reinterpret_cast<uint32_t *>(rows)[k] = reinterpret_cast<const uint32_t *>(P.rows)[k];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    WarpRing &Q = rings[threadIdx.x >> 5];
    uint32_t ring_head = 0, ring_tail = 0;   // warp-uniform
    bool exhausted = false;                  // warp-uniform
    uint32_t threshold = max(1u, min(32u, P.refill_threshold));

    Lane L;
    L.z = 0.f; L.ux = 0.f; L.uy = 0.f; L.uz = -1.f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 0; L.pid = 0; L.row = 0; L.plo = 0; L.phi = 0; L.w3 = 0; L.imp = false;
    L.one_m_g = 1.f; L.one_m_g2 = 1.f; L.two_g = 0.f; L.flip = 0; L.t_hi = 0;
    bool alive = false;                  // false: the lane is waiting (needs attention, or carries no photon)

    for (;;) {
        if (!__all_sync(0xffffffffu, alive)) {
            // ------------------------------------------------------------ deferred attention + refill
            const uint32_t waiting = __ballot_sync(0xffffffffu, !alive);
            if (__popc(waiting) >= threshold) {
                // 1. lanes whose last event tripped needs_attention(): finish or resume them, all together
                if (!alive && L.i != 0u) {
                    const uint32_t cond = resolve<IMP>(P, rows, L);
                    if (cond != ALIVE) {
                        store_raw(P, L.pid, L.ux, L.uy, L.uz, L.path_hi + L.path_lo, L.i - 1u, cond, L.row);
                        L.i = 0u;
                    } else {
                        alive = true;
                    }
                }
                // 2. lanes without a photon take a fresh one from the ring
                const uint32_t empty = __ballot_sync(0xffffffffu, !alive);
                const uint32_t need = __popc(empty);
                if (need != 0u && !(exhausted && ring_head == ring_tail)) {
                    while (!exhausted && (ring_tail - ring_head) < need)
                        if (prepare_batch<IMP>(P, rows, Q, ring_tail, lane) == 0u) exhausted = true;
                    const uint32_t avail = ring_tail - ring_head;
                    if (!alive) {
                        const uint32_t rank = __popc(empty & ((1u << lane) - 1u));
                        if (rank < avail) {
                            const uint32_t slot = (ring_head + rank) & (RING - 1);
                            L.pid = Q.pid[slot];
                            L.row = Q.row[slot];
                            const float dtau = Q.dtau[slot];
                            load_row_constants(L, rows[L.row]);
                            L.ux = P.mu0x; L.uy = 0.0f; L.uz = P.mu0z;
                            L.z = dtau * P.mu0z;
                            L.path_lo = dtau;
                            L.path_hi = 0.0f;
                            L.i = 1u;
                            const unsigned long long gid = P.photon_begin + L.pid;
                            L.plo = (uint32_t)gid; L.phi = (uint32_t)(gid >> 32);
                            alive = true;
                        }
                    }
                    ring_head += min(avail, need);
                    __syncwarp();
                }
                if (exhausted && ring_head == ring_tail) {
                    // nothing left to hand out: from now on every waiting lane is resolved immediately
                    threshold = 1u;
                    if (__ballot_sync(0xffffffffu, alive) == 0u) break;
                }
            }
        }
        // ---------------------------------------------------------------- one scattering event per live lane
        if (alive) {
            const uint4 w = philox4x32_10(L.i + 1u, TAG_EVENT, L.plo, L.phi, P.rk);
            L.i += 1u;
            scatter_and_move(L, w);
            uint32_t thi = L.t_hi;
            if (IMP) {
                const DevRow &R = rows[L.row];
                L.imp = species_is_impurity(P, R, L.i, L.plo, L.phi);
                thi = L.imp ? R.ti_hi : thi;
            }
            alive = !needs_attention(P, L, thi);
        }
    }
}

// ---- launch helper (called from the host runtime) ------------------------------------------------------------

size_t walk_smem_bytes(int n_rows, int block_threads)
{
    return ((n_rows * sizeof(DevRow) + 15) & ~size_t(15)) + (block_threads / 32) * sizeof(WarpRing);
}

template <bool IMP, int BLOCK, int MIN_BLOCKS>
static cudaError_t launch_one(const WalkParams &P, int grid, cudaStream_t stream, int *occupancy)
{
    const size_t smem = walk_smem_bytes(P.n_rows, BLOCK);
    auto kern = walk_kernel<IMP, BLOCK, MIN_BLOCKS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occupancy) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, kern, BLOCK, smem);
    kern<<<grid, BLOCK, smem, stream>>>(P);
    return cudaGetLastError();
}

template <int BLOCK, int MIN_BLOCKS>
static cudaError_t launch_variant(const WalkParams &P, bool impurity, int grid, cudaStream_t stream, int *occupancy)
{
    return impurity ? launch_one<true, BLOCK, MIN_BLOCKS>(P, grid, stream, occupancy)
                    : launch_one<false, BLOCK, MIN_BLOCKS>(P, grid, stream, occupancy);
}

// block_threads in {128, 256, 512}; blocks_per_sm is the occupancy target the variant is compiled for
// (64 / 48 / 40 registers per thread for <= 32 / 40 / 48 resident warps per SM).
// With `occupancy` non-null nothing is launched; the resident blocks per SM are returned through it.
cudaError_t launch_walk(const WalkParams &P, bool impurity, int block_threads, int blocks_per_sm, int grid,
                        cudaStream_t stream, int *occupancy)
{
    const int warps_per_sm = block_threads / 32 * blocks_per_sm;
    if (block_threads == 128) {
        if (warps_per_sm <= 32) return launch_variant<128, 8>(P, impurity, grid, stream, occupancy);
        if (warps_per_sm <= 40) return launch_variant<128, 10>(P, impurity, grid, stream, occupancy);
        return launch_variant<128, 12>(P, impurity, grid, stream, occupancy);
    }
    if (block_threads == 256) {
        if (warps_per_sm <= 32) return launch_variant<256, 4>(P, impurity, grid, stream, occupancy);
        if (warps_per_sm <= 40) return launch_variant<256, 5>(P, impurity, grid, stream, occupancy);
        return launch_variant<256, 6>(P, impurity, grid, stream, occupancy);
    }
    if (block_threads == 512) {
        if (warps_per_sm <= 32) return launch_variant<512, 2>(P, impurity, grid, stream, occupancy);
        return launch_variant<512, 3>(P, impurity, grid, stream, occupancy);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mc3d
