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

struct WarpRing {
    uint32_t pid[RING];   // photon offset in this launch
    uint32_t row[RING];   // SSP row
    float dtau[RING];     // free path of the first event
};

// Walk state of the photon a lane is carrying.
struct Lane {
    float z, ux, uy, uz;
    float path_lo, path_hi;   // path in optical-depth units: path_hi + path_lo (flushed every 256 events)
    uint32_t i;               // events completed
    uint32_t pid;             // photon offset in this launch
    uint32_t row;
    uint32_t plo, phi;        // global photon id (Philox counter words 2, 3)
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

// 40-bit absorption test u >= ssa (monte_carlo3D.py:1461), w3 = high 32 bits, low byte of w1 = low 8 bits
__device__ __forceinline__ bool absorbed40(uint32_t w3, uint32_t w1, uint32_t t_hi, uint32_t t_lo)
{
    return w3 > t_hi || (w3 == t_hi && (w1 & 0xffu) >= t_lo);
}

constexpr uint32_t ALIVE = 0;

// State handed to / returned from the rarely executed boundary code (kept out of the hot loop's registers).
struct SlowIO {
    float z, ux, uy, uz, path_add;
    uint32_t i;
};

// Event io.i just moved the photon from z_prev to io.z by dtau and one of "z > 0", "z < -tau_tot", "absorption
// word at/above the coarse threshold" holds.  Resolves the reference's termination chain
// monte_carlo3D.py:1390-1466 in its order; on a Lambertian-bottom reflection it also performs the NEXT event
// (1238-1250: cosine-law rejection sampling about +z) so the hot loop never carries a bottom_reflection flag.
// Returns the condition (0 = keep walking).
template <bool IMP>
__device__ __forceinline__ uint32_t slow_path(const WalkParams &P, const DevRow *rows, uint32_t row, SlowIO *io,
                                           float z_prev, float dtau, uint32_t w1, uint32_t w3, bool imp,
                                           uint32_t plo, uint32_t phi)
{
    const DevRow &R = rows[row];
    float z = io->z;
    io->path_add = 0.0f;
    if (z > 0.0f) {   // reflected (1390-1397): remove the part of the step above the surface
        io->path_add = -(z * dtau) / (z - z_prev);
        return 1u;
    }
    if (z < P.neg_tau_tot) {   // 1399-1459
        io->path_add = -((z + P.tau_tot) * dtau) / (z - z_prev);
        io->z = z = P.neg_tau_tot;
        const uint32_t exit_cond = (io->i == 1u) ? 3u : 2u;
        if (!P.lambert_bottom) return exit_cond;
        const uint4 b = philox4x32_10(io->i, TAG_LAMBERT, plo, phi, P.rk);
        if ((long long)b.x > P.refl_thr) return exit_cond;
        // ---- reflected by the Lambertian bottom: event i+1 happens here ----
        const uint32_t i2 = io->i + 1u;
        io->i = i2;
        const uint4 w = philox4x32_10(i2, TAG_EVENT, plo, phi, P.rk);
        float ct, st;
        for (uint32_t j = 0;; ++j) {
            const uint4 a = philox4x32_10(i2, TAG_LAMBERT | ((1u + (j >> 1)) << 8), plo, phi, P.rk);
            const float u_t = u32_to_unit((j & 1u) ? a.z : a.x);
            const float r1 = u32_to_unit((j & 1u) ? a.w : a.y);
            float s_, c_;
            sincosf(1.5707963267948966f * u_t, &s_, &c_);
            if (r1 < 2.0f * s_ * c_) { ct = c_; st = s_; break; }
        }
        float cp, sp;
        azimuth(w.y, cp, sp);
        io->ux = st * cp; io->uy = st * sp; io->uz = ct;   // muz_0 == 1 branch, 1262-1265
        const float dt2 = free_path(w.z);
        z = fmaf(dt2, ct, P.neg_tau_tot);
        io->z = z;
        io->path_add += dt2;
        const bool imp2 = IMP ? species_is_impurity(P, R, i2, plo, phi) : false;
        if (z > 0.0f) {
            io->path_add += -(z * dt2) / (z - P.neg_tau_tot);
            return 1u;
        }
        if (absorbed40(w.w, w.y, imp2 ? R.ti_hi : R.t_hi, imp2 ? R.ti_lo : R.t_lo)) return imp2 ? 5u : 4u;
        return ALIVE;
    }
    if (absorbed40(w3, w1, imp ? R.ti_hi : R.t_hi, imp ? R.ti_lo : R.t_lo)) return imp ? 5u : 4u;   // 1461-1466
    return ALIVE;
}

// One scattering event (i >= 2) of the photon in L: monte_carlo3D.py:1212-1466 for the sphere/HG branch.
// Returns the condition (0 = still walking).
template <bool IMP>
__device__ __forceinline__ uint32_t event_step(const WalkParams &P, const DevRow *rows, Lane &L)
{
    L.i += 1u;
    const uint4 w = philox4x32_10(L.i, TAG_EVENT, L.plo, L.phi, P.rk);
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
    // rotate the direction cosines (1262-1281)
    const float d2 = fmaf(L.ux, L.ux, L.uy * L.uy);
    float nx, ny, nz;
    if (d2 < 1e-24f) {   // travelling along +-z: the reference's muz_0 == +-1 branches
        const float sg = L.uz > 0.0f ? 1.0f : -1.0f;
        nx = st * cp; ny = sg * st * sp; nz = sg * ct;
    } else {
        const float inv_d = rsqrt_fast(d2), a = st * inv_d;
        const float uzc = L.uz * cp;
        nx = fmaf(a, fmaf(L.ux, uzc, -L.uy * sp), L.ux * ct);
        ny = fmaf(a, fmaf(L.uy, uzc, L.ux * sp), L.uy * ct);
        nz = fmaf(-(d2 * a), cp, L.uz * ct);
    }
    L.ux = nx; L.uy = ny; L.uz = nz;
    // free path (1014), move (1352), path length (1372)
    const float dtau = free_path(w.z);
    const float z_prev = L.z;
    L.z = fmaf(dtau, nz, z_prev);
    L.path_lo += dtau;
    bool imp = false;
    uint32_t thi = L.t_hi;
    if (IMP) {
        const DevRow &R = rows[L.row];
        imp = species_is_impurity(P, R, L.i, L.plo, L.phi);
        thi = imp ? R.ti_hi : thi;
    }
    uint32_t cond = ALIVE;
    if (L.z > 0.0f || L.z < P.neg_tau_tot || w.w >= thi) {
        SlowIO io;
        io.z = L.z; io.ux = L.ux; io.uy = L.uy; io.uz = L.uz; io.i = L.i;
        cond = slow_path<IMP>(P, rows, L.row, &io, z_prev, dtau, w.y, w.w, imp, L.plo, L.phi);
        L.z = io.z; L.ux = io.ux; L.uy = io.uy; L.uz = io.uz; L.i = io.i;
        L.path_lo += io.path_add;
    }
    if ((L.i & 255u) == 0u) {   // keyed on the photon's own event count -> independent of scheduling
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
            SlowIO io;
            io.z = z1; io.ux = P.mu0x; io.uy = 0.0f; io.uz = P.mu0z; io.i = 1u;
            uint32_t cond = slow_path<IMP>(P, rows, row, &io, 0.0f, dtau, w.y, w.w, imp, plo, phi);
            if (cond == ALIVE && io.i != 1u) {
                // reflected off the Lambertian bottom on its first step and still alive after event 2: it no
                // longer fits the ring's "fresh photon" format, so it is walked to completion here (thin slabs only)
                Lane L;
                L.z = io.z; L.ux = io.ux; L.uy = io.uy; L.uz = io.uz; L.i = io.i;
                L.path_lo = dtau + io.path_add; L.path_hi = 0.0f;
                L.pid = pid; L.row = row; L.plo = plo; L.phi = phi;
                load_row_constants(L, R);
                do { cond = event_step<IMP>(P, rows, L); } while (cond == ALIVE);
                store_raw(P, pid, L.ux, L.uy, L.uz, L.path_hi + L.path_lo, L.i - 1u, cond, row);
                survive = false;
            } else if (cond != ALIVE) {
                store_raw(P, pid, io.ux, io.uy, io.uz, dtau + io.path_add, io.i - 1u, cond, row);
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
    L.i = 0; L.pid = 0; L.row = 0; L.plo = 0; L.phi = 0;
    L.one_m_g = 1.f; L.one_m_g2 = 1.f; L.two_g = 0.f; L.flip = 0; L.t_hi = 0;
    bool alive = false;

    for (;;) {
        // ---------------------------------------------------------------- refill (warp-uniform branch)
        const uint32_t dead = __ballot_sync(0xffffffffu, !alive);
        if (__popc(dead) >= threshold) {
            if (exhausted && ring_head == ring_tail) {
                if (dead == 0xffffffffu) break;
            } else {
                const uint32_t need = __popc(dead);
                while (!exhausted && (ring_tail - ring_head) < need)
                    if (prepare_batch<IMP>(P, rows, Q, ring_tail, lane) == 0u) exhausted = true;
                const uint32_t avail = ring_tail - ring_head;
                if (!alive) {
                    const uint32_t rank = __popc(dead & ((1u << lane) - 1u));
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
                if (exhausted && ring_head == ring_tail) threshold = 32u;   // only "all lanes done" matters now
                continue;
            }
        }
        // ---------------------------------------------------------------- one scattering event per live lane
        if (alive) {
            const uint32_t cond = event_step<IMP>(P, rows, L);
            if (cond != ALIVE) {
                store_raw(P, L.pid, L.ux, L.uy, L.uz, L.path_hi + L.path_lo, L.i - 1u, cond, L.row);
                alive = false;
            }
        }
    }
}

// ---- launch helper (called from the host runtime) ------------------------------------------------------------

size_t walk_smem_bytes(int n_rows, int block_threads)
{
    return ((n_rows * sizeof(DevRow) + 15) & ~size_t(15)) + (block_threads / 32) * sizeof(WarpRing);
}

template <bool IMP, int BLOCK, int MIN_BLOCKS>
static cudaError_t launch_one(const WalkParams &P, int grid, cudaStream_t stream)
{
    const size_t smem = walk_smem_bytes(P.n_rows, BLOCK);
    auto kern = walk_kernel<IMP, BLOCK, MIN_BLOCKS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, BLOCK, smem, stream>>>(P);
    return cudaGetLastError();
}

// block_threads in {128, 256, 512}; blocks_per_sm is the occupancy target the variant was compiled for
cudaError_t launch_walk(const WalkParams &P, bool impurity, int block_threads, int blocks_per_sm, int grid,
                        cudaStream_t stream)
{
    const int warps_per_sm = block_threads / 32 * blocks_per_sm;
#define MC3D_PICK(B, M)                                                    \
    return impurity ? launch_one<true, B, M>(P, grid, stream) : launch_one<false, B, M>(P, grid, stream)
    if (block_threads == 128) {
        if (warps_per_sm <= 32) { MC3D_PICK(128, 8); } else if (warps_per_sm <= 40) { MC3D_PICK(128, 10); } else { MC3D_PICK(128, 12); }
    } else if (block_threads == 256) {
        if (warps_per_sm <= 32) { MC3D_PICK(256, 4); } else if (warps_per_sm <= 40) { MC3D_PICK(256, 5); } else { MC3D_PICK(256, 6); }
    } else if (block_threads == 512) {
        if (warps_per_sm <= 32) { MC3D_PICK(512, 2); } else { MC3D_PICK(512, 3); }
    }
#undef MC3D_PICK
    return cudaErrorInvalidValue;
}

int walk_occupancy(bool impurity, int block_threads, int blocks_per_sm, int n_rows)
{
    int nb = 0;
    const size_t smem = walk_smem_bytes(n_rows, block_threads);
#define MC3D_OCC(B, M)                                                                                            \
    do {                                                                                                            \
        if (impurity) {                                                                                             \
            cudaFuncSetAttribute(walk_kernel<true, B, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, walk_kernel<true, B, M>, B, smem);                   \
        } else {                                                                                                    \
            cudaFuncSetAttribute(walk_kernel<false, B, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, walk_kernel<false, B, M>, B, smem);                  \
        }                                                                                                           \
    } while (0)
    const int warps_per_sm = block_threads / 32 * blocks_per_sm;
    if (block_threads == 128) {
        if (warps_per_sm <= 32) MC3D_OCC(128, 8); else if (warps_per_sm <= 40) MC3D_OCC(128, 10); else MC3D_OCC(128, 12);
    } else if (block_threads == 256) {
        if (warps_per_sm <= 32) MC3D_OCC(256, 4); else if (warps_per_sm <= 40) MC3D_OCC(256, 5); else MC3D_OCC(256, 6);
    } else if (block_threads == 512) {
        if (warps_per_sm <= 32) MC3D_OCC(512, 2); else MC3D_OCC(512, 3);
    }
#undef MC3D_OCC
    return nb;
}

}  // namespace mc3d
