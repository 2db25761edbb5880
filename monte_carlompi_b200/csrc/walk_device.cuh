// walk_device.cuh -- device code shared by the init kernel (first event of every photon) and the walk kernel
// (all later events): one scattering event, the attention predicate and the deferred resolution of the
// reference's termination chain.  Reference lines are cited at each function.
#pragma once
#include "mc3d_device.cuh"

namespace mc3d {


// Walk state of the photon a lane is carrying.
struct Lane {
    float z, ux, uy, uz;
    float path_lo, path_hi;   // path in optical-depth units: path_hi + path_lo (flushed at every renormalisation)
    uint32_t i;               // events completed; 0 = the lane carries no photon
    uint32_t blk;             // next block of the photon's walk stream (3 x the number of groups started)
    uint32_t plo;             // low word of the global photon id (Philox counter word 2).  Single-case launches: the
                              // high word is a launch constant (the host never lets a launch cross 2^32)
    uint32_t phi;             // sweep launches only: high word of the photon id = case index << 8 | bits 32..39
    uint32_t row_addr;        // shared-space address of rows[row] (the hot loop loads the row constants through it)
    uint32_t key;             // key word of the last event: key16 (coarse absorption variate) in its top half
    PhiloxWalkConst pk;       // photon-constant part of a walk block's first two Philox rounds
    bool imp;                 // last event's extinction was by the impurity
};

__device__ __forceinline__ uint32_t lane_row(const Lane &L, uint32_t rows_addr) { return (L.row_addr - rows_addr) / (uint32_t)sizeof(DevRow); }

// The case a lane's photon belongs to, the high word of its id, its index in the launch.  SWEEP = false: launch constants.
template <bool SWEEP>
__device__ __forceinline__ uint32_t lane_lcase(const WalkParams &P, const Lane &L) { return SWEEP ? (L.phi >> 8) - P.case0 : 0u; }
template <bool SWEEP>
__device__ __forceinline__ const DevCase &lane_case(const WalkParams &P, const DevCase *cases, const Lane &L)
{
    return SWEEP ? cases[lane_lcase<SWEEP>(P, L)] : P.c;
}
template <bool SWEEP>
__device__ __forceinline__ uint32_t lane_phi(const WalkParams &P, const Lane &L) { return SWEEP ? L.phi : (uint32_t)(P.c.id0 >> 32); }
__device__ __forceinline__ uint32_t lane_pid(const DevCase &C, const Lane &L) { return L.plo - (uint32_t)C.id0; }

// The part of a DevRow the hot loop needs: one 16-byte and one 4-byte shared-memory load per event (the loads are
// issued before the Philox rounds and are off the critical path; keeping them out of registers buys occupancy).
struct HotRow {
    float one_m_g, one_m_g2, d_scale, d_off, omr_scale, omr_off;
    uint32_t t_hot, ti_hot;
    float neg_tau;   // bottom of the slab: a launch constant, or (sweep launches) read with the row of the photon's case
};
__device__ __forceinline__ uint32_t shared_address(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool SWEEP>
__device__ __forceinline__ HotRow load_hot_row(const WalkParams &P, uint32_t row_addr)
{
    HotRow h;
    if (SWEEP) asm("ld.shared.f32 %0, [%1+60];" : "=f"(h.neg_tau) : "r"(row_addr));
    else h.neg_tau = P.c.neg_tau_tot;
    asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
        : "=f"(h.one_m_g), "=f"(h.one_m_g2), "=f"(h.d_scale), "=f"(h.d_off) : "r"(row_addr));
    asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+16];"
        : "=r"(h.t_hot), "=f"(h.omr_scale), "=f"(h.omr_off), "=r"(h.ti_hot) : "r"(row_addr));
    return h;
}

// ---- approximate special functions: one MUFU each (the XU pipe), flush-to-zero, independent of nvcc flags ----
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_fast(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_fast(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr float LN2 = 0.6931471805599453f;

// free path -ln(u) / ln 2 = -log2(u), u = (w + 0.5) 2^-32   (monte_carlo3D.py:1014, 1036), in the walk's depth unit
__device__ __forceinline__ float free_path(uint32_t w) { return -lg2_fast(u32_to_unit(w)); }

// cos, sin of the azimuth 2 pi u, u = (w + 0.5) 2^-32   (monte_carlo3D.py:921, 1258-1259).
// Evaluated at 2 pi u - pi (inside the accurate range of sin/cos.approx) and negated.
__device__ __forceinline__ void azimuth(uint32_t w, float &cp, float &sp)
{
    const float a = fmaf(__uint2float_rn(w), 1.4629180792671596e-09f, -3.1415926528583587f);
    cp = -cos_fast(a);
    sp = -sin_fast(a);
}

__device__ __forceinline__ void store_raw(const WalkParams &P, uint32_t pid, float ux, float uy, float uz,
                                          float path, uint32_t n_scat, uint32_t cond, uint32_t row, uint32_t lcase)
{
    RawResult *dst = P.raw + pid;
    // the whole 32-byte sector is written (two 16-byte stores): a partly written sector makes L2 fetch the rest from
    // DRAM first (it showed as 16 B of reads per record in the ncu captures)
    *reinterpret_cast<float4 *>(dst) = make_float4(ux, uy, uz, path);
    *reinterpret_cast<uint4 *>(&dst->n_scat) = make_uint4(n_scat, cond | (row << 8), lcase, 0u);
}

// ice or impurity for event i (monte_carlo3D.py:1375-1383); only drawn when an impurity is present
__device__ __forceinline__ bool species_is_impurity(const WalkParams &P, const DevRow &R, uint32_t i, uint32_t plo,
                                                    uint32_t phi)
{
    if (!R.s_any) return false;
    const uint4 v = philox4x32(i >> 2, TAG_SPECIES, plo, phi, P.rk);
    const uint32_t sel = i & 3u;
    const uint32_t w = sel == 0 ? v.x : sel == 1 ? v.y : sel == 2 ? v.z : v.w;
    return w <= R.s_last;
}

constexpr uint32_t ALIVE = 0;

// The scattering part of an event, monte_carlo3D.py:1252-1281 (deflection + rotation), 1352 (move), 1372 (path), in
// two steps.  prepare_event() is everything that depends only on the event's three random words and the photon's
// SSP row: HG deflection, azimuth, free path, absorption key.  apply_event() rotates the lane's direction and moves
// the photon.  The hot loop runs them back to back; the latency-bound tail of a launch prepares the four events of
// a group together (independent instruction streams) and then applies them in order, which shortens the dependent
// chain per event to the rotation itself.  Every operation is spelled with explicit fused / unfused intrinsics, so
// both arrangements (and every kernel that inlines them) compute bit-identical values.  No termination logic here.
struct Prepared {
    float ct, st2;     // cos(theta), sin^2(theta) of the Henyey-Greenstein deflection
    float cp, sp;      // cos, sin of the azimuth
    float dtau;        // free path in the walk's depth unit
    uint32_t key;      // key16 (coarse absorption variate) in the top half
};

__device__ __forceinline__ Prepared prepare_event(const HotRow &H, uint32_t w_hg, uint32_t w_az, uint32_t w_fp)
{
    // Henyey-Greenstein inverse CDF (790-800) in a cancellation-free form:
    //   D = 1 - g + 2 g r,  s = (1 - g^2)/D,  1 - cos = (1 - g)(1 - r)(s + 1 - g)/D,  sin^2 = (1 - cos)(1 + cos)
    Prepared e;
    const float wf = __uint2float_rn(w_hg);
    const float invD = rcp_fast(fmaf(wf, H.d_scale, H.d_off));          // D = 1 - g + 2 g r
    const float omr = fmaf(wf, H.omr_scale, H.omr_off);                 // 1 - r  (r itself for a g == 0 row)
    const float s1 = fmaf(H.one_m_g2, invD, H.one_m_g);                 // s + 1 - g
    const float omc = __fmul_rn(__fmul_rn(H.one_m_g, invD), __fmul_rn(omr, s1));
    e.ct = __fsub_rn(1.0f, omc);
    e.st2 = __fmul_rn(omc, __fsub_rn(2.0f, omc));                       // sin^2
    azimuth(w_az, e.cp, e.sp);
    e.dtau = free_path(w_fp);
    // key16 = (low byte of the HG word) << 8 | low byte of the free-path word, in the top half of the key word
    e.key = __byte_perm(w_hg, w_fp, 0x0451);
    return e;
}

__device__ __forceinline__ void apply_event(Lane &L, const Prepared &e)
{
    // rotate the direction cosines, monte_carlo3D.py:1270-1281, with sqrt(1 - muz^2) taken as sqrt(mux^2 + muy^2).
    // The reference's muz_0 == +-1 branches (1262-1269) need no code here: vertical incidence enters the walk as
    // (-1e-15, 0, -1), for which this formula reproduces the muz_0 == -1 branch exactly (the host sets mu0x), and
    // the step after a Lambertian reflection is taken in resolve().  sin(theta) / d and sin(theta) d come from ONE
    // rsqrt: with q = sin^2 d^2, sin / d = sin^2 rsqrt(q) and sin d = q rsqrt(q).  The clamp only keeps an exactly
    // forward scattering (sin^2 == 0) or a (never observed) exactly vertical direction finite.
    const float d2 = fmaf(L.ux, L.ux, __fmul_rn(L.uy, L.uy));
    const float q = fmaxf(__fmul_rn(e.st2, d2), 1e-36f);
    const float rq = rsqrt_fast(q);
    const float a = __fmul_rn(e.st2, rq);
    const float uzc = __fmul_rn(L.uz, e.cp);
    const float nx = fmaf(a, fmaf(L.ux, uzc, -__fmul_rn(L.uy, e.sp)), __fmul_rn(L.ux, e.ct));
    const float ny = fmaf(a, fmaf(L.uy, uzc, __fmul_rn(L.ux, e.sp)), __fmul_rn(L.uy, e.ct));
    const float nz = fmaf(-__fmul_rn(q, rq), e.cp, __fmul_rn(L.uz, e.ct));
    L.ux = nx; L.uy = ny; L.uz = nz;
    L.z = fmaf(e.dtau, nz, L.z);
    L.path_lo = __fadd_rn(L.path_lo, e.dtau);
    L.key = e.key;
}

__device__ __forceinline__ void scatter_and_move(Lane &L, const HotRow &H, uint32_t w_hg, uint32_t w_az, uint32_t w_fp)
{
    const Prepared e = prepare_event(H, w_hg, w_az, w_fp);
    apply_event(L, e);
}

// "Something may have happened": the photon left the slab, or its key is at/above the row's coarse threshold
// t_hot = min(t16, RENORM_KEY) << 16 -- every possible absorption, plus a 2^-10 chance per event that only serves to
// renormalise the direction and flush the path accumulator (a pseudo-random but per-photon deterministic schedule,
// mean period <= 1024 events, with no extra instruction in the loop).  Resolved later, by resolve().
__device__ __forceinline__ bool needs_attention(float neg_tau, const Lane &L, uint32_t thi)
{
    return L.z > 0.0f || L.z < neg_tau || L.key >= thi;
}

// Resolve the reference's termination chain monte_carlo3D.py:1390-1466, in its order, for the event L.i that
// just moved the photon to L.z along L.uz.  Executed by all lanes of a warp that need it at once (deferred), so it
// is off the hot path.  On a Lambertian-bottom reflection it also performs the NEXT event (1238-1250: cosine-law
// rejection sampling about +z) so that the hot loop never carries a bottom_reflection flag.
// Returns the condition (0 = keep walking; the lane's state is then ready for the next group of its walk stream).
template <bool IMP>
__device__ __forceinline__ uint32_t resolve(const WalkParams &P, const DevCase &C, const uint32_t phi, const DevRow &R, Lane &L)
{
    uint32_t cond = ALIVE;
    if (L.z > 0.0f) {   // reflected (1390-1397); z - z_prev = dtau muz, so the overshoot path is z / muz
        L.path_lo = __fsub_rn(L.path_lo, __fdividef(L.z, L.uz));
        cond = 1u;
    } else if (L.z < C.neg_tau_tot) {   // 1399-1459
        L.path_lo = __fsub_rn(L.path_lo, __fdividef(__fadd_rn(L.z, C.tau_tot), L.uz));
        L.z = C.neg_tau_tot;
        cond = (L.i == 1u) ? 3u : 2u;
        if (C.lambert_bottom) {
            const uint4 b = philox4x32(L.i, TAG_LAMBERT, L.plo, phi, P.rk);
            if ((long long)b.x <= C.refl_thr) {
                // ---- reflected by the Lambertian bottom: event i+1 happens here, on its own TAG_LAMBERT blocks ----
                L.i += 1u;
                const uint4 w = philox4x32(L.i, TAG_LAMBERT, L.plo, phi, P.rk);
                float ct, st;
                for (uint32_t j = 0;; ++j) {
                    const uint4 a = philox4x32(L.i, TAG_LAMBERT | ((1u + (j >> 1)) << 8), L.plo, phi, P.rk);
                    const float u_t = u32_to_unit((j & 1u) ? a.z : a.x);
                    const float r1 = u32_to_unit((j & 1u) ? a.w : a.y);
                    float s_, c_;
                    sincosf(1.5707963267948966f * u_t, &s_, &c_);
                    if (r1 < 2.0f * s_ * c_) { ct = c_; st = s_; break; }
                }
                float cp, sp;
                azimuth(w.y, cp, sp);
                L.ux = __fmul_rn(st, cp); L.uy = __fmul_rn(st, sp); L.uz = ct;   // muz_0 == 1 branch, 1262-1265
                const float dt2 = free_path(w.z);
                L.z = fmaf(dt2, ct, C.neg_tau_tot);
                L.path_lo = __fadd_rn(L.path_lo, dt2);
                L.key = w.w;
                L.imp = IMP ? species_is_impurity(P, R, L.i, L.plo, phi) : false;
                cond = ALIVE;
                if (L.z > 0.0f) {
                    L.path_lo = __fsub_rn(L.path_lo, __fdividef(L.z, L.uz));
                    cond = 1u;
                }
            }
        }
    }
    const uint32_t key16 = L.key >> 16;
    if (cond == ALIVE) {   // 1461-1466: absorbed iff K40 = key16 << 24 | fine24 >= T40
        const uint32_t t16 = L.imp ? R.ti16 : R.t16;
        bool absorbed = key16 > t16;
        if (key16 == t16) {   // probability 2^-16: the low 24 bits of the variate come from the event's TAG_FINE block
            const uint4 f = philox4x32(L.i, TAG_FINE, L.plo, phi, P.rk);
            absorbed = (f.x >> 8) >= (L.imp ? R.ti24 : R.t24);
        }
        if (absorbed) cond = L.imp ? 5u : 4u;
    }
    if (cond == ALIVE && key16 >= RENORM_KEY) {   // keyed on the photon's own random stream: scheduling independent
        const float rn = rsqrt_fast(fmaf(L.ux, L.ux, fmaf(L.uy, L.uy, L.uz * L.uz)));
        L.ux = __fmul_rn(L.ux, rn); L.uy = __fmul_rn(L.uy, rn); L.uz = __fmul_rn(L.uz, rn);
        L.path_hi = __fadd_rn(L.path_hi, L.path_lo);
        L.path_lo = 0.0f;
    }
    return cond;
}

// Bookkeeping after the move of event L.i: the species draw (impurity runs only) and the attention predicate.
// Returns true while the photon simply keeps walking.
template <bool IMP>
__device__ __forceinline__ bool after_move(const WalkParams &P, const uint32_t phi, const DevRow *rows, uint32_t rows_addr, Lane &L,
                                           const HotRow &H)
{
    uint32_t thi = H.t_hot;
    if (IMP) {
        const DevRow &R = rows[lane_row(L, rows_addr)];
        L.imp = species_is_impurity(P, R, L.i, L.plo, phi);
        thi = L.imp ? H.ti_hot : thi;
    }
    return !needs_attention(H.neg_tau, L, thi);
}

// One scattering event of the photon in L from three words of its walk stream, up to and including the attention
// predicate.  Returns true while the photon simply keeps walking.
template <bool IMP>
__device__ __forceinline__ bool event(const WalkParams &P, const uint32_t phi, const DevRow *rows, uint32_t rows_addr, Lane &L,
                                      const HotRow &H, uint32_t w_hg, uint32_t w_az, uint32_t w_fp)
{
    L.i += 1u;
    scatter_and_move(L, H, w_hg, w_az, w_fp);
    return after_move<IMP>(P, phi, rows, rows_addr, L, H);
}

// One GROUP of the photon's walk stream: up to four events on three Philox blocks (twelve words, three per event).
// The lane stops at the first event that needs attention; if the photon survives it continues with the NEXT group
// (L.blk already points there).  EAGER computes the three blocks up front (three independent multiply chains that
// overlap the events' arithmetic: the latency-oriented form used when a warp runs almost alone, i.e. while a launch
// drains); otherwise each block is computed right before the event that first needs it.
template <bool IMP, bool EAGER, bool SWEEP = false>
__device__ __forceinline__ bool group(const WalkParams &P, const DevRow *rows, uint32_t rows_addr, Lane &L)
{
    const HotRow H = load_hot_row<SWEEP>(P, L.row_addr);
    const uint32_t phi = lane_phi<SWEEP>(P, L);
    const uint32_t n = L.blk;
    L.blk = n + GROUP_BLOCKS;
    const uint4 a = philox_walk(n, phi, L.pk, P.rk);
    uint4 b, c;
    if (EAGER) {
        b = philox_walk(n + 1u, phi, L.pk, P.rk);
        c = philox_walk(n + 2u, phi, L.pk, P.rk);
    }
    if (!event<IMP>(P, phi, rows, rows_addr, L, H, a.x, a.y, a.z)) return false;
    if (!EAGER) b = philox_walk(n + 1u, phi, L.pk, P.rk);
    if (!event<IMP>(P, phi, rows, rows_addr, L, H, a.w, b.x, b.y)) return false;
    if (!EAGER) c = philox_walk(n + 2u, phi, L.pk, P.rk);
    if (!event<IMP>(P, phi, rows, rows_addr, L, H, b.z, b.w, c.x)) return false;
    return event<IMP>(P, phi, rows, rows_addr, L, H, c.y, c.z, c.w);
}

// The same group for a warp that runs (almost) alone on its scheduler -- the tail of a launch, where the time is the
// dependent chain of the longest walk, not issue slots.  The three blocks and the direction-independent halves of
// all four events are computed first (independent streams the scheduler can overlap); what remains sequential per
// event is the rotation: ~1/2 of the latency of the throughput form.  Same values, same stream.
template <bool IMP, bool SWEEP = false>
__device__ __forceinline__ bool group_latency(const WalkParams &P, const DevRow *rows, uint32_t rows_addr, Lane &L)
{
    const HotRow H = load_hot_row<SWEEP>(P, L.row_addr);
    const uint32_t phi = lane_phi<SWEEP>(P, L);
    const uint32_t n = L.blk;
    L.blk = n + GROUP_BLOCKS;
    const uint4 a = philox_walk(n, phi, L.pk, P.rk);
    const uint4 b = philox_walk(n + 1u, phi, L.pk, P.rk);
    const uint4 c = philox_walk(n + 2u, phi, L.pk, P.rk);
    const Prepared e0 = prepare_event(H, a.x, a.y, a.z);
    const Prepared e1 = prepare_event(H, a.w, b.x, b.y);
    const Prepared e2 = prepare_event(H, b.z, b.w, c.x);
    const Prepared e3 = prepare_event(H, c.y, c.z, c.w);
    if (IMP) {   // the species draw sits between the events: plain early exits
        L.i += 1u; apply_event(L, e0);
        if (!after_move<IMP>(P, phi, rows, rows_addr, L, H)) return false;
        L.i += 1u; apply_event(L, e1);
        if (!after_move<IMP>(P, phi, rows, rows_addr, L, H)) return false;
        L.i += 1u; apply_event(L, e2);
        if (!after_move<IMP>(P, phi, rows, rows_addr, L, H)) return false;
        L.i += 1u; apply_event(L, e3);
        return after_move<IMP>(P, phi, rows, rows_addr, L, H);
    }
    // Branch-free: all four events are applied, and the lane keeps the state after the first one that needs
    // attention (selects, no early exits), so that nothing separates the special functions of the later events from
    // the top of the group (the compiler otherwise sinks them below the exits, back onto the dependent chain).
    const Prepared ev[4] = {e0, e1, e2, e3};
    bool go = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Lane N = L;
        N.i += 1u;
        apply_event(N, ev[k]);
        const bool att = needs_attention(H.neg_tau, N, H.t_hot);
        L.z = go ? N.z : L.z; L.ux = go ? N.ux : L.ux; L.uy = go ? N.uy : L.uy; L.uz = go ? N.uz : L.uz;
        L.path_lo = go ? N.path_lo : L.path_lo; L.key = go ? N.key : L.key; L.i = go ? N.i : L.i;
        go = go && !att;
    }
    return go;
}

// ---- tables in shared memory: the SSP rows, then (sweep launches) the cases ------------------------------------------
__host__ __device__ __forceinline__ size_t tables_bytes(int n_rows, uint32_t n_cases)
{
    return (size_t)n_rows * sizeof(DevRow) + (size_t)n_cases * sizeof(DevCase);
}

// Copies the launch's rows and cases into shared memory at `smem` (16-byte aligned); the caller synchronises.
__device__ __forceinline__ void stage_tables(const WalkParams &P, unsigned char *smem, int n_threads)
{
    uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
    const int n_row_words = P.n_rows * (int)(sizeof(DevRow) / 4);
    for (int k = threadIdx.x; k < n_row_words; k += n_threads) dst[k] = reinterpret_cast<const uint32_t *>(P.rows)[k];
    const int n_case_words = (int)P.n_cases * (int)(sizeof(DevCase) / 4);
    for (int k = threadIdx.x; k < n_case_words; k += n_threads)
        dst[n_row_words + k] = reinterpret_cast<const uint32_t *>(P.cases)[k];
}
__device__ __forceinline__ const DevCase *staged_cases(const WalkParams &P, const unsigned char *smem)
{
    return reinterpret_cast<const DevCase *>(smem + (size_t)P.n_rows * sizeof(DevRow));
}

// The case of photon `pid` of a sweep launch: the last one whose first photon is <= pid.
template <bool SWEEP>
__device__ __forceinline__ uint32_t find_case(const WalkParams &P, const DevCase *cases, uint32_t pid)
{
    if (!SWEEP) return 0u;
    uint32_t lo = 0u, hi = P.n_cases;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cases[mid].pid_first <= pid) lo = mid;
        else hi = mid;
    }
    return lo;
}

// ---- per-photon prologue (shared by the init kernel and the fused kernel) ----------------------------------------

// One photon in Lambertian_surface mode, events 2, 3, ... after an event 1 that did not absorb it.  Event 1 does
// not move (dtau = 0, monte_carlo3D.py:1228-1229) and is absorbed with probability 1 - R (ssa_event = R,
// 1385-1387); every later event re-emits the photon from the surface with the cosine law (1238-1250) and almost
// surely leaves through the top on event 2.  Every event >= 2 is a Lambertian reflection event (TAG_LAMBERT blocks).
template <bool IMP>
__device__ __noinline__ uint32_t lambert_surface_walk(const WalkParams &P, const DevCase &C, const uint32_t phi, const DevRow &R, Lane &L)
{
    for (;;) {
        // termination chain for event L.i: z > 0, (z < -tau_tot cannot happen), absorbed by the surface
        if (L.z > 0.0f) {
            L.path_lo = __fsub_rn(L.path_lo, __fdividef(L.z, L.uz));
            return 1u;
        }
        const uint32_t key16 = L.key >> 16;
        bool absorbed = key16 > C.surf_t16;
        if (key16 == C.surf_t16) absorbed = (philox4x32(L.i, TAG_FINE, L.plo, phi, P.rk).x >> 8) >= C.surf_t24;
        if (absorbed) return L.imp ? 5u : 4u;
        L.i += 1u;
        const uint4 w = philox4x32(L.i, TAG_LAMBERT, L.plo, phi, P.rk);
        float ct, st;
        for (uint32_t j = 0;; ++j) {
            const uint4 a = philox4x32(L.i, TAG_LAMBERT | ((1u + (j >> 1)) << 8), L.plo, phi, P.rk);
            const float u_t = u32_to_unit((j & 1u) ? a.z : a.x);
            const float r1 = u32_to_unit((j & 1u) ? a.w : a.y);
            float s_, c_;
            sincosf(1.5707963267948966f * u_t, &s_, &c_);
            if (r1 < 2.0f * s_ * c_) { ct = c_; st = s_; break; }
        }
        float cp, sp;
        azimuth(w.y, cp, sp);
        L.ux = __fmul_rn(st, cp); L.uy = __fmul_rn(st, sp); L.uz = ct;
        const float dt = free_path(w.z);
        L.z = fmaf(dt, ct, L.z);
        L.path_lo = __fadd_rn(L.path_lo, dt);
        L.key = w.w;
        L.imp = IMP ? species_is_impurity(P, R, L.i, L.plo, phi) : false;
    }
}

// Wavelength draw (monte_carlo3D.py:1515-1520: np.around(np.random.normal(wvl0, scale), 2), Box-Muller on the photon's
// TAG_FIRST block; the rounded value is an index into the SSP table) and the first event: the three draws of
// initial_pdfs (1035-1038), no deflection (1232-1237), move, direct-transmission / Lambertian-bottom /
// first-extinction absorption tests (1399-1466).  On return L holds the photon's state and `row` its SSP row.
// Returns the condition: ALIVE = the photon walks on, with L ready for group 0 of its walk stream.  `redo` is 0
// unless event 1 needed attention (resolve() ran: the photon may have been reflected by a Lambertian bottom, in
// which case event 2 has been performed as well and L.i == 2, or its direction renormalised); it then holds what a
// later resolve() of the same event needs besides the free path: key16 << 16 | impurity << 1 | 1 (Fresh::redo).
template <bool IMP>
__device__ __forceinline__ uint32_t first_event(const WalkParams &P, const DevCase &C, const uint32_t phi, const DevRow *rows,
                                                uint32_t rows_addr, uint32_t plo, Lane &L, uint32_t &row, float &dtau, uint32_t &redo)
{
    const uint4 w = philox4x32(0u, TAG_FIRST, plo, phi, P.rk);
    // Box-Muller with one MUFU per function: |error| of zn ~1e-6, i.e. 4e-6 of a 0.01 um wavelength bin at the
    // default band width (a photon in 10^5 lands in the row next to the one fp64 arithmetic picks: the rint below)
    const float zn = __fmul_rn(sqrt_fast(__fmul_rn(-1.3862943611198906f, lg2_fast(u32_to_unit(w.x)))),
                               -cos_fast(fmaf(__uint2float_rn(w.y), 1.4629180792671596e-09f, -3.1415926528583587f)));
    const int r = (int)rint(fma(C.sigma_x100, (double)zn, C.wvl0_x100)) - C.k_first;
    row = C.row_begin + (uint32_t)max(0, min(C.n_rows - 1, r));
    const DevRow &R = rows[row];
    L.plo = plo; L.phi = phi; L.row_addr = rows_addr + row * (uint32_t)sizeof(DevRow); L.blk = 0u; L.i = 1u;
    L.ux = C.mu0x; L.uy = 0.0f; L.uz = C.mu0z; L.path_hi = 0.0f;
    L.key = w.w;
    L.imp = IMP ? species_is_impurity(P, R, 1u, plo, phi) : false;
    L.pk = philox_walk_constants(plo, P.rk);
    redo = 0u;
    if (C.lambert_surface) {
        dtau = 0.0f;
        L.z = 0.0f; L.path_lo = 0.0f;
        // (a copy goes through the out-of-line call: passing L itself would pin the caller's lane state to the stack for
        // every photon of every mode -- local stores on the common path)
        Lane T = L;
        const uint32_t c = lambert_surface_walk<IMP>(P, C, phi, R, T);
        L = T;
        return c;
    }
    dtau = free_path(w.z);
    L.z = __fmul_rn(dtau, C.mu0z);
    L.path_lo = dtau;
    if (L.z < C.neg_tau_tot || (w.w >> 16) >= (L.imp ? R.ti16 : R.t16)) {
        redo = (w.w & 0xffff0000u) | (L.imp ? 2u : 0u) | 1u;
        return resolve<IMP>(P, C, phi, R, L);
    }
    return ALIVE;
}

// Finish (store the raw record, free the lane) or resume a lane whose last event needed attention.
template <bool IMP, bool SWEEP>
__device__ __forceinline__ bool resolve_lane(const WalkParams &P, const DevCase *cases, const DevRow *rows, uint32_t rows_addr, Lane &L)
{
    const uint32_t row = lane_row(L, rows_addr);
    const DevCase &C = lane_case<SWEEP>(P, cases, L);
    const uint32_t cond = resolve<IMP>(P, C, lane_phi<SWEEP>(P, L), rows[row], L);
    if (cond == ALIVE) return true;
    store_raw(P, lane_pid(C, L), L.ux, L.uy, L.uz, __fadd_rn(L.path_hi, L.path_lo), L.i - 1u, cond, row, lane_lcase<SWEEP>(P, L));
    L.i = 0u;
    return false;
}


}  // namespace mc3d
