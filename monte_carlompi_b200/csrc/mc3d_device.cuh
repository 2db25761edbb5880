// mc3d_device.cuh -- types shared by the walk / finalize / replay kernels and the host runtime.
//
// Random-number layout (identical in oracle/mc3d_oracle.c, which is how production mode is checked):
//   Philox4x32-7 (Salmon et al., SC'11: seven rounds is the Crush-resistant round count of Philox4x32; Random123's
//   philox4x32_R(7, ...)), key = (seed_lo, seed_hi), counter = (c0, c1, pid_lo, pid_hi), pid = global photon id.
//   c1 low byte is the stream tag, c1 >> 8 a sub-block:
//     TAG_WALK       c0 = block number 3 G + b of the photon's walk stream.  The walk stream is consumed in GROUPS of
//                    four events = three blocks = twelve words; slot s of group G uses words 3s, 3s+1, 3s+2:
//                      word 3s   -> r1 of Henyey_Greenstein2 (reference monte_carlo3D.py:915-916)
//                      word 3s+1 -> azimuth (921)
//                      word 3s+2 -> free path (1014)
//                    (the float conversions of these words ignore their low byte) and the event's 40-bit
//                    single-scatter-albedo variate (1020) is K40 = key16 << 24 | fine24 with
//                      key16  = (low byte of word 3s) << 8 | low byte of word 3s+2
//                      fine24 = top 24 bits of word 0 of block (c0 = event number i, TAG_FINE) -- only ever needed
//                               when key16 equals the top 16 bits of the threshold (probability 2^-16 per event).
//                    An event "needs attention" when the photon left the slab or key16 >= min(T40 >> 24, 0xffc0)
//                    (possible absorption; above 0xffc0 the direction is renormalised).  A photon that survives an
//                    event needing attention continues with slot 0 of the NEXT group: the rest of the current
//                    group is skipped.  This is a property of the photon's own history, so the stream -- and every
//                    result -- is independent of how lanes, warps and GPUs are scheduled.
//     TAG_SPECIES    c0 = i >> 2, word i & 3 -> ice/impurity choice of event i (1023); drawn only when an impurity is present
//     TAG_LAMBERT    c0 = i.  Sub-block 0: word 0 -> bottom reflectance draw of event i (1422/1453); words 1, 2, 3 ->
//                    azimuth, free path and absorption variate (K40 = w3 << 8 | w1 & 0xff) of event i when it IS a
//                    Lambertian reflection (the event after a reflecting bottom hit, or any event >= 2 in
//                    Lambertian_surface mode).  Sub-block 1+(j>>1), words 2(j&1), 2(j&1)+1 -> attempt j of the
//                    cosine-law rejection loop (1245-1246)
//     TAG_FIRST      c0 = 0: w0, w1 -> Box-Muller normal for the photon's wavelength (1519); w2 -> free path of event 1,
//                    (w3 << 8 | w2 & 0xff) -> absorption variate of event 1 (initial_pdfs, 1035-1038)
//     TAG_FINE       c0 = i: see TAG_WALK
//   A 32-bit word w maps to the open-interval uniform (w + 0.5) 2^-32.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mc3d {

constexpr uint32_t TAG_WALK = 0, TAG_SPECIES = 1, TAG_LAMBERT = 2, TAG_FIRST = 3, TAG_FINE = 4;
constexpr int PHILOX_ROUNDS = 7;
constexpr uint32_t GROUP_EVENTS = 4, GROUP_BLOCKS = 3;
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr int N_COND = 8;

// Per-wavelength row in the form the walk consumes (built on the host from mc3d_ssp_row, fp64 -> fp32/integer).
struct DevRow {
    // -- first 32 bytes: what the event loop loads (two 16-byte shared-memory loads)
    float one_m_g;    // 1 - g
    float one_m_g2;   // 1 - g^2
    float d_scale;    // 2 g 2^-32:      D = 1 - g + 2 g r = fma(float(w), d_scale, d_off), r = (w + 1/2) 2^-32
    float d_off;      // 1 - g + g 2^-32
    uint32_t t_hot;   // coarse "needs attention" threshold on the event's key word (key16 in its top half):
                      // min(t16, RENORM_KEY) << 16.  It fires on every possible absorption and, with probability
                      // 2^-10 per event, just to renormalise
    float omr_scale;  // 1 - r = fma(float(w), omr_scale, omr_off): (-2^-32, 1).  For g == 0 rows (+2^-32, 2^-33),
    float omr_off;    //   i.e. r itself, which maps the factored HG form onto the reference's `1 - 2r` branch
    uint32_t ti_hot;  // t_hot of the impurity species
    // -- resolve / prologue only
    uint32_t t16;     // ice: absorbed iff K40 >= T40 = ceil(ssa 2^40 - 1/2); t16 = T40 >> 24 (0x10000: never absorbed)
    uint32_t t24;     //      t24 = T40 & 0xffffff
    uint32_t ti16;    // same for the impurity's single-scatter albedo
    uint32_t ti24;
    uint32_t s_last;  // impurity iff species word <= s_last (and s_any)
    uint32_t s_any;   // 0 when P_ext_imp == 0 (species word never selects the impurity)
    float inv_ext;    // ln 2 / (ext_cff_mss rho_snw): metres per unit of the walk's depth scale (optical depth / ln 2)
    float neg_tau_tot;// DevCase::neg_tau_tot of the case the row belongs to (sweep launches read it with the row)
};
static_assert(sizeof(DevRow) == 64, "DevRow is 64 bytes");
constexpr uint32_t RENORM_KEY = 0xffc0u;   // a key16 at/above this also triggers renormalisation (2^-10 per event)

// Raw result of one walk (32 B, one sector, written by the lane that finished the photon).
struct __align__(16) RawResult {
    float ux, uy, uz;   // final direction cosines
    float path_tau;     // path inside the slab in units of ln 2 optical depths (DevRow::inv_ext converts to metres)
    uint32_t n_scat;    // i - 1
    uint32_t meta;      // condition | row << 8   (row in the launch's table)
    uint32_t lcase;     // sweep launches: case index in the launch
    uint32_t pad1;
};

// A photon after its first event, waiting for a lane of the walk kernel (written by the init kernel: entry pid of the
// list is photon pid of the launch).
constexpr uint32_t FRESH_DEAD = 0xffffffffu;
struct __align__(16) Fresh {
    uint32_t pid;    // photon offset in this launch
    uint32_t row;    // SSP row in the launch's table | (case index in the launch) << 12   (sweep launches);
                     // FRESH_DEAD: the photon ended on its first event (nothing to walk)
    float dtau;      // free path of the first event
    uint32_t redo;   // != 0: event 1 needed attention and the walk kernel redoes its termination chain (resolve());
                     // key16 << 16 | impurity << 1 | 1 of that event (walk_device.cuh: first_event)
};

// Scalars of one case (one MonteCarlo.run: theta_0, tau_tot, the wavelength band, the Lambertian options) as the
// device code reads them.  A single-case launch carries one in its kernel parameters (constant bank); a sweep launch
// (mc3d_run_sweep: many cases, one launch, a concatenated SSP table) stages one per case in shared memory and every
// lane finds its own through the high word of its photon id.
struct __align__(16) DevCase {
    uint64_t id0;           // global photon id of the launch's photon 0 under this case's numbering: id = id0 + pid
    uint32_t pid_first;     // first photon of the case in this launch (cases are contiguous, ascending pid ranges)
    uint32_t row_begin;     // first SSP row of the case in the launch's table (records hold row - row_begin)
    float mu0x, mu0z;       // sin(theta0), -cos(theta0)
    float neg_tau_tot;      // -tau_tot / ln 2: depths and paths are carried in units of ln 2 optical depths, so that
    float tau_tot;          //  tau_tot / ln 2   a free path is just -log2(u) (no multiply by ln 2 per event)
    double wvl0_x100, sigma_x100;  // wavelength draw in units of 0.01 um
    int64_t refl_thr;       // bottom reflects iff (int64)w <= refl_thr  (U(w) <= R)
    int32_t k_first;
    int32_t n_rows;         // rows of this case
    uint32_t lambert_bottom;
    uint32_t lambert_surface;   // run(Lambertian_surface=True): the prologue finishes every photon by itself
    uint32_t surf_t16, surf_t24;   // 40-bit threshold of the surface reflectance (ssa_event = R, monte_carlo3D.py:1385-1387)
};
static_assert(sizeof(DevCase) == 80, "DevCase is 80 bytes");

struct WalkParams {
    uint32_t rk[2 * PHILOX_ROUNDS];   // Philox round keys: rk[2r], rk[2r+1] for round r
    DevCase c;              // the case of a single-case launch (sweep launches: `cases`)
    int32_t n_rows;         // rows of the launch's table (all cases)
    uint32_t n_cases;       // sweep launches: cases overlapping this launch (staged in shared memory); else 0
    uint32_t case0;         // sweep launches: global index of cases[0] (a photon's case index is id >> 40)
    uint32_t refill_threshold;
    uint32_t drain_give;    // drain phase: a warp with <= this many photons hands them to the block's pool (0 = off)
    uint32_t drain_latency; // drain phase: latency-oriented groups (the launch runs alone: its tail is a dependent chain)
    uint32_t n_photon;      // photons in this launch (< 2^31)
    uint32_t claim;         // walk kernel: entries of `fresh` a warp claims per atomicAdd (32, 64 or 96)
    const DevRow *rows;     // [n_rows], global
    const DevCase *cases;   // [n_cases], global (sweep launches)
    uint32_t *counter;      // walk kernel: next unclaimed entry of `fresh`
    Fresh *fresh;           // [n_photon]
    RawResult *raw;         // [n_photon]
    // hand-over to the tail kernel (a call that runs alone; null = the walk kernel drains by itself)
    uint32_t *tail;         // [TAIL_WORDS][tail_cap]: state of the photons still walking when the fresh list ran out
    uint32_t *n_tail;       // entries in `tail`
    uint32_t tail_cap;      // >= lanes of the walk kernel's grid
    uint32_t pad;
};
constexpr int TAIL_WORDS = 11;   // z, ux, uy, uz, path_lo, path_hi, i, plo, phi, row, blk

struct FinalizeParams {
    const RawResult *raw;
    const DevRow *rows;
    const double *edges;     // [n_theta_bins + 1] = np.linspace(0, pi/2, n + 1), then [n_phi_bins + 1] = np.linspace(0, 2 pi, m + 1)
    uint32_t n_photon;
    int32_t n_rows;
    int32_t n_theta_bins;
    int32_t n_phi_bins;      // <= 1: zenith histogram only
    int32_t use_smem;        // tally in shared memory: 0 no (global atomics), 1 the whole table, 2 one case's rows at a time,
                             // 3 the outcome counts of every row (BRF bins global)
    int32_t win_rows;        // sweep launches: the largest n_rows of the launch's cases (the window of use_smem == 2)
    // record columns (device), any may be null
    uint8_t *condition;
    int16_t *wvl_row;
    float *theta_n;
    float *phi_n;
    uint32_t *n_scat;
    float *path_length;
    uint4 *packed;               // 16-byte packed records (mc3d_records.packed) instead of the columns, or null
    unsigned long long *tally;   // [n_rows][N_COND + n_theta_bins * max(1, n_phi_bins)] or null
    unsigned long long *n_events;
    // optional n_scat / path-length histograms (mc3d_hist_spec) and the always-on extrema
    uint32_t *extrema;           // [4] ~min n_scat, max n_scat, ~min path, max path (float bits; path >= 0 orders as uint)
    unsigned long long *hist;    // [n_scat_bins + path_bins] or null
    const double *hist_edges;    // [n_scat_bins + 1] then [path_bins + 1], np.linspace of the requested ranges
    int32_t n_scat_bins, path_bins;
    int32_t hist_smem;           // histogram staged in shared memory behind the tally block
    double path_scale;
    // sweep launches
    const DevCase *cases;        // [n_cases] or null: records hold row - cases[lcase].row_begin
    uint32_t n_cases, case0;
    unsigned long long *case_events;   // [total cases of the call] or null: events per case, index case0 + lcase
};

// fp64 replay mode (replay_kernel.cu): per-photon inputs exactly as the reference holds them
struct ReplayParams {
    double theta0_rad, tau_tot, rho_snw, r_lambert;
    uint32_t flags;
    uint32_t n_photon;
    const double *wvl, *ssa_ice, *ssa_imp, *g, *ext_cff_mss, *p_ext_imp;
    const double *init_draws;      // [3 n]
    const long long *offsets;      // [n + 1]
    const double *stream;
    int *condition;
    double *wvn, *theta_n, *phi_n, *path_length, *snow_depth;
    long long *n_scat, *consumed;
};

__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                                             uint32_t k1)
{
    const uint32_t h0 = __umulhi(PHILOX_M0, c0), l0 = PHILOX_M0 * c0;
    const uint32_t h1 = __umulhi(PHILOX_M1, c2), l1 = PHILOX_M1 * c2;
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
}

// rk: precomputed round keys rk[2r], rk[2r+1] (kernel parameter space -> constant bank operands of the LOP3s)
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            const uint32_t *__restrict__ rk)
{
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) philox_round(c0, c1, c2, c3, rk[2 * r], rk[2 * r + 1]);
    return make_uint4(c0, c1, c2, c3);
}

// Walk-stream block (tag TAG_WALK = 0) with the photon-constant part of rounds 1 and 2 hoisted out of the walk:
//   round 1 multiplies M1 by counter word 2 = pid_lo, round 2 multiplies M0 by the round-1 word 0, which depends on
//   pid only.  pB = lo(M1 pid_lo), (pC, pD) = mulhilo(M0, hi(M1 pid_lo) ^ TAG_WALK ^ rk[0]) are computed once per
//   photon (philox_walk_constants); 2 R - 2 instead of 2 R wide multiplies per block, identical output.
struct PhiloxWalkConst { uint32_t pB, pC, pD; };

__device__ __forceinline__ PhiloxWalkConst philox_walk_constants(uint32_t plo, const uint32_t *__restrict__ rk)
{
    PhiloxWalkConst k;
    const uint32_t a = __umulhi(PHILOX_M1, plo) ^ TAG_WALK ^ rk[0];
    k.pB = PHILOX_M1 * plo;
    k.pC = __umulhi(PHILOX_M0, a);
    k.pD = PHILOX_M0 * a;
    return k;
}

// == philox4x32(n, TAG_WALK, plo, phi)
__device__ __forceinline__ uint4 philox_walk(uint32_t n, uint32_t phi, const PhiloxWalkConst k, const uint32_t *__restrict__ rk)
{
    // round 1: (c0, c1, c2, c3) = (n, TAG_WALK, plo, phi)
    uint32_t c2 = __umulhi(PHILOX_M0, n) ^ phi ^ rk[1];
    uint32_t c3 = PHILOX_M0 * n;
    // round 2: word 0 of round 1 is photon-constant, its products are pC (hi) and pD (lo)
    uint32_t c0 = __umulhi(PHILOX_M1, c2) ^ k.pB ^ rk[2];
    uint32_t c1 = PHILOX_M1 * c2;
    c2 = k.pC ^ c3 ^ rk[3];
    c3 = k.pD;
#pragma unroll
    for (int r = 2; r < PHILOX_ROUNDS; ++r) philox_round(c0, c1, c2, c3, rk[2 * r], rk[2 * r + 1]);
    return make_uint4(c0, c1, c2, c3);
}

// 40-bit absorption test K40 >= T40 with T40 = t16 << 24 | t24 (t16 = 0x10000: never)
__device__ __forceinline__ bool absorbed40(unsigned long long k40, uint32_t t16, uint32_t t24)
{
    return k40 >= (((unsigned long long)t16 << 24) | t24);
}

__device__ __forceinline__ float u32_to_unit(uint32_t w)
{
    // (w + 0.5) 2^-32, never 0; rounds to 1.0f only for w > 0xffffff7f where -log gives exactly 0
    return fmaf(__uint2float_rn(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

}  // namespace mc3d
