/*
 * mc3d.h -- C ABI of libmc3d.so: the B200-native photon random walk behind monte_carloMPI's Python driver.
 *
 * The reference (amschne/monte_carloMPI) is 100 % Python and has no FFI of its own; its hot path is the inline
 * method MonteCarlo.monte_carlo3D (monte_carloMPI/monte_carlo3D.py:1111-1490) driven by the loop at
 * monte_carlo3D.py:1613-1616, bracketed by the mpi4py scatter / gather in monte_carloMPI/parallelize.py:19,36.
 * This header is the boundary a maintainer binds with ctypes/cffi to replace exactly that loop (see
 * INTEGRATION.md for the stub).  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every buffer is caller-allocated and C-contiguous; the library never keeps
 *     a caller pointer after a call returns;
 *   - every function returns 0 on success and a negative MC3D_E* code on failure; mc3d_last_error() returns a
 *     thread-local, library-owned, NUL-terminated description of the last failure on the calling thread;
 *   - all calls are synchronous unless they say otherwise (ctypes releases the GIL around them);
 *   - there is no CPU fallback: without a CUDA device mc3d_create* fails with MC3D_ENODEVICE.
 */
#ifndef MC3D_H_
#define MC3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC3D_ABI_VERSION 4

/* error codes */
#define MC3D_OK 0
#define MC3D_EINVAL (-1)    /* bad argument                                  */
#define MC3D_ENODEVICE (-2) /* no usable CUDA device                         */
#define MC3D_ECUDA (-3)     /* a CUDA runtime call failed                    */
#define MC3D_ENCCL (-4)     /* NCCL missing or an NCCL call failed           */
#define MC3D_ENOMEM (-5)    /* host or device allocation failed              */
#define MC3D_ESTREAM (-6)   /* replay: recorded stream exhausted / malformed */

/* mc3d_params.flags */
#define MC3D_FLAG_LAMBERT_BOTTOM 1u  /* run(Lambertian_bottom=True), monte_carlo3D.py:1421-1428, 1452-1459   */
#define MC3D_FLAG_LAMBERT_SURFACE 2u /* run(Lambertian_surface=True), monte_carlo3D.py:1228-1229, 1385-1387 */

/* photon outcome codes, monte_carlo3D.py:1390-1466 (README.md:62-66; 5 is undocumented there) */
#define MC3D_COND_REFLECTED 1
#define MC3D_COND_DIFFUSE_TRANSMITTED 2
#define MC3D_COND_DIRECT_TRANSMITTED 3
#define MC3D_COND_ABSORBED_ICE 4
#define MC3D_COND_ABSORBED_IMPURITY 5
#define MC3D_N_SLOTS 16
#define MC3D_N_COND 8 /* tally stride per wavelength row: index = condition, 0 = total launched, 6..7 unused */

typedef struct mc3d_ctx mc3d_ctx; /* opaque; one per process (or several); not thread-safe per context */

/* Scalars of one MonteCarlo.run call (monte_carlo3D.py:1492-1521 and config.ini). */
typedef struct mc3d_params {
    double theta0_rad;   /* self.theta_0 = pi * theta_0 / 180, monte_carlo3D.py:1508                       */
    double tau_tot;      /* snow optical depth, config.ini:5                                               */
    double rho_snw;      /* snow density [kg m-3], config.ini:11 (path lengths in metres)                  */
    double r_lambert;    /* Lambertian_reflectance, monte_carlo3D.py:1506                                  */
    double wvl0_um;      /* centre wavelength [um], monte_carlo3D.py:1519                                  */
    double sigma_um;     /* half_width / 2.355, monte_carlo3D.py:1516                                      */
    int32_t k_first;     /* table row r holds wavelength (k_first + r) / 100 um (np.around(.., 2) grid)     */
    uint32_t flags;      /* MC3D_FLAG_*                                                                    */
    int32_t n_theta_bins;/* BRF zenith bins over [0, pi/2] (post_processing.py:73-76, 435-444); 0 = none   */
    int32_t n_phi_bins;  /* azimuth bins over [0, 2 pi] for a full-hemisphere (theta, phi) BRF; 0 or 1 = zenith
                            only (the reference stores phi_n but never bins it)                               */
} mc3d_params;

/* One row of the per-wavelength SSP table, the de-duplicated form of the per-photon arrays the reference
 * builds at monte_carlo3D.py:1524-1588, 1612 (one row per distinct rounded wavelength). */
typedef struct mc3d_ssp_row {
    double wvl_um;      /* rounded wavelength of this row                                                   */
    double ssa_ice;     /* self.ssa_ice[p]                                                                  */
    double ssa_imp;     /* self.ssa_imp[p]                                                                  */
    double g;           /* self.g[p], Henyey-Greenstein asymmetry                                           */
    double ext_cff_mss; /* self.ext_cff_mss[p] = ext_ice (1 - c) + ext_imp c   [m2 kg-1]                    */
    double p_ext_imp;   /* self.P_ext_imp[p]                                                                */
} mc3d_ssp_row;

/* Per-photon outcome records, struct of arrays, index = photon id - photon_begin.  Any pointer may be NULL
 * (that column is then not copied back).  Replaces the list of tuples returned through comm.gather
 * (monte_carlo3D.py:1487-1488, 1618); wvn and snow_depth are table[wvl_row].
 *
 * `packed` (optional) selects the 16-byte packed record instead of the six columns (19 bytes): when it is non-NULL
 * the columns are ignored and photon p's record is the four 32-bit words packed[4p .. 4p+3]
 *     word 0   n_scat << 9 | wvl_row      (n_scat saturates at MC3D_PACKED_NSCAT_MAX; tables of <= 512 rows only)
 *     word 1   float bits of theta_n,     sign bit = condition bit 0
 *     word 2   float bits of phi_n,       sign bit = condition bit 1
 *     word 3   float bits of path_length, sign bit = condition bit 2      (the three floats are never negative)
 * -- one copy of 16 bytes per photon over PCIe; mc3d_unpack_records expands it into columns on the host. */
typedef struct mc3d_records {
    uint8_t *condition;  /* 1..5                                                                           */
    int16_t *wvl_row;    /* row of the SSP table                                                            */
    float *theta_n;      /* arccos(muz), [0, pi]                                                            */
    float *phi_n;        /* [0, 2 pi), 0 for an unscattered photon                                          */
    uint32_t *n_scat;    /* i - 1                                                                           */
    float *path_length;  /* metres inside the slab                                                          */
    uint32_t *packed;    /* NULL, or uint32[4 * n_photon]: the packed form (see above)                      */
} mc3d_records;
#define MC3D_PACKED_MAX_ROWS 512
#define MC3D_PACKED_NSCAT_MAX 0x7fffffu /* a photon with more scatterings is stored with this value and the call's
                                           mc3d_stats.packed_saturated is set                                */

/* Same, fp64, for replay mode (bit-for-bit comparable with the reference's tuples). */
typedef struct mc3d_records_f64 {
    int32_t *condition;
    double *wvn;
    double *theta_n;
    double *phi_n;
    int64_t *n_scat;
    double *path_length;
    double *snow_depth;
    int64_t *consumed;   /* stream values the photon consumed (must equal offsets[p+1] - offsets[p])          */
} mc3d_records_f64;

typedef struct mc3d_stats {
    uint64_t n_photon;    /* photons walked by this context in the call                                     */
    uint64_t n_events;    /* sum over photons of (n_scat + 1) = loop iterations of monte_carlo3D.py:1212      */
    double kernel_ms;     /* CUDA-event time of the walk kernel(s), max over this context's devices         */
    double total_ms;      /* host wall time of the call (uploads + kernel + copy-back + reduce)             */
    int32_t n_devices;
    int32_t sm_count;     /* of device 0 of the context                                                     */
    int32_t sm_clock_khz; /* cudaDevAttrClockRate of device 0                                               */
    int32_t grid_blocks;  /* persistent blocks launched per device                                          */
    int32_t block_threads;
    int32_t packed_saturated; /* 1 if a photon's n_scat exceeded MC3D_PACKED_NSCAT_MAX in a packed-record call */
    int32_t walk_path;    /* MC3D_PATH_FUSED or MC3D_PATH_PERSISTENT: the kernels that ran                   */
    int32_t reserved;
} mc3d_stats;

/* Optional histograms of two per-photon columns, binned on the GPU so that no records have to leave it
 * (post_processing.py:162-223: np.histogram(path_length * 100, bins=1000) and np.histogram(n_scat, bins=200), whose
 * ranges are the data's own (min, max) -- obtained here from a first pass, see mc3d_extrema). */
typedef struct mc3d_hist_spec {
    int32_t n_scat_bins;  /* 0 = no n_scat histogram                                                        */
    int32_t path_bins;    /* 0 = no path-length histogram                                                   */
    double n_scat_lo, n_scat_hi; /* np.histogram range for float64(n_scat)                                  */
    double path_lo, path_hi;     /* np.histogram range for float64(path_length) * path_scale                */
    double path_scale;    /* 100 for centimetres (post_processing.py:169)                                   */
} mc3d_hist_spec;

/* Extrema over the photons of one call (all devices of the context); always computed. */
typedef struct mc3d_extrema {
    uint32_t n_scat_min, n_scat_max;
    float path_min, path_max; /* metres */
} mc3d_extrema;

/* ---- library / device discovery ---------------------------------------------------------------------- */
int mc3d_abi_version(void);
const char *mc3d_last_error(void);
/* Number of CUDA devices and properties of device `device` (any out pointer may be NULL). */
int mc3d_query(int device, int *n_devices, int *sm_count, int *sm_clock_khz, uint64_t *global_mem_bytes,
               int *cc_major, int *cc_minor);

/* ---- contexts (replace `Parallel(...)`'s MPI.COMM_WORLD, parallelize.py:6-12) ------------------------- */
/* Single process driving n_dev devices (n_dev >= 1).  With n_dev > 1 a NCCL communicator is created with
 * ncclCommInitAll; photon-id ranges follow np.array_split (parallelize.py:14-15). */
int mc3d_create(mc3d_ctx **ctx, const int *device_ids, int n_dev);
/* One process per GPU (torchrun / mpirun style): rank 0 calls mc3d_nccl_unique_id, ships the 128 bytes to all
 * ranks by any means, every rank calls mc3d_create_rank.  world_size == 1 needs no id (may be NULL). */
int mc3d_nccl_unique_id(uint8_t id_out[128]);
int mc3d_create_rank(mc3d_ctx **ctx, int device_id, const uint8_t nccl_id[128], int rank, int world_size);
int mc3d_destroy(mc3d_ctx *ctx);

/* ---- pinned host memory for the record arrays (so copy-back runs at PCIe speed) ------------------------ */
int mc3d_host_alloc(void **ptr, uint64_t bytes);
int mc3d_host_free(void *ptr);

/* Packed layout for the record arrays of an n-photon call: byte offsets of the six mc3d_records columns (in struct
 * order, each 256-byte aligned) inside one block of *total_bytes.  Optional: when the six host pointers of a
 * single-device, single-chunk (n <= 2^26) call follow this layout inside one (pinned) block, the library returns the
 * records with ONE device-to-host copy instead of six; any other arrangement of pointers works as before. */
int mc3d_records_layout(uint64_t n_photon, uint64_t offsets[6], uint64_t *total_bytes);

/* ---- the hot path --------------------------------------------------------------------------------------
 * Production mode (fp32 walk, Philox4x32-7 keyed on (seed, photon id)).  Walks photon ids
 * [photon_begin, photon_begin + n_photon) -- replaces the loop monte_carlo3D.py:1613-1616 together with
 * initial_pdfs/populate_pdfs (monte_carlo3D.py:885-921, 1010-1044), Henyey_Greenstein2 (790-800) and the
 * per-photon wavelength draw (1515-1520).  In a multi-device context the range is split with np.array_split
 * boundaries; in a multi-rank context each rank passes its own sub-range.
 *
 *   table, n_rows   per-wavelength SSP rows; photons whose drawn wavelength falls outside are clamped to the
 *                   first / last row (cannot happen for a table covering +-7 sigma: |z| <= 6.8 with 32-bit
 *                   uniforms).
 *   records         NULL, or SoA destination for this call's photons (host memory, ideally mc3d_host_alloc'd).
 *   tally           NULL, or uint64[n_rows * (MC3D_N_COND + n_theta_bins * max(1, n_phi_bins))]: per row, outcome
 *                   counts by condition followed by the reflected-photon zenith histogram (np.histogram semantics
 *                   of post_processing.py:73-76 applied to float64(theta_n)); with n_phi_bins > 1 each zenith bin
 *                   is split into azimuth bins (np.histogram2d semantics, bin = theta_bin * n_phi_bins + phi_bin).  OVERWRITTEN with this call's
 *                   counts, reduced over all devices of the context; in a multi-rank context call
 *                   mc3d_reduce_tally afterwards for the job total.
 */
int mc3d_run(mc3d_ctx *ctx, const mc3d_params *params, const mc3d_ssp_row *table, int n_rows, uint64_t seed,
             uint64_t photon_begin, uint64_t n_photon, const mc3d_records *records, uint64_t *tally,
             mc3d_stats *stats);

/* Asynchronous variant: enqueues upload + walk + copy-back on the context's streams and returns; the record
 * and tally buffers must stay valid (and should be pinned) until mc3d_wait.  `slot` (0 .. MC3D_N_SLOTS-1) selects one
 * of sixteen independent device buffer sets, each with its own stream, so that several calls can be in flight: the
 * long-walk tail and the copy-back of one call overlap the walks of the next ones.  mc3d_wait(ctx, slot, stats)
 * blocks until that slot is complete and returns the call's statistics (the `stats` of mc3d_run_async, if not NULL,
 * is only zeroed: nothing has run yet when it returns). */
int mc3d_run_async(mc3d_ctx *ctx, int slot, const mc3d_params *params, const mc3d_ssp_row *table, int n_rows,
                   uint64_t seed, uint64_t photon_begin, uint64_t n_photon, const mc3d_records *records,
                   uint64_t *tally, mc3d_stats *stats);
int mc3d_wait(mc3d_ctx *ctx, int slot, mc3d_stats *stats);

/* ---- sweeps: many cases, the same launches --------------------------------------------------------------------
 * The reference's driver loops MonteCarlo.run over wavelengths, grain radii and zenith angles (monte_carlo3D-run.py:
 * 60-96, 112-122), one full run each.  mc3d_run_sweep walks all the cases of such a loop together: one concatenated
 * SSP table, one set of launches over the concatenated photons, so that short cases share the GPU with long ones and
 * small cases do not pay a launch (and its tail) each.
 *
 * Case c owns rows table[row_begin .. row_begin + n_rows) and n_photon photons; its photon j is photon id
 * (c << 40) + j of the stream `seed` -- the case index sits in the high bits of the Philox counter -- so the call
 * equals, bit for bit, n_cases calls mc3d_run(&cases[c].params, table + row_begin, n_rows, seed, (uint64_t)c << 40,
 * n_photon, ...) with their records concatenated in case order (wvl_row counts from the case's row_begin) and their
 * tallies stacked by table row: tally is uint64[n_rows_total * (MC3D_N_COND + n_theta_bins * max(1, n_phi_bins))].
 * Cases may share rows only as identical ranges with equal tau_tot and rho_snw (typically: one table, several
 * zenith angles); n_theta_bins / n_phi_bins are common to all cases.  case_events (NULL or uint64[n_cases]) receives
 * the events of each case.  The asynchronous form takes a slot like mc3d_run_async (finish with mc3d_wait) and a
 * sub-range [range_begin, range_begin + range_count) of the concatenated photons, which is how ranks of a multi-rank
 * context share a sweep (records are indexed from range_begin; tallies and case_events cover the sub-range). */
typedef struct mc3d_sweep_case {
    mc3d_params params;
    int32_t row_begin;
    int32_t n_rows;
    uint64_t n_photon;   /* < 2^40 */
} mc3d_sweep_case;
#define MC3D_SWEEP_MAX_CASES 1024
int mc3d_run_sweep(mc3d_ctx *ctx, const mc3d_sweep_case *cases, int n_cases, const mc3d_ssp_row *table, int n_rows_total,
                   uint64_t seed, const mc3d_records *records, uint64_t *tally, uint64_t *case_events, mc3d_stats *stats);
int mc3d_run_sweep_async(mc3d_ctx *ctx, int slot, const mc3d_sweep_case *cases, int n_cases, const mc3d_ssp_row *table,
                         int n_rows_total, uint64_t seed, uint64_t range_begin, uint64_t range_count,
                         const mc3d_records *records, uint64_t *tally, uint64_t *case_events);

/* Sum `tally` (uint64[n]) over the ranks of a multi-rank context with one ncclReduce to `root`
 * (replaces comm.gather for the reduced quantities, parallelize.py:19).  In place; no-op for world_size 1. */
int mc3d_reduce_tally(mc3d_ctx *ctx, uint64_t *tally, uint64_t n, int root);

/* Histograms (see mc3d_hist_spec).  mc3d_set_histograms configures the context: every later mc3d_run /
 * mc3d_run_async also bins its photons with np.histogram's uniform-bin semantics (NULL switches it off again; the
 * spec is copied).  mc3d_get_histograms, called after mc3d_wait / mc3d_run and before the slot is reused, copies
 * out the counts of that call summed over the context's devices -- n_scat_counts: uint64[n_scat_bins],
 * path_counts: uint64[path_bins], either may be NULL -- and the extrema (may be NULL).  In a multi-rank context sum
 * the counts with mc3d_reduce_tally and combine the extrema on the host. */
int mc3d_set_histograms(mc3d_ctx *ctx, const mc3d_hist_spec *spec);
int mc3d_get_histograms(mc3d_ctx *ctx, int slot, uint64_t *n_scat_counts, uint64_t *path_counts, mc3d_extrema *extrema);

/* Replay mode (fp64 walk that consumes the reference's own recorded random stream; correctness tool).
 * Per-photon inputs are the arrays the reference holds at monte_carlo3D.py:1575-1612: wvl, ssa_ice, ssa_imp,
 * g, ext_cff_mss, P_ext_imp (each double[n_photon]); init_draws = the 3 uniforms per photon of initial_pdfs
 * (monte_carlo3D.py:1035-1038); stream/offsets = raw uniforms consumed by photon p's walk in order
 * [r1, u_phi, u_tau, u_ssa, u_ext [, u_refl] [, (u_theta, r)...]]* , stream[offsets[p] .. offsets[p+1]).
 * n_mismatch receives the number of photons whose consumed count differs from the recorded one. */
int mc3d_replay(mc3d_ctx *ctx, const mc3d_params *params, uint64_t n_photon, const double *wvl,
                const double *ssa_ice, const double *ssa_imp, const double *g, const double *ext_cff_mss,
                const double *p_ext_imp, const double *init_draws, const int64_t *offsets, const double *stream,
                const mc3d_records_f64 *out, uint64_t *n_mismatch);

/* ---- native text writer (host code; the step after the path, monte_carlo3D.py:1621-1648) ---------------------
 * Appends one line per photon, '%d %r %r %r %d %r %r\n' % (condition, wvn, theta_n, phi_n, n_scat, path_length,
 * snow_depth) with CPython's float repr (shortest round-trip digits), byte-identical to the reference's writer.
 * Columns as in mc3d_records (float columns are widened to double first); wvn and snow_depth are per-row tables
 * indexed by wvl_row.  append != 0 appends to an existing file (the caller writes the header line first).
 * n_threads <= 0 uses every host core.  Returns the number of bytes written or a negative MC3D_E* code. */
int64_t mc3d_write_records_text(const char *path, int append, uint64_t n, const uint8_t *condition, const int16_t *wvl_row,
                                const float *theta_n, const float *phi_n, const uint32_t *n_scat, const float *path_length,
                                const double *wvn_by_row, const double *snow_depth_by_row, int n_rows, int n_threads);
/* Host code: expand n packed records (mc3d_records.packed layout) into the columns of `out` (any may be NULL;
 * out->packed is ignored).  n_threads <= 0 uses every host core. */
int mc3d_unpack_records(const uint32_t *packed, uint64_t n, const mc3d_records *out, int n_threads);
/* CPython repr(x) of one double into buf (at least 32 bytes, NUL-terminated); returns its length. */
int mc3d_py_repr(double x, char *buf);

/* Tuning knobs (optional): persistent blocks per SM (default: automatic -- enough lanes for >= 26 photons each,
 * capped by what is resident; 255 restores automatic), threads per block (128 / 256 / 512, default 256) and the
 * number of waiting lanes at which a warp resolves / refills (default 4).  0 keeps the current value. */
int mc3d_set_launch(mc3d_ctx *ctx, int blocks_per_sm, int block_threads, int refill_threshold);

/* Which kernels walk the photons.  MC3D_PATH_PERSISTENT: init -> persistent-warp walk (-> tail) -> finalize, per-photon
 * state handed through HBM.  MC3D_PATH_FUSED: one kernel from a photon's first draw to its record, nothing but the
 * record touches HBM (three full-width stages over shared-memory queues; made for walks of a few events: strongly
 * absorbing grains, thin slabs, Lambertian_surface).  MC3D_PATH_AUTO (default) picks by the expected walk length --
 * at present always the persistent path, which measures faster at every walk length.  A performance
 * choice only: both follow the same per-photon random stream and return bit-identical results. */
#define MC3D_PATH_AUTO 0
#define MC3D_PATH_FUSED 1
#define MC3D_PATH_PERSISTENT 2
int mc3d_set_walk_path(mc3d_ctx *ctx, int path);

/* How a persistent-path call ends.  When the last photon has been handed out every lane still carries one, and the
 * call ends with its longest walk (the `while` loop of one photon, monte_carlo3D.py:1212-1466, is a sequential chain).
 * mode 1: the walk kernel hands those photons to a tail kernel (dense warps; once a warp is down to four photons its
 * idle lanes prepare the photons' next events, which halves the time per event of a lone walk).  mode 0: the walk
 * kernel drains by itself, and draining warps consolidate when other calls are in flight.  mode -1 (default): 1 for a
 * synchronous call (mc3d_run, mc3d_run_sweep) that starts with no other call in flight on the context, else 0 (an
 * asynchronous call may get company before it ends, and the tail kernel would queue behind it).  A performance choice only: bit-identical
 * results either way. */
int mc3d_set_tail_kernel(mc3d_ctx *ctx, int mode);

/* Input caching (default on): a call whose SSP table, bin edges and histogram edges equal what its slot uploaded
 * last time skips the host-to-device copy (a few KB).  enabled = 0 makes every call upload its inputs again. */
int mc3d_set_input_caching(mc3d_ctx *ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* MC3D_H_ */
