// mc3d_api.cu -- host runtime and C ABI of libmc3d.so (see include/mc3d.h for the contract).
//
// What the reference does around the walk and what replaces it here:
//   Parallel._map / np.array_split + comm.scatter   (parallelize.py:14-15, 28-38)  -> photon-id ranges per device
//   the Python photon loop                          (monte_carlo3D.py:1613-1616)   -> init_kernel + walk_kernel +
//                                                                                      finalize_kernel
//   comm.gather of per-photon tuples                (parallelize.py:19)            -> each device copies its id
//                                                       range straight into the caller's SoA arrays
//   (no reference equivalent) outcome / BRF tallies                               -> integer tallies, one
//                                                       ncclReduce(sum, uint64) over NVLink when > 1 device/rank
// NCCL is resolved with dlopen at first multi-GPU use, so the library loads (and single-GPU runs work) on hosts
// without NCCL and shares the process' NCCL when the caller already loaded one.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include <nccl.h>

#include "../../include/mc3d.h"
#include "mc3d_device.cuh"

namespace mc3d {
cudaError_t launch_walk(const WalkParams &P, bool impurity, int block_threads, int blocks_per_sm, int grid,
                        cudaStream_t stream, int *occupancy);
cudaError_t launch_init(const WalkParams &P, bool impurity, int sm_count, cudaStream_t stream);
cudaError_t launch_tail(const WalkParams &P, bool impurity, int lanes, cudaStream_t stream);
cudaError_t launch_finalize(const FinalizeParams &P, int sm_count, cudaStream_t stream);
cudaError_t launch_fused(const WalkParams &P, const FinalizeParams &F, bool impurity, int sm_count, cudaStream_t stream);
cudaError_t launch_replay(const ReplayParams &P, cudaStream_t stream);
}  // namespace mc3d

using namespace mc3d;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(MC3D_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.handle) return MC3D_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(MC3D_ENCCL, "NCCL not found (dlopen libnccl.so.2: %s)", dlerror());
#define SYM(field, name)                                                        \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                 \
    if (!g_nccl.field) return fail(MC3D_ENCCL, "NCCL symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Reduce, "ncclReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.handle = h;
    return MC3D_OK;
}
}  // namespace

#define NCCL_TRY(expr)                                                                                       \
    do {                                                                                                       \
        ncclResult_t r_ = (expr);                                                                              \
        if (r_ != ncclSuccess)                                                                                 \
            return fail(MC3D_ENCCL, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------ context
namespace {

constexpr uint64_t CHUNK_PHOTONS = 1ull << 26;   // raw results: 32 B/photon -> 2 GiB per chunk buffer
constexpr int N_SLOTS = MC3D_N_SLOTS;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Slot {
    // read-only inputs of a call in one device block (one upload, skipped when nothing changed since the slot's last
    // call): DevRow[n_rows] | zenith + azimuth edges | column-histogram edges
    DevBuf<uint8_t> cst;
    uint8_t *host_cst = nullptr;            // pinned staging
    size_t host_cst_cap = 0;
    std::vector<uint8_t> cst_shadow;        // what the device block holds
    // accumulators in one device block (one memset, one copy back), in 64-bit words:
    // tally[tally_len] | n_events | extrema (2 words) + column histograms | claim counters (4 x uint32 per chunk: ring claim counter, spare, tail-list length, spare)
    DevBuf<unsigned long long> acc;
    unsigned long long *host_acc = nullptr; // pinned mirror of tally .. column histograms
    size_t host_acc_cap = 0, extras_len = 0, n_case_ev = 0;
    DevBuf<Fresh> fresh;                    // photons that survived their first event (init kernel -> walk kernel)
    DevBuf<RawResult> raw;
    DevBuf<uint32_t> tail;                  // walk kernel -> tail kernel hand-over list (calls that run alone)
    DevBuf<uint8_t> recs;                   // record columns of one chunk, packed like mc3d_records_layout
    cudaStream_t stream = nullptr;          // each slot has its own stream: two calls in flight overlap on the GPU
    std::vector<cudaEvent_t> ev;            // pairs (begin, end) around walk+finalize of each chunk
    // pending call
    bool busy = false;
    uint64_t n_photon = 0;
    size_t tally_len = 0;
    int n_chunks = 0;
    uint64_t *user_tally = nullptr;
    uint64_t *user_case_ev = nullptr;       // sweeps: events per case
    bool packed = false;                    // the pending call returns packed records
};

struct Device {
    int id = 0;
    int sm_count = 0, clock_khz = 0;
    Slot slot[N_SLOTS];
};

}  // namespace

struct mc3d_ctx {
    std::vector<Device> devs;
    std::vector<ncclComm_t> comms;   // one per device (single process) or one (multi rank)
    int rank = 0, world = 1;
    int blocks_per_sm = 0;   // 0 = automatic: enough lanes for >= 26 photons each, at most the resident capacity
    int block_threads = 256, refill_threshold = 4;
    struct Occupancy { bool impurity; int block_threads, bps_variant, n_rows; uint32_t n_cases; int resident; };
    std::vector<Occupancy> occupancy;   // resident walk-kernel blocks per SM, queried once per variant
    bool input_caching = true; // skip the upload of inputs identical to the slot's previous call
    int drain_give = -1;       // -1 = automatic (16 when other calls are in flight, else off); MC3D_DRAIN_GIVE overrides
    int tail_kernel = -1;      // -1 = automatic (a synchronous call that runs alone finishes in the tail kernel); MC3D_TAIL overrides
    bool sync_call = false;    // set by mc3d_run / mc3d_run_sweep around their enqueue
    int per_lane = 0;          // photons per lane the automatic grid aims for (0 = 26, or 13 for a call that runs alone); MC3D_PER_LANE
    int drain_latency = -1;    // -1 = automatic (on for a call that runs alone); MC3D_DRAIN_LATENCY overrides
    int walk_path = MC3D_PATH_AUTO;   // mc3d_set_walk_path / MC3D_WALK_PATH
    double fused_max_events = 0.0;    // automatic path: fused kernel when a photon is expected to end within this many events.
                                      // 0 = never: since the init kernel's fixes the persistent path is faster at every walk
                                      // length measured, 0.30 vs 0.35 ms per 1e7 photons at 2.1 events each
                                      // (profiles/r02_fused_vs_persistent.log); MC3D_FUSED_MAX_EVENTS / mc3d_set_walk_path
    std::chrono::steady_clock::time_point t0[N_SLOTS];
    mc3d_stats pending_stats[N_SLOTS];
    bool hist_on = false;
    mc3d_hist_spec hist_spec{};
    // results of the last completed call of each slot (combined over devices in mc3d_wait)
    mc3d_hist_spec done_spec[N_SLOTS]{};
    bool done_hist[N_SLOTS] = {};
    mc3d_extrema done_extrema[N_SLOTS]{};
    std::vector<uint64_t> done_counts[N_SLOTS];
    DevBuf<unsigned long long> reduce_buf;   // mc3d_reduce_tally's device staging (multi-rank contexts)
};

// ------------------------------------------------------------------------------------------------ helpers
static void array_split(uint64_t n, int parts, int k, uint64_t *begin, uint64_t *count)
{
    // np.array_split boundaries (parallelize.py:14-15): the first n % parts chunks get one extra element
    const uint64_t q = n / parts, r = n % parts;
    *begin = (uint64_t)k * q + std::min<uint64_t>(k, r);
    *count = q + ((uint64_t)k < r ? 1 : 0);
}

static void threshold40(double ssa, uint32_t *t16, uint32_t *t24)
{
    // absorbed iff (K40 + 1/2) 2^-40 >= ssa  <=>  K40 >= T40 = ceil(ssa 2^40 - 1/2); t16 = T40 >> 24 (0x10000 when
    // T40 == 2^40: never absorbed), t24 = T40 & 0xffffff
    double t = std::ceil(std::ldexp(ssa, 40) - 0.5);
    if (!(t > 0.0)) t = 0.0;   // also catches NaN
    if (t >= 1099511627776.0) t = 1099511627776.0;
    const uint64_t T = (uint64_t)t;
    *t16 = (uint32_t)(T >> 24);
    *t24 = (uint32_t)(T & 0xffffffu);
}

static bool build_rows(const mc3d_params *P, const mc3d_ssp_row *table, int n_rows, float neg_tau_tot, DevRow *out)
{
    bool impurity = false;
    for (int r = 0; r < n_rows; ++r) {
        const mc3d_ssp_row &s = table[r];
        DevRow &d = out[r];
        d.one_m_g = (float)(1.0 - s.g);
        d.one_m_g2 = (float)(1.0 - s.g * s.g);
        d.d_scale = (float)std::ldexp(2.0 * s.g, -32);
        d.d_off = (float)(1.0 - s.g + std::ldexp(s.g, -32));
        if (s.g == 0.0) { d.omr_scale = (float)std::ldexp(1.0, -32); d.omr_off = (float)std::ldexp(1.0, -33); }
        else { d.omr_scale = -(float)std::ldexp(1.0, -32); d.omr_off = 1.0f; }
        threshold40(s.ssa_ice, &d.t16, &d.t24);
        threshold40(s.ssa_imp, &d.ti16, &d.ti24);
        d.t_hot = std::min(d.t16, RENORM_KEY) << 16;
        d.ti_hot = std::min(d.ti16, RENORM_KEY) << 16;
        d.neg_tau_tot = neg_tau_tot;
        // impurity iff (w + 1/2) 2^-32 <= P_ext_imp  <=>  w <= floor(P 2^32 - 1/2)
        const double sl = std::floor(std::ldexp(s.p_ext_imp, 32) - 0.5);
        if (sl >= 0.0) {
            d.s_any = 1u;
            d.s_last = sl >= 4294967295.0 ? 0xffffffffu : (uint32_t)sl;
            impurity = true;
        } else {
            d.s_any = 0u;
            d.s_last = 0u;
        }
        d.inv_ext = (float)(0.6931471805599453 / (s.ext_cff_mss * P->rho_snw));
    }
    return impurity;
}

// Rough number of events a photon of the centre wavelength lives (only used to pick the kernel: a performance
// choice, the results do not depend on it).  It ends by absorption after ~1 / (1 - ssa) events, or by leaving a slab
// of optical depth tau through a face after ~(1 + tau)(1 + tau (1 - g)) events (a reflecting bottom sends it back).
static double expected_events(const mc3d_params *P, const mc3d_ssp_row *table, int n_rows)
{
    if (P->flags & MC3D_FLAG_LAMBERT_SURFACE) return 2.0;
    const int r = std::max(0, std::min(n_rows - 1, (int)std::lrint(P->wvl0_um * 100.0) - P->k_first));
    const mc3d_ssp_row &s = table[r];
    const double a = (1.0 - s.p_ext_imp) * (1.0 - s.ssa_ice) + s.p_ext_imp * (1.0 - s.ssa_imp);
    const double by_absorption = 1.0 / std::max(a, 1e-12);
    double by_escape = (1.0 + P->tau_tot) * (1.0 + P->tau_tot * (1.0 - s.g));
    if (P->flags & MC3D_FLAG_LAMBERT_BOTTOM) by_escape *= 1.0 + 2.0 * std::max(0.0, std::min(1.0, P->r_lambert));
    return std::min(by_absorption, by_escape);
}

// np.linspace(start, stop, n + 1): arange * step + start with the endpoint forced (numpy/_core/function_base.py)
static void linspace_edges(double start, double stop, int n_bins, double *out)
{
    const double step = (stop - start) / n_bins;
    for (int b = 0; b <= n_bins; ++b) out[b] = b * step + start;
    out[n_bins] = stop;
}

// mc3d_records_layout: columns in struct order, each 256-byte aligned
static const size_t REC_ITEM[6] = {sizeof(uint8_t), sizeof(int16_t), sizeof(float), sizeof(float), sizeof(uint32_t), sizeof(float)};
static size_t records_layout(uint64_t n, size_t off[6])
{
    size_t at = 0;
    for (int c = 0; c < 6; ++c) {
        off[c] = at;
        at += ((size_t)n * REC_ITEM[c] + 255) & ~(size_t)255;
    }
    return at;
}

static void philox_round_keys(uint64_t seed, uint32_t rk[2 * PHILOX_ROUNDS])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        rk[2 * r] = k0;
        rk[2 * r + 1] = k1;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
}

static int check_ctx(mc3d_ctx *ctx)
{
    if (!ctx || ctx->devs.empty()) return fail(MC3D_EINVAL, "null or empty context");
    return MC3D_OK;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int mc3d_abi_version(void) { return MC3D_ABI_VERSION; }

const char *mc3d_last_error(void) { return g_err; }

int mc3d_query(int device, int *n_devices, int *sm_count, int *sm_clock_khz, uint64_t *global_mem_bytes,
               int *cc_major, int *cc_minor)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        if (n_devices) *n_devices = 0;
        return fail(MC3D_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    if (n_devices) *n_devices = n;
    if (device < 0 || device >= n) return fail(MC3D_EINVAL, "device %d out of range [0, %d)", device, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    int clock = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&clock, cudaDevAttrClockRate, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (sm_clock_khz) *sm_clock_khz = clock;
    if (global_mem_bytes) *global_mem_bytes = (uint64_t)prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return MC3D_OK;
}

static int init_device(Device &d, int id)
{
    d.id = id;
    CUDA_TRY(cudaSetDevice(id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, id));
    if (prop.major < 10)
        return fail(MC3D_ENODEVICE, "device %d is sm_%d%d; libmc3d is built for sm_100a (B200) only", id, prop.major,
                    prop.minor);
    d.sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaDeviceGetAttribute(&d.clock_khz, cudaDevAttrClockRate, id));
    for (Slot &s : d.slot) CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    return MC3D_OK;
}

static void apply_env(mc3d_ctx *ctx)
{
    const char *e = getenv("MC3D_DRAIN_GIVE");        // experiments only; results do not depend on it
    if (e && *e) ctx->drain_give = std::max(0, std::min(31, atoi(e)));
    e = getenv("MC3D_PER_LANE");                      // experiments only
    if (e && *e) ctx->per_lane = std::max(0, atoi(e));
    e = getenv("MC3D_TAIL");
    if (e && *e) ctx->tail_kernel = atoi(e) ? 1 : 0;
    e = getenv("MC3D_DRAIN_LATENCY");
    if (e && *e) ctx->drain_latency = atoi(e) ? 1 : 0;
    e = getenv("MC3D_WALK_PATH");                     // same: "fused" | "persistent" | "auto"
    if (e && !strcmp(e, "fused")) ctx->walk_path = MC3D_PATH_FUSED;
    if (e && !strcmp(e, "persistent")) ctx->walk_path = MC3D_PATH_PERSISTENT;
    e = getenv("MC3D_FUSED_MAX_EVENTS");
    if (e && *e) ctx->fused_max_events = atof(e);
}

int mc3d_create(mc3d_ctx **out, const int *device_ids, int n_dev)
{
    if (!out) return fail(MC3D_EINVAL, "ctx out pointer is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(MC3D_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    if (n_dev < 1 || n_dev > n) return fail(MC3D_EINVAL, "n_dev %d out of range [1, %d]", n_dev, n);
    mc3d_ctx *ctx = new (std::nothrow) mc3d_ctx();
    if (!ctx) return fail(MC3D_ENOMEM, "out of host memory");
    ctx->devs.resize(n_dev);
    std::vector<int> ids(n_dev);
    for (int k = 0; k < n_dev; ++k) {
        ids[k] = device_ids ? device_ids[k] : k;
        if (ids[k] < 0 || ids[k] >= n) { delete ctx; return fail(MC3D_EINVAL, "device id %d out of range", ids[k]); }
        int rc = init_device(ctx->devs[k], ids[k]);
        if (rc) { mc3d_destroy(ctx); return rc; }
    }
    if (n_dev > 1) {
        int rc = load_nccl();
        if (rc) { mc3d_destroy(ctx); return rc; }
        ctx->comms.assign(n_dev, nullptr);
        ncclResult_t r = g_nccl.CommInitAll(ctx->comms.data(), n_dev, ids.data());
        if (r != ncclSuccess) {
            ctx->comms.clear();
            mc3d_destroy(ctx);
            return fail(MC3D_ENCCL, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
        }
    }
    apply_env(ctx);
    *out = ctx;
    return MC3D_OK;
}

int mc3d_nccl_unique_id(uint8_t id_out[128])
{
    if (!id_out) return fail(MC3D_EINVAL, "id_out is null");
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    return MC3D_OK;
}

int mc3d_create_rank(mc3d_ctx **out, int device_id, const uint8_t nccl_id[128], int rank, int world_size)
{
    if (!out) return fail(MC3D_EINVAL, "ctx out pointer is null");
    *out = nullptr;
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(MC3D_EINVAL, "bad rank %d / world %d", rank, world_size);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(MC3D_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n) return fail(MC3D_EINVAL, "device id %d out of range", device_id);
    mc3d_ctx *ctx = new (std::nothrow) mc3d_ctx();
    if (!ctx) return fail(MC3D_ENOMEM, "out of host memory");
    ctx->devs.resize(1);
    ctx->rank = rank;
    ctx->world = world_size;
    int rc = init_device(ctx->devs[0], device_id);
    if (rc) { mc3d_destroy(ctx); return rc; }
    if (world_size > 1) {
        if (!nccl_id) { mc3d_destroy(ctx); return fail(MC3D_EINVAL, "nccl_id is required for world_size > 1"); }
        rc = load_nccl();
        if (rc) { mc3d_destroy(ctx); return rc; }
        ncclUniqueId id;
        memcpy(&id, nccl_id, 128);
        ctx->comms.assign(1, nullptr);
        ncclResult_t r = g_nccl.CommInitRank(&ctx->comms[0], world_size, id, rank);
        if (r != ncclSuccess) {
            ctx->comms.clear();
            mc3d_destroy(ctx);
            return fail(MC3D_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        }
    }
    apply_env(ctx);
    *out = ctx;
    return MC3D_OK;
}

int mc3d_destroy(mc3d_ctx *ctx)
{
    if (!ctx) return MC3D_OK;
    for (size_t k = 0; k < ctx->comms.size(); ++k)
        if (ctx->comms[k] && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comms[k]);
    if (ctx->reduce_buf.p && !ctx->devs.empty() && cudaSetDevice(ctx->devs[0].id) == cudaSuccess) ctx->reduce_buf.release();
    for (Device &d : ctx->devs) {
        if (cudaSetDevice(d.id) != cudaSuccess) continue;
        for (Slot &s : d.slot) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.cst.release(); s.acc.release(); s.fresh.release(); s.raw.release(); s.tail.release(); s.recs.release();
            if (s.host_cst) cudaFreeHost(s.host_cst);
            if (s.host_acc) cudaFreeHost(s.host_acc);
            for (cudaEvent_t e : s.ev) cudaEventDestroy(e);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
    }
    delete ctx;
    return MC3D_OK;
}

int mc3d_host_alloc(void **ptr, uint64_t bytes)
{
    if (!ptr) return fail(MC3D_EINVAL, "ptr is null");
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(MC3D_ENOMEM, "cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return MC3D_OK;
}

int mc3d_host_free(void *ptr)
{
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return MC3D_OK;
}

int mc3d_set_input_caching(mc3d_ctx *ctx, int enabled)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    ctx->input_caching = enabled != 0;
    return MC3D_OK;
}

int mc3d_records_layout(uint64_t n_photon, uint64_t offsets[6], uint64_t *total_bytes)
{
    if (!offsets || !total_bytes) return fail(MC3D_EINVAL, "null argument");
    size_t off[6];
    *total_bytes = records_layout(n_photon, off);
    for (int c = 0; c < 6; ++c) offsets[c] = off[c];
    return MC3D_OK;
}

int mc3d_set_launch(mc3d_ctx *ctx, int blocks_per_sm, int block_threads, int refill_threshold)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (block_threads != 0 && block_threads != 128 && block_threads != 256 && block_threads != 512)
        return fail(MC3D_EINVAL, "block_threads must be 128, 256 or 512");
    if (blocks_per_sm < 0 || (blocks_per_sm > 16 && blocks_per_sm != 255)) return fail(MC3D_EINVAL, "blocks_per_sm out of range");
    if (refill_threshold < 0 || refill_threshold > 32) return fail(MC3D_EINVAL, "refill_threshold must be in [1, 32]");
    if (blocks_per_sm) ctx->blocks_per_sm = blocks_per_sm == 255 ? 0 : blocks_per_sm;
    if (block_threads) ctx->block_threads = block_threads;
    if (refill_threshold) ctx->refill_threshold = refill_threshold;
    return MC3D_OK;
}

int mc3d_set_walk_path(mc3d_ctx *ctx, int path)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (path != MC3D_PATH_AUTO && path != MC3D_PATH_FUSED && path != MC3D_PATH_PERSISTENT)
        return fail(MC3D_EINVAL, "path must be MC3D_PATH_AUTO, MC3D_PATH_FUSED or MC3D_PATH_PERSISTENT");
    ctx->walk_path = path;
    return MC3D_OK;
}

int mc3d_set_tail_kernel(mc3d_ctx *ctx, int mode)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (mode < -1 || mode > 1) return fail(MC3D_EINVAL, "mode must be -1 (automatic), 0 or 1");
    ctx->tail_kernel = mode;
    return MC3D_OK;
}

static int validate_run(const mc3d_params *P, const mc3d_ssp_row *table, int n_rows)
{
    if (!P || !table) return fail(MC3D_EINVAL, "params / table is null");
    if (n_rows < 1 || n_rows > 2048) return fail(MC3D_EINVAL, "n_rows %d out of range [1, 2048] (rows are staged in shared memory)", n_rows);
    if (P->n_theta_bins < 0 || P->n_theta_bins > 65536) return fail(MC3D_EINVAL, "n_theta_bins out of range");
    if (P->n_phi_bins < 0 || P->n_phi_bins > 4096) return fail(MC3D_EINVAL, "n_phi_bins out of range");
    if ((uint64_t)n_rows * (N_COND + (uint64_t)P->n_theta_bins * std::max(1, P->n_phi_bins)) > (1ull << 28))
        return fail(MC3D_EINVAL, "tally block too large (rows x zenith x azimuth bins)");
    if (!(P->tau_tot > 0.0)) return fail(MC3D_EINVAL, "tau_tot must be positive");
    if (!(P->rho_snw > 0.0)) return fail(MC3D_EINVAL, "rho_snw must be positive");
    if (!(P->theta0_rad >= 0.0 && P->theta0_rad < 1.5707963267948966))
        return fail(MC3D_EINVAL, "theta0 must be in [0, pi/2)");
    if (!(P->sigma_um >= 0.0)) return fail(MC3D_EINVAL, "sigma must be >= 0");
    for (int r = 0; r < n_rows; ++r) {
        if (!(table[r].ext_cff_mss > 0.0)) return fail(MC3D_EINVAL, "row %d: ext_cff_mss must be positive", r);
        if (!(table[r].g > -1.0 && table[r].g < 1.0)) return fail(MC3D_EINVAL, "row %d: g must be in (-1, 1)", r);
    }
    return MC3D_OK;
}

// One call: one case (mc3d_run / mc3d_run_async) or a sweep of cases walked by the same launches (mc3d_run_sweep*).
namespace {
struct HostCase {
    mc3d_params params;
    int row_begin, n_rows;   // the case's rows in the call's table
    uint64_t id_begin;       // global photon id of the case's first photon
    uint64_t first;          // its position in the call's concatenated photon space
    uint64_t n_photon;
};
struct Job {
    std::vector<HostCase> cases;
    bool sweep = false;
    const mc3d_ssp_row *table = nullptr;
    int n_rows = 0;                       // rows of the whole table
    uint64_t seed = 0;
    uint64_t range_begin = 0, range_count = 0;   // the part of the concatenated photon space this context walks
    const mc3d_records *rec = nullptr;    // indexed from range_begin
    uint64_t *tally = nullptr;
    uint64_t *case_events = nullptr;      // [cases.size()] or null (sweeps)
    int n_theta_bins = 0, n_phi_bins = 0;
    bool sync = false;                    // issued by mc3d_run / mc3d_run_sweep: nothing else can be enqueued before it ends
};

DevCase make_devcase(const mc3d_params *P, int row_begin, int n_rows)
{
    DevCase c;
    memset(&c, 0, sizeof c);
    c.row_begin = (uint32_t)row_begin;
    c.n_rows = n_rows;
    c.mu0x = (float)std::sin(P->theta0_rad);
    if (c.mu0x == 0.0f) c.mu0x = -1e-12f;   // vertical incidence: see apply_event (reproduces the muz_0 == -1 branch)
    c.mu0z = (float)(-std::cos(P->theta0_rad));
    c.tau_tot = (float)(P->tau_tot / 0.6931471805599453);   // the walk's depth unit is ln 2 optical depths
    c.neg_tau_tot = -c.tau_tot;
    c.wvl0_x100 = P->wvl0_um * 100.0;
    c.sigma_x100 = P->sigma_um * 100.0;
    c.k_first = P->k_first;
    const double t = std::floor(std::ldexp(P->r_lambert, 32) - 0.5);
    c.refl_thr = t < -1.0 ? -1 : (t > 4294967295.0 ? 4294967295ll : (long long)t);
    c.lambert_bottom = (P->flags & MC3D_FLAG_LAMBERT_BOTTOM) ? 1u : 0u;
    c.lambert_surface = (P->flags & MC3D_FLAG_LAMBERT_SURFACE) ? 1u : 0u;
    threshold40(P->r_lambert, &c.surf_t16, &c.surf_t24);
    return c;
}
}  // namespace

static int run_job(mc3d_ctx *ctx, int slot_idx, const Job &J);

static int start_job(mc3d_ctx *ctx, int slot_idx, const Job &J)
{
    for (Device &d : ctx->devs)
        if (d.slot[slot_idx].busy) return fail(MC3D_EINVAL, "slot %d is busy; call mc3d_wait first", slot_idx);
    int rc = run_job(ctx, slot_idx, J);
    if (rc) {
        // a failure half way (allocation, launch, NCCL): drain whatever was enqueued and give the slot back, so the
        // context stays usable and no copy into the caller's buffers is left in flight; the error text is kept
        for (Device &d : ctx->devs) {
            if (cudaSetDevice(d.id) == cudaSuccess && d.slot[slot_idx].stream) cudaStreamSynchronize(d.slot[slot_idx].stream);
            d.slot[slot_idx].busy = false;
        }
        (void)cudaGetLastError();
    }
    return rc;
}

int mc3d_run_async(mc3d_ctx *ctx, int slot_idx, const mc3d_params *P, const mc3d_ssp_row *table, int n_rows,
                   uint64_t seed, uint64_t photon_begin, uint64_t n_photon, const mc3d_records *rec,
                   uint64_t *tally, mc3d_stats *stats)
{
    if (stats) memset(stats, 0, sizeof *stats);   // filled by mc3d_wait (the call has not run yet)
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (slot_idx < 0 || slot_idx >= N_SLOTS) return fail(MC3D_EINVAL, "slot must be in [0, %d)", N_SLOTS);
    rc = validate_run(P, table, n_rows);
    if (rc) return rc;
    Job J;
    J.cases.push_back(HostCase{*P, 0, n_rows, photon_begin, 0, n_photon});
    J.table = table; J.n_rows = n_rows; J.seed = seed;
    J.range_begin = 0; J.range_count = n_photon;
    J.rec = rec; J.tally = tally;
    J.n_theta_bins = P->n_theta_bins; J.n_phi_bins = P->n_phi_bins;
    J.sync = ctx->sync_call;
    return start_job(ctx, slot_idx, J);
}

int mc3d_run_sweep_async(mc3d_ctx *ctx, int slot_idx, const mc3d_sweep_case *cases, int n_cases, const mc3d_ssp_row *table,
                         int n_rows_total, uint64_t seed, uint64_t range_begin, uint64_t range_count,
                         const mc3d_records *rec, uint64_t *tally, uint64_t *case_events)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (slot_idx < 0 || slot_idx >= N_SLOTS) return fail(MC3D_EINVAL, "slot must be in [0, %d)", N_SLOTS);
    if (!cases || n_cases < 1 || n_cases > MC3D_SWEEP_MAX_CASES)
        return fail(MC3D_EINVAL, "n_cases must be in [1, %d]", MC3D_SWEEP_MAX_CASES);
    if (!table || n_rows_total < 1 || n_rows_total > 2048)
        return fail(MC3D_EINVAL, "n_rows_total %d out of range [1, 2048] (rows are staged in shared memory)", n_rows_total);
    Job J;
    J.sweep = true;
    J.table = table; J.n_rows = n_rows_total; J.seed = seed;
    J.rec = rec; J.tally = tally; J.case_events = case_events;
    J.n_theta_bins = cases[0].params.n_theta_bins; J.n_phi_bins = cases[0].params.n_phi_bins;
    uint64_t total = 0;
    for (int c = 0; c < n_cases; ++c) {
        const mc3d_sweep_case &s = cases[c];
        if (s.row_begin < 0 || s.n_rows < 1 || (int64_t)s.row_begin + s.n_rows > n_rows_total)
            return fail(MC3D_EINVAL, "case %d: rows [%d, %d) outside the table of %d rows", c, s.row_begin, s.row_begin + s.n_rows, n_rows_total);
        rc = validate_run(&s.params, table + s.row_begin, s.n_rows);
        if (rc) return rc;
        if (s.params.n_theta_bins != J.n_theta_bins || s.params.n_phi_bins != J.n_phi_bins)
            return fail(MC3D_EINVAL, "case %d: all cases of a sweep share n_theta_bins / n_phi_bins", c);
        if (s.n_photon >= (1ull << 40)) return fail(MC3D_EINVAL, "case %d: n_photon must be < 2^40", c);
        // a row carries its case's slab depth and density: cases may share rows only if those agree
        for (int o = 0; o < c; ++o) {
            const mc3d_sweep_case &t = cases[o];
            const bool disjoint = t.row_begin + t.n_rows <= s.row_begin || s.row_begin + s.n_rows <= t.row_begin;
            const bool same = t.row_begin == s.row_begin && t.n_rows == s.n_rows;
            if (!disjoint && !(same && t.params.tau_tot == s.params.tau_tot && t.params.rho_snw == s.params.rho_snw))
                return fail(MC3D_EINVAL, "cases %d and %d: cases share rows only as identical ranges with equal tau_tot and rho_snw", o, c);
        }
        J.cases.push_back(HostCase{s.params, s.row_begin, s.n_rows, (uint64_t)c << 40, total, s.n_photon});
        total += s.n_photon;
    }
    if (range_begin > total || range_count > total - range_begin)
        return fail(MC3D_EINVAL, "photon range [%llu, +%llu) outside the sweep's %llu photons", (unsigned long long)range_begin,
                    (unsigned long long)range_count, (unsigned long long)total);
    J.range_begin = range_begin; J.range_count = range_count;
    J.sync = ctx->sync_call;
    return start_job(ctx, slot_idx, J);
}

int mc3d_run_sweep(mc3d_ctx *ctx, const mc3d_sweep_case *cases, int n_cases, const mc3d_ssp_row *table, int n_rows_total,
                   uint64_t seed, const mc3d_records *rec, uint64_t *tally, uint64_t *case_events, mc3d_stats *stats)
{
    if (stats) memset(stats, 0, sizeof *stats);
    uint64_t total = 0;
    for (int c = 0; cases && c < n_cases; ++c) total += cases[c].n_photon;
    if (check_ctx(ctx) == MC3D_OK) ctx->sync_call = true;
    int rc = mc3d_run_sweep_async(ctx, 0, cases, n_cases, table, n_rows_total, seed, 0, total, rec, tally, case_events);
    if (check_ctx(ctx) == MC3D_OK) ctx->sync_call = false;
    if (rc) return rc;
    return mc3d_wait(ctx, 0, stats);
}

static int run_job(mc3d_ctx *ctx, int slot_idx, const Job &J)
{
    ctx->t0[slot_idx] = std::chrono::steady_clock::now();

    const int n_dev = (int)ctx->devs.size();
    const int n_rows = J.n_rows;
    const int n_phi = std::max(1, J.n_phi_bins);
    const size_t stride = N_COND + (size_t)J.n_theta_bins * n_phi;
    const size_t tally_len = (size_t)n_rows * stride;
    const size_t n_edges = (size_t)J.n_theta_bins + 1 + (size_t)n_phi + 1;
    const size_t n_case_ev = (J.sweep && J.case_events) ? J.cases.size() : 0;

    WalkParams W;
    memset(&W, 0, sizeof W);
    philox_round_keys(J.seed, W.rk);
    W.n_rows = n_rows;
    W.refill_threshold = (uint32_t)ctx->refill_threshold;
    bool lone = true;
    {   // drain-phase consolidation pays when other launches can use the issue slots it frees, i.e. when other
        // calls are in flight on this context; a call running alone is bound by the dependent chain of its longest
        // walks instead and takes the latency-oriented drain loop
        bool others_busy = false;
        for (int s = 0; s < N_SLOTS; ++s) others_busy |= (s != slot_idx && ctx->devs[0].slot[s].busy);
        W.drain_give = ctx->drain_give >= 0 ? (uint32_t)ctx->drain_give : (others_busy ? 16u : 0u);
        W.drain_latency = ctx->drain_latency >= 0 ? (uint32_t)ctx->drain_latency : (others_busy ? 0u : 1u);
        lone = !others_busy;
    }
    double longest = 0.0;   // expected events per photon of the longest-lived case
    for (const HostCase &c : J.cases) longest = std::max(longest, expected_events(&c.params, J.table + c.row_begin, c.n_rows));
    const bool fused = ctx->walk_path == MC3D_PATH_FUSED || (ctx->walk_path == MC3D_PATH_AUTO && longest <= ctx->fused_max_events);
    W.claim = longest <= 4.0 ? 96u : (longest <= 16.0 ? 64u : 32u);   // fewer, larger claims of the fresh list when walks are short

    ctx->done_hist[slot_idx] = ctx->hist_on;
    ctx->done_spec[slot_idx] = ctx->hist_spec;

    mc3d_stats &st = ctx->pending_stats[slot_idx];
    memset(&st, 0, sizeof st);
    st.n_devices = n_dev;
    st.sm_count = ctx->devs[0].sm_count;
    st.sm_clock_khz = ctx->devs[0].clock_khz;
    st.block_threads = ctx->block_threads;
    st.n_photon = J.range_count;
    st.walk_path = fused ? MC3D_PATH_FUSED : MC3D_PATH_PERSISTENT;

    for (int k = 0; k < n_dev; ++k) {
        Device &d = ctx->devs[k];
        Slot &s = d.slot[slot_idx];
        uint64_t off, cnt;
        array_split(J.range_count, n_dev, k, &off, &cnt);   // this device's part of the range, relative to range_begin
        CUDA_TRY(cudaSetDevice(d.id));
        // Chunks of <= 2^26 photons.  A single-case launch never crosses a multiple of 2^32 in the global photon id
        // (its lanes carry only the low id word; the high word is a launch constant); sweep lanes carry both words.
        std::vector<std::pair<uint64_t, uint32_t>> chunks;   // (offset in this device's range, count)
        for (uint64_t c_off = 0; c_off < cnt;) {
            uint64_t c = std::min<uint64_t>(CHUNK_PHOTONS, cnt - c_off);
            if (!J.sweep) {
                const uint64_t gid = J.cases[0].id_begin + J.range_begin + off + c_off;
                c = std::min<uint64_t>(c, 0x100000000ull - (gid & 0xffffffffull));
            }
            chunks.emplace_back(c_off, (uint32_t)c);
            c_off += c;
        }
        const int n_chunks = (int)chunks.size();
        const uint64_t chunk_cap = std::min<uint64_t>(cnt, CHUNK_PHOTONS);
        // sweep: the cases overlapping each chunk, as consecutive case indices [lo, hi]
        std::vector<std::pair<int, int>> chunk_cases(n_chunks, std::make_pair(0, 0));
        size_t n_case_entries = 0;
        if (J.sweep) {
            for (int c = 0; c < n_chunks; ++c) {
                const uint64_t g0 = J.range_begin + off + chunks[c].first, g1 = g0 + chunks[c].second;   // [g0, g1)
                int lo = 0, hi = (int)J.cases.size() - 1;
                while (lo < hi && J.cases[lo].first + J.cases[lo].n_photon <= g0) ++lo;
                while (hi > lo && J.cases[hi].first >= g1) --hi;
                chunk_cases[c] = std::make_pair(lo, hi);
                n_case_entries += (size_t)(hi - lo + 1);
            }
        }

        // ---- buffers
        const bool hist_on = ctx->hist_on;
        const mc3d_hist_spec hs = ctx->hist_spec;
        const size_t n_hist = hist_on ? (size_t)hs.n_scat_bins + (size_t)hs.path_bins : 0;
        const size_t n_hist_edges = hist_on ? n_hist + 2 : 0;
        const size_t rows_bytes = (size_t)n_rows * sizeof(DevRow);
        const size_t edges_bytes = (n_edges + n_hist_edges) * sizeof(double);
        const size_t cst_bytes = rows_bytes + edges_bytes + n_case_entries * sizeof(DevCase);
        s.extras_len = 2 + n_hist;
        s.n_case_ev = n_case_ev;
        const size_t acc_copy = tally_len + 1 + s.extras_len + n_case_ev;          // words copied back
        const size_t acc_words = acc_copy + 2 * (size_t)std::max(n_chunks, 1);    // + two words (4 counters) per chunk
        const bool cst_moved = s.cst.cap < cst_bytes;
        CUDA_TRY(s.cst.ensure(cst_bytes));
        CUDA_TRY(s.acc.ensure(acc_words));
        if (!fused) {   // the fused kernel keeps a photon in registers from its first draw to its record
            CUDA_TRY(s.fresh.ensure(std::max<uint64_t>(chunk_cap, 1)));
            CUDA_TRY(s.raw.ensure(std::max<uint64_t>(chunk_cap, 1)));
        }
        if (s.host_cst_cap < cst_bytes) {
            if (s.host_cst) cudaFreeHost(s.host_cst);
            s.host_cst = nullptr;
            CUDA_TRY(cudaHostAlloc((void **)&s.host_cst, cst_bytes, cudaHostAllocPortable));
            s.host_cst_cap = cst_bytes;
        }
        if (s.host_acc_cap < acc_copy) {
            if (s.host_acc) cudaFreeHost(s.host_acc);
            s.host_acc = nullptr;
            CUDA_TRY(cudaHostAlloc((void **)&s.host_acc, acc_copy * sizeof(unsigned long long), cudaHostAllocPortable));
            s.host_acc_cap = acc_copy;
        }
        const mc3d_records *rec = J.rec;
        const bool want_packed = rec != nullptr && rec->packed != nullptr && cnt != 0;
        if (want_packed)
            for (const HostCase &c : J.cases)
                if (c.n_rows > MC3D_PACKED_MAX_ROWS)
                    return fail(MC3D_EINVAL, "packed records support tables of <= %d rows per case (got %d): use the record columns", MC3D_PACKED_MAX_ROWS, c.n_rows);
        const bool want_rec = rec != nullptr && cnt != 0 && !want_packed;
        void *const host_col[6] = {rec ? (void *)rec->condition : nullptr, rec ? (void *)rec->wvl_row : nullptr,
                                   rec ? (void *)rec->theta_n : nullptr,   rec ? (void *)rec->phi_n : nullptr,
                                   rec ? (void *)rec->n_scat : nullptr,    rec ? (void *)rec->path_length : nullptr};
        size_t rec_off[6];
        if (want_rec) CUDA_TRY(s.recs.ensure(records_layout(chunk_cap, rec_off)));
        if (want_packed) CUDA_TRY(s.recs.ensure((size_t)chunk_cap * 16));
        // one copy instead of six when the caller's arrays sit in one block packed like the device's
        bool packed_host = want_rec && n_dev == 1 && n_chunks == 1;
        if (packed_host) {
            records_layout(cnt, rec_off);
            for (int c = 0; c < 6; ++c)
                packed_host = packed_host && host_col[c] != nullptr && (uint8_t *)host_col[c] == (uint8_t *)host_col[0] + rec_off[c];
        }
        while ((int)s.ev.size() < 2 * std::max(n_chunks, 1)) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreate(&e));
            s.ev.push_back(e);
        }
        DevRow *d_rows = reinterpret_cast<DevRow *>(s.cst.p);
        double *d_edges = reinterpret_cast<double *>(s.cst.p + rows_bytes);
        double *d_hist_edges = d_edges + n_edges;
        DevCase *d_cases = reinterpret_cast<DevCase *>(s.cst.p + rows_bytes + edges_bytes);
        unsigned long long *d_tally = s.acc.p, *d_extras = s.acc.p + tally_len + 1;
        unsigned long long *d_case_ev = n_case_ev ? d_extras + s.extras_len : nullptr;
        uint32_t *d_counters = reinterpret_cast<uint32_t *>(s.acc.p + acc_copy);

        // ---- inputs: staged in pinned memory, uploaded only when they differ from what the slot's block holds
        DevRow *h_rows = reinterpret_cast<DevRow *>(s.host_cst);
        double *h_edges = reinterpret_cast<double *>(s.host_cst + rows_bytes);
        DevCase *h_cases = reinterpret_cast<DevCase *>(s.host_cst + rows_bytes + edges_bytes);
        memset(h_rows, 0, rows_bytes);   // rows no case uses stay zero
        bool impurity = false;
        for (const HostCase &c : J.cases) {
            const float neg_tau = make_devcase(&c.params, c.row_begin, c.n_rows).neg_tau_tot;
            impurity |= build_rows(&c.params, J.table + c.row_begin, c.n_rows, neg_tau, h_rows + c.row_begin);
        }
        if (J.n_theta_bins > 0) linspace_edges(0.0, 1.5707963267948966, J.n_theta_bins, h_edges);   // np.linspace(0, pi/2, n + 1)
        else h_edges[0] = 0.0;
        linspace_edges(0.0, 6.283185307179586, n_phi, h_edges + J.n_theta_bins + 1);               // np.linspace(0, 2 pi, m + 1)
        if (hist_on) {
            double *h_hist = h_edges + n_edges;
            if (hs.n_scat_bins > 0) linspace_edges(hs.n_scat_lo, hs.n_scat_hi, hs.n_scat_bins, h_hist);
            else h_hist[0] = 0.0;
            if (hs.path_bins > 0) linspace_edges(hs.path_lo, hs.path_hi, hs.path_bins, h_hist + hs.n_scat_bins + 1);
            else h_hist[hs.n_scat_bins + 1] = 0.0;
        }
        {   // sweep: per chunk, the cases it overlaps with their position in the chunk
            size_t at = 0;
            for (int c = 0; J.sweep && c < n_chunks; ++c) {
                const uint64_t g0 = J.range_begin + off + chunks[c].first;
                for (int q = chunk_cases[c].first; q <= chunk_cases[c].second; ++q) {
                    const HostCase &hc = J.cases[q];
                    DevCase dc = make_devcase(&hc.params, hc.row_begin, hc.n_rows);
                    dc.id0 = hc.id_begin - hc.first + g0;                       // id = id0 + pid (modulo 2^64)
                    dc.pid_first = (uint32_t)(std::max(hc.first, g0) - g0);      // < 2^26
                    h_cases[at++] = dc;
                }
            }
        }
        if (!ctx->input_caching || cst_moved || s.cst_shadow.size() != cst_bytes || memcmp(s.cst_shadow.data(), s.host_cst, cst_bytes) != 0) {
            CUDA_TRY(cudaMemcpyAsync(s.cst.p, s.host_cst, cst_bytes, cudaMemcpyHostToDevice, s.stream));
            s.cst_shadow.assign(s.host_cst, s.host_cst + cst_bytes);
        }
        // tallies, event count, extrema (the minima are kept complemented, so everything starts at 0), column
        // histograms, events per case and the claim counters
        CUDA_TRY(cudaMemsetAsync(s.acc.p, 0, acc_words * sizeof(unsigned long long), s.stream));

        // ---- launch configuration: persistent grid, SM count x resident blocks
        // A launch ends with a drain phase in which lanes that found no more photons idle while the longest walks
        // finish; its cost grows with the number of lanes, so a launch gets only as many lanes as keep >= 26 photons
        // per lane (13 for a call that runs alone: its ramp-down is bound by latency, not by issue slots; at most the
        // resident capacity, at least one block per SM).  Small launches then leave room for the next calls in flight
        // on the other slots' streams.
        const int bps_variant = ctx->blocks_per_sm > 0 ? ctx->blocks_per_sm : 1024 / ctx->block_threads;
        const uint32_t max_chunk_cases = [&] { int m = 0; for (auto &cc : chunk_cases) m = std::max(m, cc.second - cc.first + 1); return (uint32_t)(J.sweep ? m : 0); }();
        int resident = 0;
        for (const mc3d_ctx::Occupancy &o : ctx->occupancy)
            if (o.impurity == impurity && o.block_threads == ctx->block_threads && o.bps_variant == bps_variant && o.n_rows == n_rows &&
                o.n_cases == max_chunk_cases)
                resident = o.resident;
        if (resident == 0 && !fused) {
            WalkParams Wq = W;
            Wq.n_cases = max_chunk_cases;
            CUDA_TRY(launch_walk(Wq, impurity, ctx->block_threads, bps_variant, 0, s.stream, &resident));
            if (resident > 0) ctx->occupancy.push_back({impurity, ctx->block_threads, bps_variant, n_rows, max_chunk_cases, resident});
        }
        if (resident < 1 && !fused)
            return fail(MC3D_ECUDA, "walk kernel does not fit on an SM (block %d, rows %d, cases %u)", ctx->block_threads, n_rows, max_chunk_cases);
        resident = std::max(1, std::min(resident, bps_variant));
        if (ctx->blocks_per_sm == 0) {
            const uint64_t per_lane = ctx->per_lane > 0 ? (uint64_t)ctx->per_lane : (lone ? 13 : 26);
            const uint64_t want_blocks = (std::min<uint64_t>(cnt, CHUNK_PHOTONS) / per_lane + ctx->block_threads - 1) / ctx->block_threads;
            const int per_sm = (int)std::min<uint64_t>(resident, std::max<uint64_t>(1, want_blocks / d.sm_count));
            resident = per_sm;
        }
        st.grid_blocks = d.sm_count * resident;
        // a synchronous call that runs alone hands its last photons over to the tail kernel (dense warps, helper lanes).
        // (Only a synchronous call is known to stay alone: the tail kernel's blocks could not become resident behind the
        // persistent grids of calls enqueued after an asynchronous one, and its stream would stall.)
        const bool use_tail = !fused && (ctx->tail_kernel >= 0 ? ctx->tail_kernel != 0 : (lone && J.sync));
        if (use_tail) CUDA_TRY(s.tail.ensure((size_t)TAIL_WORDS * st.grid_blocks * ctx->block_threads));   // (before the timed span)

        size_t case_at = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const uint64_t c_off = chunks[c].first;
            const uint32_t c_cnt = chunks[c].second;
            WalkParams Wc = W;
            Wc.n_photon = c_cnt;
            Wc.rows = d_rows;
            if (J.sweep) {
                Wc.n_cases = (uint32_t)(chunk_cases[c].second - chunk_cases[c].first + 1);
                Wc.case0 = (uint32_t)chunk_cases[c].first;
                Wc.cases = d_cases + case_at;
                case_at += Wc.n_cases;
            } else {
                const HostCase &hc = J.cases[0];
                Wc.c = make_devcase(&hc.params, 0, n_rows);
                Wc.c.id0 = hc.id_begin + J.range_begin + off + c_off;
            }
            Wc.counter = d_counters + 4 * c;
            Wc.fresh = s.fresh.p;
            Wc.raw = s.raw.p;
            const int want = (int)((c_cnt + ctx->block_threads - 1) / ctx->block_threads);
            const int grid = std::max(1, std::min(st.grid_blocks, want));
            CUDA_TRY(cudaEventRecord(s.ev[2 * c], s.stream));
            if (!fused) {
                CUDA_TRY(launch_init(Wc, impurity, d.sm_count, s.stream));
                const int lanes = grid * ctx->block_threads;
                if (use_tail) {
                    Wc.tail = s.tail.p;
                    Wc.n_tail = d_counters + 4 * c + 2;
                    Wc.tail_cap = (uint32_t)lanes;
                }
                CUDA_TRY(launch_walk(Wc, impurity, ctx->block_threads, bps_variant, grid, s.stream, nullptr));
                if (use_tail) CUDA_TRY(launch_tail(Wc, impurity, lanes, s.stream));
            }
            FinalizeParams F;
            memset(&F, 0, sizeof F);
            F.raw = s.raw.p;
            F.rows = d_rows;
            F.edges = d_edges;
            F.n_photon = c_cnt;
            F.n_rows = n_rows;
            F.n_theta_bins = J.n_theta_bins;
            F.n_phi_bins = J.n_phi_bins;
            if (J.sweep) {
                F.cases = Wc.cases;
                F.n_cases = Wc.n_cases;
                F.case0 = Wc.case0;
                F.case_events = d_case_ev;
                for (int q = chunk_cases[c].first; q <= chunk_cases[c].second; ++q) F.win_rows = std::max(F.win_rows, J.cases[q].n_rows);
            }
            if (want_rec) {
                records_layout(c_cnt, rec_off);
                F.condition = host_col[0] ? s.recs.p + rec_off[0] : nullptr;
                F.wvl_row = host_col[1] ? reinterpret_cast<int16_t *>(s.recs.p + rec_off[1]) : nullptr;
                F.theta_n = host_col[2] ? reinterpret_cast<float *>(s.recs.p + rec_off[2]) : nullptr;
                F.phi_n = host_col[3] ? reinterpret_cast<float *>(s.recs.p + rec_off[3]) : nullptr;
                F.n_scat = host_col[4] ? reinterpret_cast<uint32_t *>(s.recs.p + rec_off[4]) : nullptr;
                F.path_length = host_col[5] ? reinterpret_cast<float *>(s.recs.p + rec_off[5]) : nullptr;
            }
            if (want_packed) F.packed = reinterpret_cast<uint4 *>(s.recs.p);
            F.tally = d_tally;
            F.n_events = d_tally + tally_len;
            F.extrema = reinterpret_cast<uint32_t *>(d_extras);
            if (n_hist) {
                F.hist = d_extras + 2;
                F.hist_edges = d_hist_edges;
                F.n_scat_bins = hs.n_scat_bins;
                F.path_bins = hs.path_bins;
                F.path_scale = hs.path_scale;
            }
            if (fused) CUDA_TRY(launch_fused(Wc, F, impurity, d.sm_count, s.stream));
            else CUDA_TRY(launch_finalize(F, d.sm_count, s.stream));
            CUDA_TRY(cudaEventRecord(s.ev[2 * c + 1], s.stream));
            if (want_packed) {
                CUDA_TRY(cudaMemcpyAsync(rec->packed + 4 * (off + c_off), s.recs.p, (size_t)c_cnt * 16, cudaMemcpyDeviceToHost, s.stream));
            } else if (packed_host) {
                CUDA_TRY(cudaMemcpyAsync(host_col[0], s.recs.p, rec_off[5] + (size_t)c_cnt * REC_ITEM[5], cudaMemcpyDeviceToHost, s.stream));
            } else if (want_rec) {
                const uint64_t o = off + c_off;   // this chunk's first photon in the caller's arrays
                for (int c = 0; c < 6; ++c)
                    if (host_col[c])
                        CUDA_TRY(cudaMemcpyAsync((uint8_t *)host_col[c] + o * REC_ITEM[c], s.recs.p + rec_off[c],
                                                 (size_t)c_cnt * REC_ITEM[c], cudaMemcpyDeviceToHost, s.stream));
            }
        }
        if (k != 0 || n_dev == 1)   // device 0 of a multi-device context copies after the reduce below
            CUDA_TRY(cudaMemcpyAsync(s.host_acc, s.acc.p, acc_copy * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        s.busy = true;
        s.n_photon = cnt;
        s.tally_len = tally_len;
        s.n_chunks = n_chunks;
        s.user_tally = J.tally;
        s.user_case_ev = J.case_events;
        s.packed = want_packed;
    }

    // ---- the only collective: sum the tally blocks of all devices of this process onto device 0 (NVLink)
    if (n_dev > 1) {
        NCCL_TRY(g_nccl.GroupStart());
        for (int k = 0; k < n_dev; ++k) {
            Device &d = ctx->devs[k];
            Slot &s = d.slot[slot_idx];
            NCCL_TRY(g_nccl.Reduce(s.acc.p, s.acc.p, tally_len + 1, ncclUint64, ncclSum, 0, ctx->comms[k], s.stream));
        }
        NCCL_TRY(g_nccl.GroupEnd());
    }
    if (n_dev > 1) {
        Device &d = ctx->devs[0];
        Slot &s = d.slot[slot_idx];
        CUDA_TRY(cudaSetDevice(d.id));
        CUDA_TRY(cudaMemcpyAsync(s.host_acc, s.acc.p, (tally_len + 1 + s.extras_len + s.n_case_ev) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    }
    return MC3D_OK;
}

int mc3d_wait(mc3d_ctx *ctx, int slot_idx, mc3d_stats *stats)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (slot_idx < 0 || slot_idx >= N_SLOTS) return fail(MC3D_EINVAL, "slot must be in [0, %d)", N_SLOTS);
    if (!ctx->devs[0].slot[slot_idx].busy) return fail(MC3D_EINVAL, "slot %d has no call in flight", slot_idx);
    mc3d_stats &st = ctx->pending_stats[slot_idx];
    double kernel_ms = 0.0;
    int first_err = MC3D_OK;
    for (Device &d : ctx->devs) {
        Slot &s = d.slot[slot_idx];
        cudaSetDevice(d.id);
        cudaError_t e = cudaStreamSynchronize(s.stream);
        s.busy = false;
        if (e != cudaSuccess) {
            if (!first_err) first_err = fail(MC3D_ECUDA, "device %d: %s", d.id, cudaGetErrorString(e));
            continue;
        }
        double ms = 0.0;
        for (int c = 0; c < s.n_chunks; ++c) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, s.ev[2 * c], s.ev[2 * c + 1]) == cudaSuccess) ms += t;
        }
        kernel_ms = std::max(kernel_ms, ms);
    }
    if (first_err) return first_err;
    {   // extrema and column histograms: combined over the devices on the host (a few KB)
        uint32_t e[4] = {0xffffffffu, 0u, 0xffffffffu, 0u};
        const mc3d_hist_spec &hs = ctx->done_spec[slot_idx];
        const size_t n_hist = ctx->done_hist[slot_idx] ? (size_t)hs.n_scat_bins + (size_t)hs.path_bins : 0;
        std::vector<uint64_t> &counts = ctx->done_counts[slot_idx];
        counts.assign(n_hist, 0);
        for (Device &d : ctx->devs) {
            Slot &s = d.slot[slot_idx];
            const unsigned long long *h_extras = s.host_acc + s.tally_len + 1;
            uint32_t de[4];
            memcpy(de, h_extras, sizeof de);
            de[0] = ~de[0];
            de[2] = ~de[2];
            e[0] = std::min(e[0], de[0]); e[1] = std::max(e[1], de[1]);
            e[2] = std::min(e[2], de[2]); e[3] = std::max(e[3], de[3]);
            for (size_t k = 0; k < n_hist && 2 + k < s.extras_len; ++k) counts[k] += h_extras[2 + k];
        }
        mc3d_extrema &x = ctx->done_extrema[slot_idx];
        if (e[0] > e[1]) { e[0] = e[1] = 0u; e[2] = e[3] = 0u; }   // no photons
        x.n_scat_min = e[0]; x.n_scat_max = e[1];
        st.packed_saturated = (ctx->devs[0].slot[slot_idx].packed && e[1] > MC3D_PACKED_NSCAT_MAX) ? 1 : 0;
        memcpy(&x.path_min, &e[2], 4); memcpy(&x.path_max, &e[3], 4);
    }
    Slot &s0 = ctx->devs[0].slot[slot_idx];
    st.n_events = s0.host_acc[s0.tally_len];
    if (s0.user_tally) memcpy(s0.user_tally, s0.host_acc, s0.tally_len * sizeof(uint64_t));
    if (s0.user_case_ev && s0.n_case_ev) {   // events per case: summed over the devices on the host
        for (size_t c = 0; c < s0.n_case_ev; ++c) s0.user_case_ev[c] = 0;
        for (Device &d : ctx->devs) {
            const Slot &s = d.slot[slot_idx];
            const unsigned long long *ev = s.host_acc + s.tally_len + 1 + s.extras_len;
            for (size_t c = 0; c < s0.n_case_ev; ++c) s0.user_case_ev[c] += ev[c];
        }
    }
    st.kernel_ms = kernel_ms;
    st.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ctx->t0[slot_idx]).count();
    if (stats) *stats = st;
    return MC3D_OK;
}

int mc3d_run(mc3d_ctx *ctx, const mc3d_params *params, const mc3d_ssp_row *table, int n_rows, uint64_t seed,
             uint64_t photon_begin, uint64_t n_photon, const mc3d_records *records, uint64_t *tally,
             mc3d_stats *stats)
{
    if (check_ctx(ctx) == MC3D_OK) ctx->sync_call = true;
    int rc = mc3d_run_async(ctx, 0, params, table, n_rows, seed, photon_begin, n_photon, records, tally, stats);
    if (check_ctx(ctx) == MC3D_OK) ctx->sync_call = false;
    if (rc) return rc;
    return mc3d_wait(ctx, 0, stats);
}

int mc3d_set_histograms(mc3d_ctx *ctx, const mc3d_hist_spec *spec)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!spec) { ctx->hist_on = false; return MC3D_OK; }
    if (spec->n_scat_bins < 0 || spec->path_bins < 0 || spec->n_scat_bins > (1 << 20) || spec->path_bins > (1 << 20))
        return fail(MC3D_EINVAL, "histogram bin counts must be in [0, 2^20]");
    if (spec->n_scat_bins > 0 && !(spec->n_scat_hi > spec->n_scat_lo))
        return fail(MC3D_EINVAL, "n_scat histogram range must have hi > lo (np.histogram widens an empty range by +-0.5)");
    if (spec->path_bins > 0 && !(spec->path_hi > spec->path_lo))
        return fail(MC3D_EINVAL, "path histogram range must have hi > lo (np.histogram widens an empty range by +-0.5)");
    if (spec->path_bins > 0 && !(spec->path_scale > 0.0)) return fail(MC3D_EINVAL, "path_scale must be positive");
    ctx->hist_spec = *spec;
    ctx->hist_on = spec->n_scat_bins > 0 || spec->path_bins > 0;
    return MC3D_OK;
}

int mc3d_get_histograms(mc3d_ctx *ctx, int slot_idx, uint64_t *n_scat_counts, uint64_t *path_counts, mc3d_extrema *extrema)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (slot_idx < 0 || slot_idx >= N_SLOTS) return fail(MC3D_EINVAL, "slot must be in [0, %d)", N_SLOTS);
    if (ctx->devs[0].slot[slot_idx].busy) return fail(MC3D_EINVAL, "slot %d is still in flight; call mc3d_wait first", slot_idx);
    if (extrema) *extrema = ctx->done_extrema[slot_idx];
    if (n_scat_counts || path_counts) {
        if (!ctx->done_hist[slot_idx]) return fail(MC3D_EINVAL, "the last call on slot %d ran without mc3d_set_histograms", slot_idx);
        const mc3d_hist_spec &hs = ctx->done_spec[slot_idx];
        const std::vector<uint64_t> &c = ctx->done_counts[slot_idx];
        if (c.size() != (size_t)hs.n_scat_bins + (size_t)hs.path_bins) return fail(MC3D_EINVAL, "no histogram result on slot %d", slot_idx);
        if (n_scat_counts && hs.n_scat_bins) memcpy(n_scat_counts, c.data(), hs.n_scat_bins * sizeof(uint64_t));
        if (path_counts && hs.path_bins) memcpy(path_counts, c.data() + hs.n_scat_bins, hs.path_bins * sizeof(uint64_t));
    }
    return MC3D_OK;
}

int mc3d_reduce_tally(mc3d_ctx *ctx, uint64_t *tally, uint64_t n, int root)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!tally && n) return fail(MC3D_EINVAL, "tally is null");
    if (ctx->world == 1) return MC3D_OK;
    if (root < 0 || root >= ctx->world) return fail(MC3D_EINVAL, "root %d out of range", root);
    Device &d = ctx->devs[0];
    cudaStream_t stream0 = d.slot[0].stream;
    CUDA_TRY(cudaSetDevice(d.id));
    CUDA_TRY(ctx->reduce_buf.ensure(std::max<uint64_t>(n, 1)));   // kept by the context: no allocation per call
    unsigned long long *buf = ctx->reduce_buf.p;
    cudaError_t e = cudaMemcpyAsync(buf, tally, n * sizeof(uint64_t), cudaMemcpyHostToDevice, stream0);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) r = g_nccl.Reduce(buf, buf, n, ncclUint64, ncclSum, root, ctx->comms[0], stream0);
    if (e == cudaSuccess && r == ncclSuccess && ctx->rank == root)
        e = cudaMemcpyAsync(tally, buf, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream0);
    cudaError_t e2 = cudaStreamSynchronize(stream0);
    if (r != ncclSuccess) return fail(MC3D_ENCCL, "ncclReduce failed: %s", g_nccl.GetErrorString(r));
    if (e != cudaSuccess) return fail(MC3D_ECUDA, "tally copy failed: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return fail(MC3D_ECUDA, "stream sync failed: %s", cudaGetErrorString(e2));
    return MC3D_OK;
}

int mc3d_replay(mc3d_ctx *ctx, const mc3d_params *P, uint64_t n, const double *wvl, const double *ssa_ice,
                const double *ssa_imp, const double *g, const double *ext_cff_mss, const double *p_ext_imp,
                const double *init_draws, const int64_t *offsets, const double *stream, const mc3d_records_f64 *out,
                uint64_t *n_mismatch)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!P || !wvl || !ssa_ice || !ssa_imp || !g || !ext_cff_mss || !p_ext_imp || !init_draws || !offsets || !out)
        return fail(MC3D_EINVAL, "null argument");
    if (n_mismatch) *n_mismatch = 0;
    if (n == 0) return MC3D_OK;
    if (n >= (1ull << 31)) return fail(MC3D_EINVAL, "replay supports < 2^31 photons per call");
    if (offsets[0] != 0) return fail(MC3D_ESTREAM, "offsets[0] must be 0");
    for (uint64_t p = 0; p < n; ++p)
        if (offsets[p + 1] < offsets[p]) return fail(MC3D_ESTREAM, "offsets must be non-decreasing (photon %llu)", (unsigned long long)p);
    const uint64_t n_stream = (uint64_t)offsets[n];
    if (n_stream && !stream) return fail(MC3D_EINVAL, "stream is null");
    Device &d = ctx->devs[0];
    cudaStream_t stream0 = d.slot[0].stream;
    CUDA_TRY(cudaSetDevice(d.id));

    // one arena: 6 + 3 per-photon double inputs, offsets, stream, then outputs
    const size_t n_in = 9 * n, n_outd = 5 * n;
    double *d_in = nullptr, *d_stream = nullptr, *d_outd = nullptr;
    long long *d_off = nullptr, *d_outl = nullptr;
    int *d_cond = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_in); cudaFree(d_stream); cudaFree(d_outd); cudaFree(d_off); cudaFree(d_outl); cudaFree(d_cond);
    };
#define TRY_OR_CLEAN(expr)                                                                       \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) { cleanup(); return fail(MC3D_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } \
    } while (0)
    TRY_OR_CLEAN(cudaMalloc((void **)&d_in, n_in * sizeof(double)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_stream, std::max<uint64_t>(n_stream, 1) * sizeof(double)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_outd, n_outd * sizeof(double)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_off, (n + 1) * sizeof(long long)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_outl, 2 * n * sizeof(long long)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_cond, n * sizeof(int)));
    const double *ins[6] = {wvl, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp};
    for (int k = 0; k < 6; ++k)
        TRY_OR_CLEAN(cudaMemcpyAsync(d_in + k * n, ins[k], n * sizeof(double), cudaMemcpyHostToDevice, stream0));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_in + 6 * n, init_draws, 3 * n * sizeof(double), cudaMemcpyHostToDevice, stream0));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_off, offsets, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream0));
    if (n_stream)
        TRY_OR_CLEAN(cudaMemcpyAsync(d_stream, stream, n_stream * sizeof(double), cudaMemcpyHostToDevice, stream0));

    ReplayParams R;
    memset(&R, 0, sizeof R);
    R.theta0_rad = P->theta0_rad; R.tau_tot = P->tau_tot; R.rho_snw = P->rho_snw; R.r_lambert = P->r_lambert;
    R.flags = P->flags;
    R.n_photon = (uint32_t)n;
    R.wvl = d_in; R.ssa_ice = d_in + n; R.ssa_imp = d_in + 2 * n; R.g = d_in + 3 * n; R.ext_cff_mss = d_in + 4 * n;
    R.p_ext_imp = d_in + 5 * n; R.init_draws = d_in + 6 * n;
    R.offsets = d_off; R.stream = d_stream;
    R.condition = d_cond;
    R.wvn = d_outd; R.theta_n = d_outd + n; R.phi_n = d_outd + 2 * n; R.path_length = d_outd + 3 * n;
    R.snow_depth = d_outd + 4 * n;
    R.n_scat = d_outl; R.consumed = d_outl + n;
    TRY_OR_CLEAN(launch_replay(R, stream0));

    std::vector<long long> consumed(n);
#define BACK(dst, src, T) \
    if (dst) TRY_OR_CLEAN(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, stream0))
    BACK(out->condition, d_cond, int);
    BACK(out->wvn, R.wvn, double);
    BACK(out->theta_n, R.theta_n, double);
    BACK(out->phi_n, R.phi_n, double);
    BACK(out->path_length, R.path_length, double);
    BACK(out->snow_depth, R.snow_depth, double);
    BACK(out->n_scat, R.n_scat, long long);
    BACK(consumed.data(), R.consumed, long long);
#undef BACK
    TRY_OR_CLEAN(cudaStreamSynchronize(stream0));
#undef TRY_OR_CLEAN
    cleanup();
    uint64_t mism = 0;
    for (uint64_t p = 0; p < n; ++p) {
        if (out->consumed) out->consumed[p] = consumed[p];
        if (consumed[p] != offsets[p + 1] - offsets[p]) ++mism;
    }
    if (n_mismatch) *n_mismatch = mism;
    return MC3D_OK;
}

}  // extern "C"
