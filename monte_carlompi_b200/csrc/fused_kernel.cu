// fused_kernel.cu -- the short-walk path: one kernel, nothing but the finished record touches HBM.
//
// Strongly absorbing grains (NIR wavelengths beyond ~1.4 um, large effective radii: config C3 of BASELINE.json) or
// optically thin slabs end a photon after a handful of events.  The persistent three-kernel path then spends its
// time moving per-photon state through HBM (16 B `fresh` written + read, 32 B raw result written + read, then the
// record: ~115 B per photon against 16-19 B of record), on three launches, and on lanes that wait for the rest of
// their group of four events.  Here a warp runs the same device functions as three full-width stages that feed each
// other through two small queues in shared memory:
//
//   intake    32 consecutive photon ids: wavelength draw + first event (first_event, monte_carlo3D.py:1515-1520,
//             1035-1038, 1232-1237).  Survivors -> queue A (state after the first event, 10 words), the rest -> queue D.
//             Runs when lanes are free and queue A is empty.
//   walk      ONE event of every photon the warp's lanes carry (scatter_and_move + the attention predicate, then
//             resolve() for the lanes that need it: monte_carlo3D.py:1212-1466).  Finished -> queue D; the freed lane
//             takes the next photon from queue A.  A lane remembers its position in its group of the walk stream
//             (block number, slot, the words of the last Philox block not used yet), so it consumes exactly the
//             words group() would.
//   finalize  pops 32 finished photons from D: angles, record and tallies (finalize_photon, monte_carlo3D.py:1468-1490,
//             post_processing.py:73-76).
//
// Every stage runs with (nearly) all 32 lanes busy however short or uneven the walks are; the price is one warp vote
// and the slot bookkeeping per event (the persistent kernel votes once per four), which is why the host picks this
// kernel only for short walks (mc3d_api.cu: expected_events).  The photon's random stream depends only on (seed, photon id, event number) and the
// arithmetic is the same inlined code, so the records are bit-identical to the persistent path's (tested).
#include <algorithm>

#include "finalize_device.cuh"
#include "walk_device.cuh"

namespace mc3d {

constexpr int FQ_CAP = 64;        // entries per queue and warp: a stage runs when >= 32 are waiting, and adds <= 32
constexpr int FQ_WALK_WORDS = 10; // z, ux, uy, uz, path_lo, path_hi, i, plo, phi, row_addr of a photon after its first event
constexpr int FQ_DONE_WORDS = 8;  // pid, ux, uy, uz, path, n_scat, cond | row << 8, lcase

struct FusedQueues {
    uint32_t walk[FQ_WALK_WORDS][32];
    uint32_t done[FQ_DONE_WORDS][FQ_CAP];
};

__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float u2f(uint32_t x) { return __uint_as_float(x); }

// Append the lanes with `pred` to a queue (warp-uniform length `cnt`): returns this lane's entry.
__device__ __forceinline__ uint32_t queue_slot(bool pred, uint32_t lane, uint32_t &cnt)
{
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    const uint32_t at = cnt + __popc(m & ((1u << lane) - 1u));
    cnt += __popc(m);
    return at;
}

__device__ __forceinline__ void push_done(FusedQueues &Q, uint32_t at, uint32_t pid, const Lane &L, uint32_t cond, uint32_t row,
                                          uint32_t lcase)
{
    Q.done[0][at] = pid; Q.done[1][at] = f2u(L.ux); Q.done[2][at] = f2u(L.uy); Q.done[3][at] = f2u(L.uz);
    Q.done[4][at] = f2u(__fadd_rn(L.path_hi, L.path_lo)); Q.done[5][at] = L.i - 1u; Q.done[6][at] = cond | (row << 8);
    Q.done[7][at] = lcase;
}

// Finalize the last `k` entries of queue D (k <= 32).
template <bool SWEEP>
__device__ __forceinline__ void flush_done(const WalkParams &P, const FinalizeParams &F, FinalizeBlock &B, FusedQueues &Q,
                                           const DevRow *rows, const DevCase *cases, uint32_t lane, uint32_t &cnt, uint32_t k)
{
    __syncwarp();
    if (lane < k) {
        const uint32_t at = cnt - k + lane;
        const uint32_t meta = Q.done[6][at], row = meta >> 8, lcase = Q.done[7][at];
        const DevCase &C = SWEEP ? cases[lcase] : P.c;
        finalize_photon<SWEEP>(F, B, Q.done[0][at], u2f(Q.done[1][at]), u2f(Q.done[2][at]), u2f(Q.done[3][at]), u2f(Q.done[4][at]),
                               Q.done[5][at], meta & 0xffu, row, row - C.row_begin, lcase, rows[row].inv_ext);
    }
    cnt -= k;
    __syncwarp();
}

template <bool IMP, bool SWEEP, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2) fused_kernel(const __grid_constant__ WalkParams P, const __grid_constant__ FinalizeParams F)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned int statics[6];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    const DevCase *cases = staged_cases(P, smem_raw);
    stage_tables(P, smem_raw, BLOCK);
    FusedQueues *queues = reinterpret_cast<FusedQueues *>(smem_raw + ((tables_bytes(P.n_rows, P.n_cases) + 15) & ~(size_t)15));
    FinalizeBlock B;
    finalize_begin<BLOCK>(F, B, reinterpret_cast<unsigned int *>(queues + BLOCK / 32), statics);   // ends with __syncthreads()
    const uint32_t rows_addr = shared_address(rows);
    const uint32_t lane = threadIdx.x & 31u;
    FusedQueues &Q = queues[threadIdx.x >> 5];
    uint32_t n_walk = 0u, n_done = 0u;                            // warp-uniform queue lengths
    const uint32_t batch_stride = gridDim.x * BLOCK;
    uint32_t batch = blockIdx.x * BLOCK + (threadIdx.x & ~31u);   // first photon of the warp's next intake

    // the photon this lane is walking: its state, and its place in its group of the walk stream (first block of the
    // group << 2 | slot, and the words of the group's last Philox block that later slots use)
    Lane L;
    L.z = 0.f; L.ux = 0.f; L.uy = 0.f; L.uz = -1.f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 0u; L.blk = 0u; L.plo = 0u; L.phi = 0u; L.row_addr = rows_addr; L.key = 0u; L.imp = false;
    L.pk.pB = L.pk.pC = L.pk.pD = 0u;
    uint32_t bs = 0u, w0 = 0u, w1 = 0u, w2 = 0u;
    bool alive = false;

    // A lane is walking (alive), waiting (its last event needs attention: L.i != 0, !alive) or free (L.i == 0).
    const uint32_t threshold = max(1u, min(32u, 2u * P.refill_threshold));
    for (;;) {
        const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
        if (32u - (uint32_t)__popc(alive_mask) >= threshold) {
            // ---- service, one uniform branch for all lanes that are not walking (as in the persistent kernel):
            // 1. waiting lanes run the termination chain of their last event together
            uint32_t cond = ALIVE, row = 0u, lcase = 0u, pid = 0u;
            const bool waiting = !alive && L.i != 0u;
            if (waiting) {
                row = lane_row(L, rows_addr);
                lcase = lane_lcase<SWEEP>(P, L);
                const DevCase &C = SWEEP ? cases[lcase] : P.c;
                pid = lane_pid(C, L);
                cond = resolve<IMP>(P, C, lane_phi<SWEEP>(P, L), rows[row], L);
                alive = cond == ALIVE;                       // bs already points to slot 0 of the next group
            }
            const bool ended = waiting && !alive;
            const uint32_t ad = queue_slot(ended, lane, n_done);
            if (ended) {
                push_done(Q, ad, pid, L, cond, row, lcase);
                L.i = 0u;
            }
            if (n_done >= 32u) flush_done<SWEEP>(P, F, B, Q, rows, cases, lane, n_done, 32u);
            // 2. intake when queue A is empty: 32 new photons through their first event
            if (n_walk == 0u && batch < P.n_photon) {
                const uint32_t npid = batch + lane;
                batch += batch_stride;
                Lane N;
                uint32_t nrow = 0u, redo, nlcase = 0u, ncond = 1u;
                float dtau;
                const bool have = npid < P.n_photon;
                if (have) {
                    nlcase = find_case<SWEEP>(P, cases, npid);
                    const DevCase &C = SWEEP ? cases[nlcase] : P.c;
                    const uint64_t id = C.id0 + npid;
                    ncond = first_event<IMP>(P, C, (uint32_t)(id >> 32), rows, rows_addr, (uint32_t)id, N, nrow, dtau, redo);
                }
                const bool walks = have && ncond == ALIVE;
                const uint32_t aw = queue_slot(walks, lane, n_walk);
                if (walks) {
                    Q.walk[0][aw] = f2u(N.z); Q.walk[1][aw] = f2u(N.ux); Q.walk[2][aw] = f2u(N.uy); Q.walk[3][aw] = f2u(N.uz);
                    Q.walk[4][aw] = f2u(N.path_lo); Q.walk[5][aw] = f2u(N.path_hi); Q.walk[6][aw] = N.i;
                    Q.walk[7][aw] = N.plo; Q.walk[8][aw] = N.phi; Q.walk[9][aw] = N.row_addr;
                }
                const uint32_t nd = queue_slot(have && !walks, lane, n_done);
                if (have && !walks) push_done(Q, nd, npid, N, ncond, nrow, nlcase);
                __syncwarp();
            }
            // 3. free lanes take over photons waiting in queue A
            const uint32_t free_mask = __ballot_sync(0xffffffffu, !alive);
            if (n_walk != 0u) {
                const uint32_t rank = __popc(free_mask & ((1u << lane) - 1u));
                const uint32_t k = min(n_walk, (uint32_t)__popc(free_mask));
                if (!alive && rank < k) {
                    const uint32_t at = n_walk - 1u - rank;
                    L.z = u2f(Q.walk[0][at]); L.ux = u2f(Q.walk[1][at]); L.uy = u2f(Q.walk[2][at]); L.uz = u2f(Q.walk[3][at]);
                    L.path_lo = u2f(Q.walk[4][at]); L.path_hi = u2f(Q.walk[5][at]); L.i = Q.walk[6][at];
                    L.plo = Q.walk[7][at]; L.phi = Q.walk[8][at]; L.row_addr = Q.walk[9][at];
                    L.pk = philox_walk_constants(L.plo, P.rk);
                    bs = 0u;
                    alive = true;
                }
                n_walk -= k;
                __syncwarp();
            } else if (free_mask == 0xffffffffu && batch >= P.n_photon) {
                break;   // nothing left to take in, nobody walking or waiting
            }
            if (n_done >= 32u) flush_done<SWEEP>(P, F, B, Q, rows, cases, lane, n_done, 32u);
        }
        // ---- walk: one event of every walking photon; a lane whose event needs attention stops and waits
        if (alive) {
            const uint32_t slot = bs & 3u, n = bs >> 2;
            L.blk = n + GROUP_BLOCKS;                        // where the photon continues when this event needs attention
            const uint32_t phi = lane_phi<SWEEP>(P, L);
            const HotRow H = load_hot_row<SWEEP>(P, L.row_addr);
            // the event's words (HG deflection, azimuth, free path) and the ones left for the group's later slots,
            // picked with masks (slots differ from lane to lane: no branches):
            //   slot 0: block a -> a.x a.y a.z | left a.w        slot 2: block c -> w0 w1 c.x | left c.y c.z c.w
            //   slot 1: block b -> w0 b.x b.y  | left b.z b.w    slot 3: w0 w1 w2
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (slot < 3u) v = philox_walk(n + slot, phi, L.pk, P.rk);
            const uint32_t m0 = slot == 0u ? 0xffffffffu : 0u, m1 = slot == 1u ? 0xffffffffu : 0u;
            const uint32_t m2 = slot == 2u ? 0xffffffffu : 0u, m3 = slot == 3u ? 0xffffffffu : 0u;
            const uint32_t e0 = (v.x & m0) | (w0 & ~m0);
            const uint32_t e1 = (v.y & m0) | (v.x & m1) | (w1 & (m2 | m3));
            const uint32_t e2 = (v.z & m0) | (v.y & m1) | (v.x & m2) | (w2 & m3);
            w0 = (v.w & m0) | (v.z & m1) | (v.y & m2);
            w1 = (v.w & m1) | (v.z & m2);
            w2 = v.w;
            alive = event<IMP>(P, phi, rows, rows_addr, L, H, e0, e1, e2);
            // next slot (after slot 3: slot 0 of the next group); an event that needs attention ends the group
            bs = alive ? bs + 1u + (m3 & 8u) : (bs & ~3u) + (GROUP_BLOCKS << 2);
        }
    }
    if (n_done) flush_done<SWEEP>(P, F, B, Q, rows, cases, lane, n_done, n_done);
    finalize_flush<BLOCK>(F, B);
}

template <bool IMP, bool SWEEP>
static cudaError_t launch_fused_variant(const WalkParams &P, const FinalizeParams &F, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    FinalizeParams Q = F;
    const size_t tab_bytes = ((tables_bytes(P.n_rows, P.n_cases) + 15) & ~(size_t)15) + (BLOCK / 32) * sizeof(FusedQueues);
    // the tally block shares the SM with the tables and the queues: keep two blocks resident
    const size_t smem = tab_bytes + finalize_plan_smem(Q, tab_bytes < 72 * 1024 ? 108 * 1024 - tab_bytes : 0, 16 * 1024, false);
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
