// walk_kernel.cu -- production-mode photon random walk (fp32, sm_100a).
//
// Replaces the Python loop reference monte_carloMPI/monte_carlo3D.py:1613-1616 and everything it calls per
// event: populate_pdfs (885-921, 1010-1025), initial_pdfs (1027-1044), Henyey_Greenstein2 (790-800) and the walk
// body monte_carlo3D (1111-1490), plus the per-photon wavelength draw (1515-1520).
//
// Execution model
//   * persistent warps; every lane walks one photon at a time.  The hot loop runs one GROUP of the photon's walk
//     stream per iteration: four scattering events on three Philox4x32-7 blocks (three words per event -> HG
//     deflection, azimuth, free path; the coarse absorption key is made of the low bytes the float conversions
//     ignore), each event = rotate, move, and ONE predicate "left the top / hit the bottom / maybe absorbed / due
//     for renormalisation".  A lane that trips it simply stops and waits; if its photon survives it continues with
//     the next group (mc3d_device.cuh: the random-number layout).
//   * when `refill_threshold` lanes of a warp are waiting, the warp takes one uniform branch: the waiting lanes
//     are resolved together (the reference's termination chain in its order; finished photons store one 32-byte
//     raw record) and the empty ones pop a fresh photon from a per-warp shared-memory ring.  The ring is refilled
//     32 (short walks: 96) photons per atomicAdd (warp-aggregated by construction) with coalesced 512-byte loads from the
//     `fresh` list the init kernel wrote (wavelength draw, SSP row and the deflection-free first event,
//     monte_carlo3D.py:1232-1237, happen there; entry pid is photon pid, the ones that ended on their first event
//     are marked dead and dropped by a warp vote as the ring is filled).  So the divergent part of a refill is a
//     handful of shared-memory loads, and walk-length divergence is bounded by the threshold instead of by the
//     longest walk in the warp.  Nothing heavy is inlined next to the event loop: no spills at 64 registers.
//   * when the id range is exhausted, what the warp does with the photons it still carries depends on the call:
//       - a synchronous call that runs alone: its tail is the dependent chain of its longest walks.  The lanes hand
//         their photons to the tail kernel (tail_kernel.cu: dense warps, helper lanes) and leave;
//       - other launches share the GPU (the host sets P.drain_give): the warp drains by itself, and since draining
//         warps empty out exponentially (a warp would run ~4 mean walk lengths for its last lane) they consolidate
//         through a small per-block pool in shared memory: a warp with few photons left hands them (eleven words of
//         state each) to the pool and exits, warps with free lanes take them over, so the issue slots go to the
//         co-resident launches instead of to mostly-empty warps;
//       - an asynchronous call that happens to start alone: drains by itself with the latency-oriented group.
//   * angles, records and tallies are produced from the raw records by the coalesced finalize kernel
//     (finalize_kernel.cu).
//   * per-photon results depend only on (seed, photon id): bit-identical for any grid, block or GPU count.
#include "walk_device.cuh"

namespace mc3d {

#ifdef MC3D_TIMELINE
// debug build (make timeline): per-warp %globaltimer at kernel entry, at exhaustion of the fresh list and at exit
__device__ unsigned long long g_timeline[3 * 16384];
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TIMELINE(slot) do { if (lane == 0) { const uint32_t w_ = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); if (w_ < 16384u) g_timeline[3 * w_ + (slot)] = now_ns(); } } while (0)
#else
#define TIMELINE(slot) do { } while (0)
#endif

constexpr int RING = 128;  // entries per warp; a refill adds at most P.claim <= 96 to fewer than 32 leftovers

// Fresh photons staged for the lanes of one warp.
struct WarpRing {
    uint4 entry[RING];   // Fresh{pid, row, dtau, pad}
};

// Photons in flight handed from one draining warp of the block to another (phase B).  The entries and `count` are
// only touched under `lock`.  `draining` counts the warps currently in phase B: a warp donates only while another
// one is there to receive, and no warp leaves while the pool holds photons, so nothing is ever stranded.
constexpr int POOL_CAP = 64;
constexpr int POOL_WORDS = 11;   // z, ux, uy, uz, path_lo, path_hi, i, plo, row_addr, blk, phi (sweep launches)
struct DrainPool {
    uint32_t lock, count, draining, pad;
    uint32_t word[POOL_WORDS][POOL_CAP];
};

// Pool words are read and written with shared-memory atomics: ordering comes from the lock, but this keeps every
// cross-warp access explicit (and compute-sanitizer's racecheck, which does not model locks, quiet).
__device__ __forceinline__ uint32_t pool_get(uint32_t *p) { return atomicOr(p, 0u); }
__device__ __forceinline__ void pool_put(uint32_t *p, uint32_t v) { atomicExch(p, v); }

// Try-lock: a warp that does not get the pool simply walks on and tries again at its next event.
__device__ __forceinline__ bool pool_try_acquire(DrainPool &D, uint32_t lane)
{
    uint32_t got = 0u;
    if (lane == 0) got = atomicCAS(&D.lock, 0u, 1u) == 0u ? 1u : 0u;
    got = __shfl_sync(0xffffffffu, got, 0);
    if (got) __threadfence_block();
    return got != 0u;
}
__device__ __forceinline__ void pool_release(DrainPool &D, uint32_t lane)
{
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(&D.lock, 0u);
}

// One warp vote per group of four events: the vote + threshold test costs ~14 issue cycles, and a stopped lane idles
// at most three events before it is noticed.  (Strongly absorbing or optically thin media, a few events per photon,
// take the fused one-thread-per-photon kernel instead: fused_kernel.cu.)
// SWEEP: the launch walks photons of several cases (mc3d_run_sweep); a lane then carries the high word of its photon
// id (which names the case) and reads the slab bottom with its SSP row.  Everything else is the same code.
template <bool IMP, bool SWEEP, int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) walk_kernel(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    const DevCase *cases = staged_cases(P, smem_raw);
    WarpRing *rings = reinterpret_cast<WarpRing *>(smem_raw + tables_bytes(P.n_rows, P.n_cases));
    DrainPool &D = *reinterpret_cast<DrainPool *>(rings + BLOCK / 32);
    if (threadIdx.x == 0) { D.lock = 0u; D.count = 0u; D.draining = 0u; D.pad = 0u; }
    stage_tables(P, smem_raw, BLOCK);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t rows_addr = shared_address(rows);
    WarpRing &Q = rings[threadIdx.x >> 5];
    uint32_t ring_head = 0, ring_tail = 0;   // warp-uniform, free-running
    const uint32_t threshold = max(1u, min(32u, P.refill_threshold));
    const uint32_t n_fresh = P.n_photon;     // one entry per photon of the launch; the ones that ended on event 1 are dead
    // photons claimed per atomicAdd: 32, or 64 / 96 for short walks, where one claim per 32 photons would run into the
    // rate at which L2 serialises returning atomics on one address (~1 per ns)
    const uint32_t claim = max(32u, min(96u, P.claim));

    Lane L;
    L.z = 0.f; L.ux = 0.f; L.uy = 0.f; L.uz = -1.f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 0; L.blk = 0; L.plo = 0; L.phi = 0; L.row_addr = rows_addr; L.key = 0; L.imp = false;
    L.pk.pB = L.pk.pC = L.pk.pD = 0;
    bool alive = false;   // false: the lane is waiting (its last event needs attention, or it carries no photon)
    TIMELINE(0);

    // ---- phase A: steady state.  Lanes that stop just wait; when `threshold` of them are waiting the warp takes
    // one uniform branch that resolves them together and refills the empty ones from the ring.
    for (;;) {
        const uint32_t waiting = __ballot_sync(0xffffffffu, !alive);
        if (__popc(waiting) >= threshold) {
            if (!alive && L.i != 0u) alive = resolve_lane<IMP, SWEEP>(P, cases, rows, rows_addr, L);
            const uint32_t empty = __ballot_sync(0xffffffffu, !alive);
            const uint32_t need = __popc(empty);
            bool exhausted = false;
            while ((ring_tail - ring_head) < need) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(P.counter, claim);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_fresh) { exhausted = true; break; }
                const uint32_t got = min(claim, n_fresh - base);
                for (uint32_t k0 = 0u; k0 < got; k0 += 32u) {   // live entries only, compacted by a warp vote
                    uint4 f = make_uint4(0u, FRESH_DEAD, 0u, 0u);
                    if (k0 + lane < got) f = *reinterpret_cast<const uint4 *>(P.fresh + base + k0 + lane);
                    const uint32_t live = __ballot_sync(0xffffffffu, f.y != FRESH_DEAD);
                    if (f.y != FRESH_DEAD) Q.entry[(ring_tail + __popc(live & ((1u << lane) - 1u))) & (RING - 1)] = f;
                    ring_tail += __popc(live);
                }
                __syncwarp();
            }
            const uint32_t avail = ring_tail - ring_head;
            if (!alive) {
                const uint32_t rank = __popc(empty & ((1u << lane) - 1u));
                if (rank < avail) {
                    const uint4 f = Q.entry[(ring_head + rank) & (RING - 1)];
                    const float dtau = __uint_as_float(f.z);
                    const DevCase &C = SWEEP ? cases[f.y >> 12] : P.c;
                    if (SWEEP) {
                        const uint64_t id = C.id0 + f.x;
                        L.plo = (uint32_t)id;
                        L.phi = (uint32_t)(id >> 32);
                    } else {
                        L.plo = (uint32_t)C.id0 + f.x;
                    }
                    L.pk = philox_walk_constants(L.plo, P.rk);
                    L.row_addr = rows_addr + (SWEEP ? f.y & 0xfffu : f.y) * (uint32_t)sizeof(DevRow);
                    L.blk = 0u;
                    L.ux = C.mu0x; L.uy = 0.0f; L.uz = C.mu0z;
                    L.z = __fmul_rn(dtau, C.mu0z);
                    L.path_lo = dtau;
                    L.path_hi = 0.0f;
                    L.i = 1u;
                    // f.w != 0 (Fresh::redo): event 1 needed attention -- a reflecting Lambertian bottom on the
                    // first step, or a fine absorption test / renormalisation.  The lane starts out waiting, so the
                    // next resolve pass redoes that event's chain (same draws, same result)
                    alive = f.w == 0u;
                    if (!alive) { L.key = f.w & 0xffff0000u; L.imp = (f.w & 2u) != 0u; }
                }
            }
            ring_head += min(avail, need);
            __syncwarp();
            if (exhausted) break;   // the ring is empty and the list has been handed out: drain
        }
        if (alive) alive = group<IMP, false, SWEEP>(P, rows, rows_addr, L);
    }

    // ---- phase B: drain.  Nothing left to hand out; every lane finishes the photon it carries.  A launch that runs
    // alone (drain_give == 0) is now bound by the dependent chain of its longest walks: group_latency().  With other
    // launches in flight the issue slots matter instead: eager group, and the warps consolidate through the block's
    // pool (see DrainPool).
    TIMELINE(1);
    if (P.tail) {
        // a call that runs alone: its tail is the dependent chain of its longest walks, and that belongs to a kernel of
        // its own (tail_kernel.cu: dense warps, helper lanes).  Resolve what is waiting, hand the walking photons over.
        if (!alive && L.i != 0u) alive = resolve_lane<IMP, SWEEP>(P, cases, rows, rows_addr, L);
        const uint32_t m = __ballot_sync(0xffffffffu, alive);
        uint32_t at = 0u;
        if (lane == 0 && m) at = atomicAdd(P.n_tail, (uint32_t)__popc(m));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
        if (alive) {
            uint32_t *t = P.tail + at;
            const uint32_t cap = P.tail_cap;
            t[0] = __float_as_uint(L.z); t[cap] = __float_as_uint(L.ux); t[2 * cap] = __float_as_uint(L.uy);
            t[3 * cap] = __float_as_uint(L.uz); t[4 * cap] = __float_as_uint(L.path_lo); t[5 * cap] = __float_as_uint(L.path_hi);
            t[6 * cap] = L.i; t[7 * cap] = L.plo; t[8 * cap] = L.phi; t[9 * cap] = lane_row(L, rows_addr); t[10 * cap] = L.blk;
        }
        TIMELINE(2);
        return;
    }
    if (lane == 0) atomicAdd(&D.draining, 1u);
    const uint32_t give_max = P.drain_give;   // 0: plain drain loop (the launch runs alone)
    // Unlocked peek at the pool: (count << 8) | draining, lane 0's view, so every branch on it is warp-uniform.
    // Refreshed once per iteration, one event stale; every decision is re-made under the lock.
    uint32_t peek = 0u;
    for (;;) {
        if (!alive && L.i != 0u) alive = resolve_lane<IMP, SWEEP>(P, cases, rows, rows_addr, L);
        const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
        const uint32_t n_alive = __popc(alive_mask);
        if (give_max == 0u) {
            if (n_alive == 0u) { TIMELINE(2); break; }
        } else if (n_alive < 32u) {
            const bool take = (peek >> 8) != 0u;
            const bool give = !take && n_alive != 0u && n_alive <= give_max && (peek & 0xffu) > 1u;
            if ((take || give || n_alive == 0u) && pool_try_acquire(D, lane)) {
                const uint32_t c = __shfl_sync(0xffffffffu, lane == 0 ? pool_get(&D.count) : 0u, 0);
                const uint32_t draining = __shfl_sync(0xffffffffu, lane == 0 ? pool_get(&D.draining) : 0u, 0);
                bool leave = false, taken = false;
                if (c != 0u) {   // take over photons another warp left behind
                    const uint32_t k = min(c, 32u - n_alive);
                    const uint32_t rank = __popc(~alive_mask & ((1u << lane) - 1u));
                    if (!alive && rank < k) {
                        const uint32_t e = c - 1u - rank;
                        L.z = __uint_as_float(pool_get(&D.word[0][e])); L.ux = __uint_as_float(pool_get(&D.word[1][e]));
                        L.uy = __uint_as_float(pool_get(&D.word[2][e])); L.uz = __uint_as_float(pool_get(&D.word[3][e]));
                        L.path_lo = __uint_as_float(pool_get(&D.word[4][e]));
                        L.path_hi = __uint_as_float(pool_get(&D.word[5][e]));
                        L.i = pool_get(&D.word[6][e]); L.plo = pool_get(&D.word[7][e]); L.row_addr = pool_get(&D.word[8][e]);
                        L.blk = pool_get(&D.word[9][e]);
                        if (SWEEP) L.phi = pool_get(&D.word[10][e]);
                        taken = true;
                    }
                    __syncwarp();
                    if (lane == 0) pool_put(&D.count, c - k);
                } else if (n_alive == 0u) {   // nothing carried, nothing waiting: this warp is done
                    if (lane == 0) atomicSub(&D.draining, 1u);
                    leave = true;
                } else if (n_alive <= give_max && draining > 1u && c + n_alive <= (uint32_t)POOL_CAP) {
                    // hand the photons over to the warps that stay and leave
                    if (alive) {
                        const uint32_t e = c + __popc(alive_mask & ((1u << lane) - 1u));
                        pool_put(&D.word[0][e], __float_as_uint(L.z)); pool_put(&D.word[1][e], __float_as_uint(L.ux));
                        pool_put(&D.word[2][e], __float_as_uint(L.uy)); pool_put(&D.word[3][e], __float_as_uint(L.uz));
                        pool_put(&D.word[4][e], __float_as_uint(L.path_lo));
                        pool_put(&D.word[5][e], __float_as_uint(L.path_hi));
                        pool_put(&D.word[6][e], L.i); pool_put(&D.word[7][e], L.plo); pool_put(&D.word[8][e], L.row_addr);
                        pool_put(&D.word[9][e], L.blk);
                        if (SWEEP) pool_put(&D.word[10][e], L.phi);
                    }
                    if (lane == 0) { pool_put(&D.count, c + n_alive); atomicSub(&D.draining, 1u); }
                    leave = true;
                }
                pool_release(D, lane);
                if (leave) { TIMELINE(2); break; }
                if (taken) {   // outside the lock: rebuild the Philox state of the photon just taken over
                    L.pk = philox_walk_constants(L.plo, P.rk);
                    alive = true;
                }
            }
        }
        if (give_max != 0u)
            peek = __shfl_sync(0xffffffffu, lane == 0 ? (pool_get(&D.count) << 8) | min(pool_get(&D.draining), 255u) : 0u, 0);
        if (alive) alive = P.drain_latency ? group_latency<IMP, SWEEP>(P, rows, rows_addr, L) : group<IMP, true, SWEEP>(P, rows, rows_addr, L);
    }
}

#ifdef MC3D_TIMELINE
extern "C" int mc3d_debug_timeline(unsigned long long *out, int n_words)
{
    return (int)cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * (size_t)n_words);
}
#endif

// ---- launch helper (called from the host runtime) ------------------------------------------------------------

size_t walk_smem_bytes(int n_rows, uint32_t n_cases, int block_threads)
{
    return tables_bytes(n_rows, n_cases) + (block_threads / 32) * sizeof(WarpRing) + sizeof(DrainPool);
}

template <bool IMP, bool SWEEP, int BLOCK, int MIN_BLOCKS>
static cudaError_t launch_one(const WalkParams &P, int grid, cudaStream_t stream, int *occupancy)
{
    const size_t smem = walk_smem_bytes(P.n_rows, P.n_cases, BLOCK);
    auto kern = walk_kernel<IMP, SWEEP, BLOCK, MIN_BLOCKS>;
    if (smem > 48 * 1024) {   // only large SSP tables need the opt-in (the default table + rings + pool is ~14 KB)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (occupancy) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy, kern, BLOCK, smem);
    kern<<<grid, BLOCK, smem, stream>>>(P);
    return cudaGetLastError();
}

template <int BLOCK, int MIN_BLOCKS>
static cudaError_t launch_variant(const WalkParams &P, bool impurity, int grid, cudaStream_t stream, int *occupancy)
{
    if (P.n_cases) {
        if (impurity) return launch_one<true, true, BLOCK, MIN_BLOCKS>(P, grid, stream, occupancy);
        return launch_one<false, true, BLOCK, MIN_BLOCKS>(P, grid, stream, occupancy);
    }
    if (impurity) return launch_one<true, false, BLOCK, MIN_BLOCKS>(P, grid, stream, occupancy);
    return launch_one<false, false, BLOCK, MIN_BLOCKS>(P, grid, stream, occupancy);
}

// block_threads in {128, 256, 512}; blocks_per_sm is the occupancy target the variant is compiled for (64 / 48
// registers per thread for <= 32 / 40 resident warps per SM; 32 warps is the default: the loop is bound by the
// issue port and more resident warps do not raise its rate).
// With `occupancy` non-null nothing is launched; the resident blocks per SM are returned through it.
cudaError_t launch_walk(const WalkParams &P, bool impurity, int block_threads, int blocks_per_sm, int grid,
                        cudaStream_t stream, int *occupancy)
{
    const int warps_per_sm = block_threads / 32 * blocks_per_sm;
    if (block_threads == 128) {
        if (warps_per_sm <= 32) return launch_variant<128, 8>(P, impurity, grid, stream, occupancy);
        return launch_variant<128, 10>(P, impurity, grid, stream, occupancy);
    }
    if (block_threads == 256) {
        if (warps_per_sm <= 32) return launch_variant<256, 4>(P, impurity, grid, stream, occupancy);
        return launch_variant<256, 5>(P, impurity, grid, stream, occupancy);
    }
    if (block_threads == 512) return launch_variant<512, 2>(P, impurity, grid, stream, occupancy);
    return cudaErrorInvalidValue;
}

}  // namespace mc3d
