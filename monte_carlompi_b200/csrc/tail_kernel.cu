// tail_kernel.cu -- the end of a call that runs alone on the GPU.
//
// When the walk kernel has handed out its last photon, every lane still carries one, and the launch ends when the
// longest of those walks ends: a walk is a sequential chain, one event after the other (reference
// monte_carloMPI/monte_carlo3D.py:1212-1466, the `while` loop of one photon).  With other calls in flight that time is
// filled by their work (walk_kernel.cu: drain consolidation); a call that runs alone just waits for it -- for a
// 10^6-photon launch of config C2 about as long as everything before (profiles/r02_lone_launch_timeline.log), for
// visible wavelengths (10^6-event walks) far longer.  So such a call's walk kernel stops at that point and leaves the
// walking photons in a list (WalkParams::tail), and this kernel finishes them:
//
//   1. dense warps: 32 photons per warp again, the latency-oriented group of four events (walk_device.cuh:
//      group_latency), lanes resolved as soon as they stop;
//   2. helper lanes: once a warp is down to <= 4 photons, 28+ of its lanes have nothing to do while the others crawl
//      along their chains at one event per ~125 ns -- three quarters of which is not the chain (the rotation) but the
//      Philox blocks, the Henyey-Greenstein deflection, the azimuth and the free path of the event, which depend on
//      the photon's stream only.  Each lane therefore PREPARES one of the next eight events (two groups) of one of
//      the warp's photons (prepare_event: same code, same words, same values), publishes it in shared memory, and the
//      photon's own lane only applies the eight rotations and moves (apply_event) with the attention predicate after
//      each.  An event that needs attention ends the step; the photon is resolved and continues with its next group
//      as always (mc3d_device.cuh: the random-number layout), so the results are bit-identical (tested).
#include <algorithm>

#include "walk_device.cuh"

namespace mc3d {

constexpr int COOP_MAX = 4;      // photons per warp at which the lanes turn into helpers (8 prepared events each)
constexpr int COOP_EVENTS = 8;   // = 32 / COOP_MAX: two groups of the walk stream

struct __align__(16) CoopEntry {
    float ct, st2, cp, sp;       // Prepared, 16-byte load
    float dtau;
    uint32_t key;
    uint32_t pad[2];
};

template <bool IMP, bool SWEEP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) tail_kernel(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    const DevCase *cases = staged_cases(P, smem_raw);
    CoopEntry *coop_all = reinterpret_cast<CoopEntry *>(smem_raw + ((tables_bytes(P.n_rows, P.n_cases) + 15) & ~(size_t)15));
    stage_tables(P, smem_raw, BLOCK);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t rows_addr = shared_address(rows);
    CoopEntry *coop = coop_all + (threadIdx.x >> 5) * 64;   // two buffers of 32 prepared events per warp
    const uint32_t n_tail = *P.n_tail;
    const uint32_t first = (blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5)) * 32u;
    if (first >= n_tail) return;   // whole warps

    Lane L;
    L.z = 0.f; L.ux = 0.f; L.uy = 0.f; L.uz = -1.f; L.path_lo = 0.f; L.path_hi = 0.f;
    L.i = 0u; L.blk = 0u; L.plo = 0u; L.phi = 0u; L.row_addr = rows_addr; L.key = 0u; L.imp = false;
    L.pk.pB = L.pk.pC = L.pk.pD = 0u;
    bool alive = first + lane < n_tail;
    if (alive) {
        const uint32_t *t = P.tail + first + lane;
        const uint32_t cap = P.tail_cap;
        L.z = __uint_as_float(t[0]); L.ux = __uint_as_float(t[cap]); L.uy = __uint_as_float(t[2 * cap]);
        L.uz = __uint_as_float(t[3 * cap]); L.path_lo = __uint_as_float(t[4 * cap]); L.path_hi = __uint_as_float(t[5 * cap]);
        L.i = t[6 * cap]; L.plo = t[7 * cap]; L.phi = t[8 * cap];
        L.row_addr = rows_addr + t[9 * cap] * (uint32_t)sizeof(DevRow); L.blk = t[10 * cap];
        L.pk = philox_walk_constants(L.plo, P.rk);
    }

    for (;;) {
        if (!alive && L.i != 0u) alive = resolve_lane<IMP, SWEEP>(P, cases, rows, rows_addr, L);
        const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
        const uint32_t n_alive = __popc(alive_mask);
        if (n_alive == 0u) break;
        if (IMP || n_alive > (uint32_t)COOP_MAX) {   // (impurity runs draw the species between events: no helpers)
            if (alive) alive = group_latency<IMP, SWEEP>(P, rows, rows_addr, L);
            continue;
        }
        // ---- helper lanes: lane (q, s) prepares event s of eight (two groups) of the warp's q-th photon.  The step is
        // software-pipelined: while the photons' own lanes apply the eight events of step t -- a dependent chain of
        // ~50 cycles per event that leaves most issue slots empty -- all lanes prepare step t+1 on the assumption
        // that no event of step t needs attention (then the photon simply continues six blocks further on).  One
        // basic block, no divergence: lanes without a photon run the rotations on their idle state under a false mask.
        const uint32_t q = lane >> 3, s = lane & 7u;
        uint32_t m = alive_mask;
        for (uint32_t j = 0; j < q; ++j) m &= m - 1u;
        const uint32_t src = (m ? __ffs(m) : __ffs(alive_mask)) - 1u;   // fewer than four photons: spare lanes shadow the first
        const uint32_t my_q = min((uint32_t)__popc(alive_mask & ((1u << lane) - 1u)), (uint32_t)COOP_MAX - 1u);
        const uint32_t o_plo = __shfl_sync(0xffffffffu, L.plo, src), o_phi = __shfl_sync(0xffffffffu, L.phi, src);
        const uint32_t o_row = __shfl_sync(0xffffffffu, L.row_addr, src);
        uint32_t o_blk = __shfl_sync(0xffffffffu, L.blk, src);
        const uint32_t phi = SWEEP ? o_phi : (uint32_t)(P.c.id0 >> 32);
        const PhiloxWalkConst pk = philox_walk_constants(o_plo, P.rk);
        const HotRow Ho = load_hot_row<SWEEP>(P, o_row);
        const HotRow H = load_hot_row<SWEEP>(P, L.row_addr);
        // slot e of a group uses words 3e .. 3e+2 of the group's twelve: blocks (3e) >> 2 and (3e + 2) >> 2
        const uint32_t e = s & 3u, boff0 = GROUP_BLOCKS * (s >> 2) + ((3u * e) >> 2), boff1 = GROUP_BLOCKS * (s >> 2) + ((3u * e + 2u) >> 2);
        auto prepare_at = [&](uint32_t blk) {
            const uint4 va = philox_walk(blk + boff0, phi, pk, P.rk);
            const uint4 vb = philox_walk(blk + boff1, phi, pk, P.rk);
            const uint32_t w0 = e == 0u ? va.x : e == 1u ? va.w : e == 2u ? va.z : vb.y;
            const uint32_t w1 = e == 0u ? va.y : e == 1u ? vb.x : e == 2u ? va.w : vb.z;
            const uint32_t w2 = e == 0u ? va.z : e == 1u ? vb.y : e == 2u ? vb.x : vb.w;
            return prepare_event(Ho, w0, w1, w2);
        };
        auto publish = [&](CoopEntry *buf, const Prepared &E) {
            *reinterpret_cast<float4 *>(&buf[lane].ct) = make_float4(E.ct, E.st2, E.cp, E.sp);
            *reinterpret_cast<uint2 *>(&buf[lane].dtau) = make_uint2(__float_as_uint(E.dtau), E.key);
        };
        uint32_t cur = 0u;
        publish(coop, prepare_at(o_blk));
        __syncwarp();
        for (;;) {
            const Prepared En = prepare_at(o_blk + 2u * GROUP_BLOCKS);   // step t+1, speculatively
            const CoopEntry *mine = coop + cur * 32u + my_q * COOP_EVENTS;
            Prepared ev[COOP_EVENTS];
#pragma unroll
            for (int k = 0; k < COOP_EVENTS; ++k) {
                const float4 a = *reinterpret_cast<const float4 *>(&mine[k].ct);
                const uint2 b = *reinterpret_cast<const uint2 *>(&mine[k].dtau);
                ev[k].ct = a.x; ev[k].st2 = a.y; ev[k].cp = a.z; ev[k].sp = a.w;
                ev[k].dtau = __uint_as_float(b.x); ev[k].key = b.y;
            }
            bool go = alive, second = false;
#pragma unroll
            for (int k = 0; k < COOP_EVENTS; ++k) {
                if (k == 4) second = go;   // the first group ran to its end: the photon enters the second one
                Lane N = L;
                N.i += 1u;
                apply_event(N, ev[k]);
                const bool att = needs_attention(H.neg_tau, N, H.t_hot);
                L.z = go ? N.z : L.z; L.ux = go ? N.ux : L.ux; L.uy = go ? N.uy : L.uy; L.uz = go ? N.uz : L.uz;
                L.path_lo = go ? N.path_lo : L.path_lo; L.key = go ? N.key : L.key; L.i = go ? N.i : L.i;
                go = go && !att;
            }
            L.blk += alive ? (second ? 2u * GROUP_BLOCKS : GROUP_BLOCKS) : 0u;
            const uint32_t stopped = __ballot_sync(0xffffffffu, alive && !go);
            alive = go;
            if (stopped) break;   // somebody needs attention: resolve, then set the helpers up again
            o_blk += 2u * GROUP_BLOCKS;
            cur ^= 1u;
            publish(coop + cur * 32u, En);
            __syncwarp();
        }
        __syncwarp();
    }
}

template <bool IMP, bool SWEEP>
static cudaError_t launch_tail_variant(const WalkParams &P, int lanes, cudaStream_t stream)
{
    constexpr int BLOCK = 128;
    const size_t smem = ((tables_bytes(P.n_rows, P.n_cases) + 15) & ~(size_t)15) + (BLOCK / 32) * 64 * sizeof(CoopEntry);
    auto kern = tail_kernel<IMP, SWEEP, BLOCK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int grid = std::max(1, (lanes + BLOCK - 1) / BLOCK);
    kern<<<grid, BLOCK, smem, stream>>>(P);
    return cudaGetLastError();
}

// `lanes`: lanes of the walk kernel's grid (an upper bound of the list's length; warps beyond it leave at once).
cudaError_t launch_tail(const WalkParams &P, bool impurity, int lanes, cudaStream_t stream)
{
    if (P.n_cases) return impurity ? launch_tail_variant<true, true>(P, lanes, stream) : launch_tail_variant<false, true>(P, lanes, stream);
    return impurity ? launch_tail_variant<true, false>(P, lanes, stream) : launch_tail_variant<false, false>(P, lanes, stream);
}

}  // namespace mc3d
