// init_kernel.cu -- the per-photon prologue, one thread per photon, fully coalesced.
//
//   * wavelength draw: wvls = np.around(np.random.normal(wvl0, scale), 2)          reference monte_carlo3D.py:1515-1520
//     (Box-Muller on the photon's own Philox block; the rounded value is an index into the SSP table)
//   * Lambertian_surface mode (monte_carlo3D.py:1228-1229, 1238-1250, 1385-1387): the snow is replaced by a
//     Lambertian reflector, every photon ends after one or two events, so it is finished right here
//   * first event: the three draws of initial_pdfs (monte_carlo3D.py:1035-1038), no deflection (1232-1237),
//     move, direct-transmission / Lambertian-bottom / first-extinction absorption tests (1399-1466)
//
// Photons that are still walking after event 1 are appended to the `fresh` list (warp-aggregated append: one
// atomicAdd per warp); the walk kernel's lanes pick them up from there.  Photons that end on event 1 store their
// raw record here.  Keeping this code out of the walk kernel is what lets the walk kernel run at 48 registers per
// thread (40 resident warps per SM) without spilling in its event loop.
#include <algorithm>

#include "walk_device.cuh"

namespace mc3d {

// One photon in Lambertian_surface mode.  Event 1 does not move (dtau = 0, monte_carlo3D.py:1228-1229) and is
// absorbed with probability 1 - R (ssa_event = R, 1385-1387); every later event re-emits the photon from the
// surface with the cosine law (1238-1250) and almost surely leaves through the top on event 2.
template <bool IMP>
__device__ __noinline__ void lambert_surface_photon(const WalkParams &P, const DevRow &R, uint32_t pid, uint32_t row,
                                                    uint32_t plo, uint32_t phi, uint4 w, bool imp)
{
    float z = 0.0f, path = 0.0f, ux = P.mu0x, uy = 0.0f, uz = P.mu0z;
    uint32_t i = 1u, cond = ALIVE;
    for (;;) {
        // termination chain for event i: z > 0, (z < -tau_tot cannot happen), absorbed by the surface
        if (z > 0.0f) {
            path -= __fdividef(z, uz);
            cond = 1u;
        } else if (w.w > P.surf_t_hi || (w.w == P.surf_t_hi && (w.y & 0xffu) >= P.surf_t_lo)) {
            cond = imp ? 5u : 4u;
        }
        if (cond != ALIVE) break;
        i += 1u;
        w = philox4x32_10(i, TAG_EVENT, plo, phi, P.rk);
        float ct, st;
        for (uint32_t j = 0;; ++j) {
            const uint4 a = philox4x32_10(i, TAG_LAMBERT | ((1u + (j >> 1)) << 8), plo, phi, P.rk);
            const float u_t = u32_to_unit((j & 1u) ? a.z : a.x);
            const float r1 = u32_to_unit((j & 1u) ? a.w : a.y);
            float s_, c_;
            sincosf(1.5707963267948966f * u_t, &s_, &c_);
            if (r1 < 2.0f * s_ * c_) { ct = c_; st = s_; break; }
        }
        float cp, sp;
        azimuth(w.y, cp, sp);
        ux = st * cp; uy = st * sp; uz = ct;
        const float dt = free_path(w.z);
        z = fmaf(dt, ct, z);
        path += dt;
        imp = IMP ? species_is_impurity(P, R, i, plo, phi) : false;
    }
    store_raw(P, pid, ux, uy, uz, path, i - 1u, cond, row);
}

template <bool IMP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) init_kernel(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevRow *rows = reinterpret_cast<DevRow *>(smem_raw);
    for (int k = threadIdx.x; k < P.n_rows * (int)(sizeof(DevRow) / 4); k += BLOCK)
        reinterpret_cast<uint32_t *>(rows)[k] = reinterpret_cast<const uint32_t *>(P.rows)[k];
    __syncthreads();
    const uint32_t rows_addr = shared_address(rows);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t phi = (uint32_t)(P.photon_begin >> 32);
    // whole warps iterate together so that the ballot below is always executed by all 32 lanes
    for (uint32_t base = (blockIdx.x * BLOCK + (threadIdx.x & ~31u)); base < P.n_photon; base += gridDim.x * BLOCK) {
        const uint32_t pid = base + lane;
        bool survive = false, continue_outer = false;
        float dtau = 0.0f;
        uint32_t row = 0;
        if (pid < P.n_photon) {
            const uint32_t plo = (uint32_t)P.photon_begin + pid;
            const uint4 wv = philox4x32_10(0u, TAG_WAVELENGTH, plo, phi, P.rk);
            const float zn = sqrtf(-2.0f * logf(u32_to_unit(wv.x))) * cospif(2.0f * u32_to_unit(wv.y));
            const int r = (int)rint(P.wvl0_x100 + P.sigma_x100 * (double)zn) - P.k_first;
            row = (uint32_t)max(0, min(P.n_rows - 1, r));
            const DevRow &R = rows[row];
            const uint4 w = philox4x32_10(1u, TAG_EVENT, plo, phi, P.rk);
            const bool imp = IMP ? species_is_impurity(P, R, 1u, plo, phi) : false;
            if (P.lambert_surface) {
                lambert_surface_photon<IMP>(P, R, pid, row, plo, phi, w, imp);
                continue_outer = true;
            }
            dtau = free_path(w.z);
            const float z1 = dtau * P.mu0z;
            survive = !continue_outer;
            if (continue_outer) {
            } else if (z1 < P.neg_tau_tot || w.w >= (imp ? R.ti_hi : R.t_hi)) {
                Lane L;
                L.z = z1; L.ux = P.mu0x; L.uy = 0.0f; L.uz = P.mu0z; L.i = 1u; L.path_lo = dtau; L.path_hi = 0.0f;
                L.plo = plo; L.row_addr = rows_addr + row * (uint32_t)sizeof(DevRow); L.w3 = w.w; L.imp = imp;
                L.pk = philox_event_constants(plo, P.rk);
                bool alive = resolve_lane<IMP>(P, rows, rows_addr, L);
                if (alive && L.i != 1u) {
                    // reflected off the Lambertian bottom on its first step and still alive after event 2: it no
                    // longer has the "fresh photon" state, so it is walked to completion here (thin slabs only)
                    do {
                        alive = event<IMP>(P, rows, rows_addr, L);
                        if (!alive) alive = resolve_lane<IMP>(P, rows, rows_addr, L);
                    } while (alive);
                }
                survive = alive;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, survive);
        uint32_t slot0 = 0;
        if (lane == 0 && m) slot0 = atomicAdd(P.n_fresh, (uint32_t)__popc(m));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (survive) {
            Fresh f;
            f.pid = pid; f.row = row; f.dtau = dtau; f.pad = 0u;
            *reinterpret_cast<uint4 *>(P.fresh + slot0 + __popc(m & ((1u << lane) - 1u))) = *reinterpret_cast<uint4 *>(&f);
        }
    }
}

cudaError_t launch_init(const WalkParams &P, bool impurity, int sm_count, cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    const size_t smem = (size_t)P.n_rows * sizeof(DevRow);
    const long long want = ((long long)P.n_photon + BLOCK - 1) / BLOCK;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * 8));
    if (impurity) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(init_kernel<true, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        init_kernel<true, BLOCK><<<grid, BLOCK, smem, stream>>>(P);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(init_kernel<false, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        init_kernel<false, BLOCK><<<grid, BLOCK, smem, stream>>>(P);
    }
    return cudaGetLastError();
}

}  // namespace mc3d
