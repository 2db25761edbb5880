// replay_kernel.cu -- fp64 replay mode: the walk of reference monte_carloMPI/monte_carlo3D.py:1111-1490 driven by
// the reference's OWN recorded random stream (np.random.rand / uniform values in consumption order), one thread
// per photon.  A correctness tool, not a throughput path: it must reproduce the reference's per-photon
// condition and n_scat exactly and its angles / path length to ~1e-9 (libm-vs-CUDA ulp differences only).
//
// Compiled with -fmad=false: the reference's Python scalar arithmetic never fuses a multiply with an add, and the
// operation order below is the reference's.  numpy scalar `x**2` (libm pow) is evaluated as x*x.
#include "mc3d_device.cuh"

namespace mc3d {

struct Stream {
    const double *v;
    long long n, pos;
    bool exhausted;
    __device__ double next()
    {
        if (pos >= n) { exhausted = true; ++pos; return 0.5; }
        return v[pos++];
    }
};

// monte_carlo3D.py:790-800
__device__ double henyey_greenstein2(double g, double r)
{
    if (g == 0) return 1 - 2 * r;
    const double g2 = g * g;
    const double q = (1 - g2) / (1 - g + 2 * g * r);
    return (1. / (2. * g)) * (1 + g2 - q * q);
}

__global__ void __launch_bounds__(128) replay_kernel(const ReplayParams P)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_photon) return;
    constexpr double PI = 3.141592653589793, TWO_PIE = 2 * 3.141592653589793;
    const bool lambert_bottom = P.flags & 1u, lambert_surface = P.flags & 2u;
    const double g = P.g[p], ssa_ice = P.ssa_ice[p], ssa_imp = P.ssa_imp[p], p_ext_imp = P.p_ext_imp[p];
    const double ext_cff_mss = P.ext_cff_mss[p];
    const double ext_cff = ext_cff_mss * P.rho_snw;                      // 1355-1356
    Stream S{P.stream + P.offsets[p], P.offsets[p + 1] - P.offsets[p], 0, false};

    double mux_0 = sin(P.theta0_rad), muy_0 = 0, muz_0 = -cos(P.theta0_rad);   // 1121-1123
    double muz2_0 = muz_0 * muz_0;
    double mux_n = 0, muy_n = 0, muz_n = 0;
    double z_prev = 0, z = 0, path_length = 0;
    bool bottom_reflection = false;
    int condition = 0;
    long long i = 0;

    while (condition == 0) {
        i += 1;
        double r1 = 0, u_phi = 0, u_tau, u_ssa, u_ext;
        if (i == 1) {                                                   // initial_pdfs, 1035-1038
            u_tau = P.init_draws[3ull * p]; u_ssa = P.init_draws[3ull * p + 1]; u_ext = P.init_draws[3ull * p + 2];
        } else {                                                        // populate_pdfs, 915-921, 1014-1023
            r1 = S.next(); u_phi = S.next(); u_tau = S.next(); u_ssa = S.next(); u_ext = S.next();
        }
        if (S.exhausted) break;
        double dtau = -log(u_tau);
        if (lambert_surface && i == 1) dtau = 0;                        // 1228-1229

        double costheta = 1, sintheta = 0;
        if (i == 1) {
        } else if (lambert_surface || bottom_reflection) {              // 1238-1250
            mux_0 = 0.; muy_0 = 0.; muz_0 = 1.;
            for (;;) {
                const double u_theta = S.next(), r = S.next();
                if (S.exhausted) break;
                const double theta_rand = 0.0 + (PI / 2 - 0.0) * u_theta;
                if (r < 2 * sin(theta_rand) * cos(theta_rand)) {
                    costheta = cos(theta_rand);
                    sintheta = sqrt(1 - costheta * costheta);
                    break;
                }
            }
            if (S.exhausted) break;
        } else {                                                        // 1252-1253
            costheta = henyey_greenstein2(g, r1);
            sintheta = sqrt(1 - costheta * costheta);
        }

        if (i > 1) {                                                    // 1255-1285
            const double phi = u_phi * TWO_PIE;
            const double cosphi = cos(phi), sinphi = sin(phi);
            if (muz_0 == 1) {
                mux_n = sintheta * cosphi; muy_n = sintheta * sinphi; muz_n = costheta;
            } else if (muz_0 == -1) {
                mux_n = sintheta * cosphi; muy_n = -sintheta * sinphi; muz_n = -costheta;
            } else {
                const double den = sqrt(1 - muz2_0);
                mux_n = (sintheta * (mux_0 * muz_0 * cosphi - muy_0 * sinphi)) / den + mux_0 * costheta;
                muy_n = (sintheta * (muy_0 * muz_0 * cosphi + mux_0 * sinphi)) / den + muy_0 * costheta;
                muz_n = -den * sintheta * cosphi + muz_0 * costheta;
            }
            bottom_reflection = false;
        } else {                                                        // 1344-1347
            mux_n = mux_0; muy_n = muy_0; muz_n = muz_0;
        }

        z_prev = z;
        z = z_prev + dtau * muz_n;                                      // 1352
        if (i > 1) {                                                    // 1364-1369
            mux_0 = mux_n; muy_0 = muy_n; muz_0 = muz_n;
            muz2_0 = muz_0 * muz_0;
        }
        path_length += dtau / ext_cff;                                  // 1372

        int ext_state;
        double ssa_event;
        if (u_ext > p_ext_imp) { ext_state = 1; ssa_event = ssa_ice; }  // 1375-1383
        else { ext_state = 2; ssa_event = ssa_imp; }
        if (lambert_surface) ssa_event = P.r_lambert;                   // 1385-1387

        if (z > 0) {                                                    // 1390-1397
            condition = 1;
            path_length += -((z * dtau) / ((z - z_prev) * ext_cff));
        } else if (z < -P.tau_tot) {                                    // 1399-1459
            path_length += -(((z + P.tau_tot) * dtau) / ((z - z_prev) * ext_cff));
            const double dtau_correction = -(((z + P.tau_tot) / (z_prev - z)) * dtau);
            z = z - (muz_n * dtau_correction);
            const int exit_cond = (i == 1) ? 3 : 2;
            if (lambert_bottom) {
                const double reflectance_rand = S.next();
                if (S.exhausted) break;
                if (reflectance_rand <= P.r_lambert) bottom_reflection = true;
                else condition = exit_cond;
            } else {
                condition = exit_cond;
            }
        } else if (u_ssa >= ssa_event) {                                // 1461-1466
            condition = (ext_state == 1) ? 4 : 5;
        }
    }

    P.condition[p] = condition;
    P.wvn[p] = 1. / P.wvl[p];                                           // 1468
    P.theta_n[p] = acos(muz_0);                                         // 1469
    double phi_n;
    if (i == 1) phi_n = 0.;                                             // 1472-1485
    else if (mux_0 > 0 && muy_0 > 0) phi_n = atan(muy_0 / mux_0);
    else if (mux_0 < 0 && muy_0 > 0) phi_n = atan(muy_0 / mux_0) + PI;
    else if (mux_0 < 0 && muy_0 < 0) phi_n = atan(muy_0 / mux_0) + PI;
    else if (mux_0 > 0 && muy_0 < 0) phi_n = atan(muy_0 / mux_0) + TWO_PIE;
    else phi_n = nan("");   // the reference raises UnboundLocalError here
    P.phi_n[p] = phi_n;
    P.n_scat[p] = i - 1;                                                // 1487
    P.path_length[p] = path_length;
    P.snow_depth[p] = P.tau_tot / (ext_cff_mss * P.rho_snw);            // 1612
    P.consumed[p] = S.pos;
}

cudaError_t launch_replay(const ReplayParams &P, cudaStream_t stream)
{
    const int block = 128;
    const int grid = (int)((P.n_photon + block - 1) / block);
    replay_kernel<<<grid > 0 ? grid : 1, block, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace mc3d
