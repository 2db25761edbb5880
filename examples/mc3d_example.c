/* mc3d_example.c -- the C ABI of libmc3d.so from plain C (no Python, no torch): the known-answer case of the
 * reference's test() (monte_carlo3D.py:1849-1866: tau_tot 2, ssa 0.9, g 0.75, normal incidence, black bottom;
 * van de Hulst 1980 / Wang et al. 1995: albedo ~0.09739, total transmittance ~0.66096).
 *
 *   gcc -std=c99 -I include examples/mc3d_example.c -L monte_carlompi_b200 -l:libmc3d.so \
 *       -Wl,-rpath,$PWD/monte_carlompi_b200 -o mc3d_example && ./mc3d_example [n_photon]
 *
 * Exit code 0 = ran and matched the known answer, 2 = no CUDA device (the library has no CPU fallback), 1 = error. */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "mc3d.h"

int main(int argc, char **argv)
{
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 2000000ull;
    if (mc3d_abi_version() != MC3D_ABI_VERSION) {
        fprintf(stderr, "ABI mismatch: library %d, header %d\n", mc3d_abi_version(), MC3D_ABI_VERSION);
        return 1;
    }
    mc3d_ctx *ctx = NULL;
    int rc = mc3d_create(&ctx, NULL, 1);
    if (rc == MC3D_ENODEVICE) {
        fprintf(stderr, "mc3d_create: %s\n", mc3d_last_error());
        return 2;
    }
    if (rc != MC3D_OK) {
        fprintf(stderr, "mc3d_create failed (%d): %s\n", rc, mc3d_last_error());
        return 1;
    }
    /* one wavelength row: 0.50 um, monochromatic (sigma = 0) */
    mc3d_ssp_row row = {0.50, 0.9, 0.3, 0.75, 16.4, 0.0};
    mc3d_params p = {0.0, 2.0, 300.0, 1.0, 0.50, 0.0, 50, 0u /* black bottom */, 18, 0};
    uint64_t tally[MC3D_N_COND + 18];
    mc3d_stats st;
    /* a first, small call: the CUDA runtime loads each kernel on its first launch (not part of any later call's time) */
    rc = mc3d_run(ctx, &p, &row, 1, 1ull, 0, 1000, NULL, tally, &st);
    if (rc == MC3D_OK) rc = mc3d_run(ctx, &p, &row, 1, 20190603ull, 0, n, NULL, tally, &st);
    if (rc != MC3D_OK) {
        fprintf(stderr, "mc3d_run failed (%d): %s\n", rc, mc3d_last_error());
        mc3d_destroy(ctx);
        return 1;
    }
    const double refl = (double)tally[MC3D_COND_REFLECTED] / (double)n;
    const double trans = (double)(tally[MC3D_COND_DIFFUSE_TRANSMITTED] + tally[MC3D_COND_DIRECT_TRANSMITTED]) / (double)n;
    printf("photons %llu events %llu kernel %.3f ms (%d SMs, grid %d x %d)\n", (unsigned long long)st.n_photon,
           (unsigned long long)st.n_events, st.kernel_ms, st.sm_count, st.grid_blocks, st.block_threads);
    printf("albedo %.5f (known answer 0.09739)  transmittance %.5f (0.66096)\n", refl, trans);
    mc3d_destroy(ctx);
    const double tol = 4.0 * sqrt(0.25 / (double)n) + 3e-4;
    return (fabs(refl - 0.09739) < tol && fabs(trans - 0.66096) < tol && tally[0] == n) ? 0 : 1;
}
