#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- statistical golden data from the UNMODIFIED reference at 10^6 photons
(BASELINE.json configs[1]: spheres, HG, default wavelength / grain size / zenith), for the production-mode
3-sigma tests.  Runs K single-rank reference processes (the reference's MPI ranks are independent apart from one
scatter and one gather) with distinct seeds and stores per-wavelength outcome counts, the 137-bin BRF zenith
histogram of reflected photons (post_processing.py:73-76), an n_scat histogram and path-length moments.

    python oracle/make_golden_stats.py [n_photon_total] [n_proc] [config]     # config c2: ~48 core-minutes at 10^6

Configs: ``c2`` (default; stats_c2_reference.npz), ``slab_lb`` (finite slab tau_tot = 3 over a Lambertian bottom
R = 0.5: every condition 1-4 occurs, bottom reflections feed back into the walk, monte_carlo3D.py:1238-1262, 1418-1466)
and ``impurity`` (tau_tot = 3, black carbon 1e-5 by mass: the ice / impurity species draw, monte_carlo3D.py:1020-1025,
condition 5), written to stats_<config>_reference.npz.
"""
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

CONFIGS = {
    'c2': dict(wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=15., tau_tot=1e6, Lambertian_bottom=True,
               Lambertian_reflectance=0.5, fixture='spectral', n_theta_bins=137, base_seed=7000),
    'slab_lb': dict(wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=15., tau_tot=3.0, Lambertian_bottom=True,
                    Lambertian_reflectance=0.5, fixture='spectral', n_theta_bins=137, base_seed=17000),
    # BASELINE.json configs[3]: visible wavelength, weakly absorbing ice, large grains -- thousands of scatterings per
    # photon (mean ~2700, walks beyond 10^5 events): the long-walk regime (renormalisation, path accumulation in fp32)
    'vis': dict(wvl0=0.53, half_width=0.085, rds_snw=1000., theta_0=15., tau_tot=1e6, Lambertian_bottom=True,
                Lambertian_reflectance=0.5, fixture='const-vis', n_theta_bins=137, base_seed=37000),
    'impurity': dict(wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=30., tau_tot=3.0, imp_cnc=1e-5,
                     Lambertian_bottom=True, Lambertian_reflectance=0.5, fixture='spectral', n_theta_bins=137,
                     base_seed=27000),
}
CFG = dict(CONFIGS['c2'])


def work(args):
    rank, n, optics, cfg = args
    CFG.clear()
    CFG.update(cfg)
    from oracle import ref_shim
    r = ref_shim.run_reference(n, CFG['wvl0'], CFG['half_width'], CFG['rds_snw'], CFG['theta_0'],
                               seed=CFG['base_seed'] + rank, optics_dir=optics,
                               model_kwargs=dict(tau_tot=CFG['tau_tot'], imp_cnc=CFG.get('imp_cnc', 0.0)),
                               run_kwargs=dict(Lambertian_bottom=CFG['Lambertian_bottom'],
                                               Lambertian_reflectance=CFG['Lambertian_reflectance']), record=False)
    return {k: r[k] for k in ('condition', 'wvl', 'theta_n', 'n_scat', 'path_length')}


def main():
    n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
    n_proc = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    name = sys.argv[3] if len(sys.argv) > 3 else 'c2'
    CFG.clear()
    CFG.update(CONFIGS[name])
    from monte_carlompi_b200 import ssp_fixtures
    optics = os.path.join(tempfile.mkdtemp(prefix='mc3d_optics_'), CFG['fixture'])
    ssp_fixtures.write_optics_dir(optics, CFG['fixture'], (int(CFG['rds_snw']),))
    n_chunks = n_proc * 4
    sizes = [len(c) for c in np.array_split(np.arange(n_total), n_chunks)]
    with mp.Pool(n_proc) as pool:
        parts = pool.map(work, [(k, sizes[k], optics, dict(CFG)) for k in range(n_chunks)], chunksize=1)
    cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    k = np.rint(cat['wvl'] * 100).astype(np.int64)
    k_lo, k_hi = int(k.min()), int(k.max())
    nb = CFG['n_theta_bins']
    counts = np.zeros((k_hi - k_lo + 1, 8), np.int64)
    brf = np.zeros((k_hi - k_lo + 1, nb), np.int64)
    for kk in range(k_lo, k_hi + 1):
        m = k == kk
        counts[kk - k_lo, 0] = m.sum()
        for c in range(1, 6):
            counts[kk - k_lo, c] = (m & (cat['condition'] == c)).sum()
        brf[kk - k_lo] = np.histogram(cat['theta_n'][m & (cat['condition'] == 1)], bins=nb, range=(0., np.pi / 2))[0]
    refl = cat['condition'] == 1
    out = os.path.join(ROOT, 'tests', 'golden', 'stats_%s_reference.npz' % name)
    np.savez_compressed(out, config=np.array(repr(dict(CFG, n_photon=n_total))), k_first=k_lo, counts=counts, brf=brf,
                        n_scat_hist=np.bincount(np.minimum(cat['n_scat'], 4095), minlength=4096),
                        n_scat_sq_sum=(cat['n_scat'].astype(np.float64) ** 2).sum(),
                        n_scat_log2_hist=np.bincount(np.floor(np.log2(np.maximum(cat['n_scat'], 1))).astype(np.int64), minlength=32),
                        n_scat_sum=cat['n_scat'].sum(), path_sum=cat['path_length'].sum(),
                        path_sq_sum=(cat['path_length'] ** 2).sum(),
                        path_sum_reflected=cat['path_length'][refl].sum())
    print('wrote', out, 'n', n_total, 'cond', np.bincount(cat['condition'], minlength=6)[1:], 'mean n_scat',
          cat['n_scat'].mean())


if __name__ == '__main__':
    main()
