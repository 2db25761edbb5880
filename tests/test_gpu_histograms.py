"""n_scat / path-length histograms binned on the GPU (post_processing.py:162-223) against np.histogram of the
records of the same run -- through the C ABI and through the driver surface."""
import numpy as np
import pytest

import gpu_util
from monte_carlompi_b200 import engine

pytestmark = pytest.mark.gpu


def _np_hist(x, bins, rng=None):
    return np.histogram(x, bins=bins, range=rng)


@pytest.mark.parametrize('kind,tau,n,bins', [('nir', 1e6, 300000, (200, 1000)), ('nir', 3.0, 100000, (37, 64)),
                                             ('vis', 10.0, 20000, (200, 1000)), ('nir', 1e6, 5000, (20000, 70000))])
def test_abi_histograms_equal_numpy(kind, tau, n, bins):
    ctx = gpu_util.context()
    table = gpu_util.const_table(0.992, 0.89) if kind == 'nir' else gpu_util.const_table(0.999989859099, 0.89, ext=6.6)
    p = engine.make_params(np.pi * 15. / 180., tau, 300., 0.5, 0.5, 0.0, 50, lambert_bottom=True, n_theta_bins=0)
    ctx.set_histograms()
    rec, _, _ = ctx.run(p, table, 11, 0, n, records=True, tally=False)
    ns_min, ns_max, pl_min, pl_max = ctx.extrema(0)
    assert ns_min == rec['n_scat'].min() and ns_max == rec['n_scat'].max()
    assert np.float32(pl_min) == rec['path_length'].min() and np.float32(pl_max) == rec['path_length'].max()
    path_cm = rec['path_length'].astype(np.float64) * 100.
    want_ns, _ = _np_hist(rec['n_scat'], bins[0])
    want_pl, _ = _np_hist(path_cm, bins[1])
    ctx.set_histograms(bins[0], (float(ns_min), float(ns_max)), bins[1], (pl_min * 100., pl_max * 100.), 100.)
    try:
        ctx.run(p, table, 11, 0, n, records=False, tally=False)
        ns, pl = ctx.histograms(0)
        assert ns.sum() == n and pl.sum() == n
        assert np.array_equal(ns, want_ns.astype(np.uint64))
        assert np.array_equal(pl, want_pl.astype(np.uint64))
        # a narrower range drops the outliers exactly like np.histogram(range=...)
        lo, hi = np.percentile(path_cm, [10, 90])
        ctx.set_histograms(0, (0., 1.), 50, (lo, hi), 100.)
        ctx.run(p, table, 11, 0, n, records=False, tally=False)
        ns2, pl2 = ctx.histograms(0)
        assert ns2.size == 0 and np.array_equal(pl2, _np_hist(path_cm, 50, (lo, hi))[0].astype(np.uint64))
    finally:
        ctx.set_histograms()


def test_histogram_spec_is_validated():
    ctx = gpu_util.context()
    with pytest.raises(engine.Mc3dError):
        ctx.set_histograms(10, (3., 3.), 0, (0., 1.))
    with pytest.raises(engine.Mc3dError):
        ctx.set_histograms(0, (0., 1.), 10, (0., 1.), path_scale=0.)
    with pytest.raises(engine.Mc3dError):
        ctx.set_histograms(-1, (0., 1.), 0, (0., 1.))
    ctx.set_histograms()


def test_driver_histograms_match_post_processing(run_dir, optics_root):
    from monte_carloMPI import monte_carlo3D
    mc = monte_carlo3D.MonteCarlo(optics_dir=optics_root['spectral'], output_dir=str(run_dir / 'o'), devices=[0], seed=5)
    n = 200000
    h = mc.histograms(n, 1.3, 0.085, 100., theta_0=15., Lambertian_bottom=True, Lambertian_reflectance=0.5)
    mc.run(n, 1.3, 0.085, 100., theta_0=15., Lambertian_bottom=True, Lambertian_reflectance=0.5, write_output=False)
    rec = mc.last_records
    # the reference: np.histogram(data['path_length[m]'] * 100, bins=1000), np.histogram(data['n_scat'], bins=200)
    want_pl, edges_pl = np.histogram(rec['path_length'].astype(np.float64) * 100, bins=1000)
    want_ns, edges_ns = np.histogram(rec['n_scat'].astype(np.int64), bins=200)
    assert np.array_equal(h['n_scat'][0], want_ns.astype(np.uint64)) and np.array_equal(h['n_scat'][1], edges_ns)
    assert np.array_equal(h['path_length_cm'][0], want_pl.astype(np.uint64))
    assert np.array_equal(h['path_length_cm'][1], edges_pl)
    mc.close()
