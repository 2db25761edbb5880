"""Production mode (fp32 walk, Philox4x32-7 keyed on (seed, photon id)) through the C ABI.

Parity is shown three ways:
  1. against the oracle's production-mode restatement fed the SAME Philox draws (fp64, reference arithmetic): per-photon
     condition / n_scat / wavelength identical for all but a few fp32-vs-fp64 branch flips, float columns to fp32 accuracy;
  2. against statistics of the UNMODIFIED reference at 10^6 photons (tests/golden/stats_c2_reference.npz): outcome
     fractions and the 137 BRF zenith bins within 3 sigma binomial (north_star's production bar);
  3. size-independent properties at full size: counts add up, tallies equal np.histogram of the records, results do
     not depend on how the photon range is split or how the kernel is launched.
"""
import ast
import os

import numpy as np
import pytest

import golden_util as gu
import gpu_util
from monte_carlompi_b200 import engine

pytestmark = pytest.mark.gpu

SIGMA13 = 0.085 / 2.355


def _run(P, rows, seed, begin, n, **kw):
    return gpu_util.context().run(P, rows, seed, begin, n, **kw)


# ---------------------------------------------------------------------------------------------- 1. same-stream parity
CASES = {
    # name: (theta0, tau_tot, R, lambert_bottom, table, wvl0, sigma, k_first)
    'c2_semi_infinite': (15., 1e6, 0.5, True, ('spectral', 100, 104, 156, 0.0), 1.3, SIGMA13, 104),
    'slab_tau3_lambert': (15., 3.0, 0.5, True, ('spectral', 100, 104, 156, 0.0), 1.3, SIGMA13, 104),
    'thin_normal_incidence': (0., 0.5, 0.5, True, ('spectral', 100, 104, 156, 0.0), 1.3, SIGMA13, 104),
    'black_bottom_60deg': (60., 3.0, 1.0, False, ('spectral', 250, 74, 126, 0.0), 1.0, SIGMA13, 74),
    'impurity': (30., 3.0, 0.5, True, ('spectral', 100, 104, 156, 1e-5), 1.3, SIGMA13, 104),
    'bright_bottom_R1': (15., 1.0, 1.0, True, ('spectral', 100, 104, 156, 0.0), 1.3, SIGMA13, 104),
    'wide_band_clamped_table': (15., 4.0, 0.3, True, ('spectral', 100, 20, 45, 0.0), 0.33, 0.26 / 2.355, 20),
    # config C3's absorption-terminated corner: 2.1 um, 1000 um grains (about two events per photon)
    'nir_large_grains_short_walks': (30., 1e6, 0.5, True, ('spectral', 1000, 184, 236, 0.0), 2.1, SIGMA13, 184),
    # thin slab over a perfectly reflecting Lambertian bottom: most photons reach it on their FIRST step
    'thin_slab_mirror_bottom': (15., 0.3, 1.0, True, ('spectral', 100, 64, 116, 0.0), 0.9, SIGMA13, 64),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_same_philox_stream_as_oracle(name):
    from oracle import oracle
    th, tau, R, lb, tab, wvl0, sig, k0 = CASES[name]
    rows = gpu_util.fixture_table(*tab)
    Pe, Po = gpu_util.both_params(th, tau, R, wvl0, sig, k0, lb)
    n, seed, begin = 200000, 20190603, 12345
    rec, tally, st = _run(Pe, rows, seed, begin, n)
    # both kernel paths (the one-kernel short-walk path and the persistent three-kernel path) walk every case:
    # bit-identical records, tallies and event counts, whichever the library would pick by itself
    ctx = gpu_util.context()
    try:
        for path in ('fused', 'persistent'):
            ctx.set_walk_path(path)
            rec2, tally2, st2 = _run(Pe, rows, seed, begin, n)
            assert st2['walk_path'] == {'fused': engine.PATH_FUSED, 'persistent': engine.PATH_PERSISTENT}[path]
            for col in rec:
                assert np.array_equal(rec[col], rec2[col]), (path, col)
            assert np.array_equal(tally, tally2) and st2['n_events'] == st['n_events'], path
    finally:
        ctx.set_walk_path('auto')
    o = oracle.philox(Po, rows, seed, begin, n, n_threads=os.cpu_count())
    same = (rec['condition'] == o['condition']) & (rec['n_scat'] == o['n_scat']) & (rec['wvl_row'] == o['wvl_row'])
    # fp32 vs fp64 flips a branch for a few photons per 10^5 (rint at a wavelength bin edge, z within an ulp of 0)
    assert same.mean() >= 0.9995, same.mean()
    assert (rec['wvl_row'] == o['wvl_row']).mean() >= 0.9999
    for col, tol in (('theta_n', 2e-4), ('path_length', 2e-3)):
        a, b = rec[col][same].astype(np.float64), o[col][same]
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
        assert np.median(rel) < 2e-6 and np.percentile(rel, 99.9) < tol, (col, np.median(rel), rel.max())
    dphi = np.abs(rec['phi_n'][same].astype(np.float64) - o['phi_n'][same])
    dphi = np.minimum(dphi, 2 * np.pi - dphi)
    assert np.median(dphi) < 5e-6 and np.percentile(dphi, 99) < 1e-3
    # outcome counts can differ only by the flipped photons
    cg, co = np.bincount(rec['condition'], minlength=6), np.bincount(o['condition'], minlength=6)
    assert np.abs(cg - co).sum() <= 2 * (~same).sum()
    assert abs(int(st['n_events']) - o['n_events']) <= np.abs(rec['n_scat'].astype(np.int64) - o['n_scat'])[~same].sum()


def test_isotropic_and_backward_scattering_rows():
    # g == 0 takes the reference's `1 - 2 r` branch (monte_carlo3D.py:794-795); g < 0 as in its debug constants (1878)
    from oracle import oracle
    for g in (0.0, -0.89):
        rows = gpu_util.const_table(0.9, g)
        Pe, Po = gpu_util.both_params(45., 5.0, 1.0, 0.5, 0.0, 50, False)
        rec, _, _ = _run(Pe, rows, 7, 0, 100000)
        o = oracle.philox(Po, rows, 7, 0, 100000, n_threads=os.cpu_count())
        same = (rec['condition'] == o['condition']) & (rec['n_scat'] == o['n_scat'])
        assert same.mean() >= 0.9995, (g, same.mean())


def test_lambertian_surface_mode():
    # monte_carlo3D.py:1228-1250, 1385-1387: the snow replaced by a Lambertian reflector of reflectance R
    from oracle import oracle
    from monte_carlompi_b200 import post
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    th = np.pi * 40. / 180.
    n, R = 400000, 0.7
    pe = engine.make_params(th, 5.0, 300., R, 1.3, SIGMA13, 104, lambert_bottom=False, lambert_surface=True, n_theta_bins=137)
    po = oracle.make_params(th, 5.0, 300., R, 1.3, SIGMA13, 104, lambert_bottom=False, lambert_surface=True, n_theta_bins=137)
    rec, tally, st = _run(pe, rows, 31, 0, n)
    o = oracle.philox(po, rows, 31, 0, n, n_threads=os.cpu_count())
    same = (rec['condition'] == o['condition']) & (rec['n_scat'] == o['n_scat']) & (rec['wvl_row'] == o['wvl_row'])
    assert same.mean() >= 0.9995
    refl = rec['condition'] == 1
    assert abs(refl.mean() - R) < 4 * np.sqrt(R * (1 - R) / n) and set(np.unique(rec['condition'])) == {1, 4}
    assert (rec['n_scat'][refl] >= 1).all() and (rec['n_scat'][~refl] == 0).all()
    assert np.abs(rec['path_length'][refl]).max() < 1e-6          # nothing travelled inside the "snow"
    a, b = rec['theta_n'][same & refl].astype(np.float64), o['theta_n'][same & refl]
    assert np.percentile(np.abs(a - b), 99.9) < 1e-5
    # a Lambertian reflector: BRF == R in every zenith bin (within noise), the normalisation check of brf()
    mid, brf = post.brf_from_tally(tally, rows)
    inner = slice(10, 127)
    assert abs(np.mean(brf[inner]) - R) < 0.01 and np.std(brf[inner]) < 0.05
    assert np.array_equal(tally, gpu_util.tally_from_records(rec, len(rows), 137))


# ---------------------------------------------------------------------------------------------- 2. vs the reference
def _stats_file(name='c2'):
    return os.path.join(gu.GOLDEN_DIR, 'stats_%s_reference.npz' % name)


# c2: BASELINE.json configs[1] (semi-infinite).  slab_lb: tau_tot = 3 over a Lambertian bottom R = 0.5 (all of
# conditions 1-4; bottom reflections re-enter the walk).  impurity: tau_tot = 3 with 1e-5 black carbon (species draw,
# condition 5).  Each is 10^6 photons of the unmodified reference (oracle/make_golden_stats.py).  vis: BASELINE.json
# configs[3], visible wavelength, ~2700 events per photon, walks beyond 10^5 events -- 3000 photons only: the
# reference appends to six position arrays per event (monte_carlo3D.py:1350-1362), quadratic in the walk length.
@pytest.mark.parametrize('name', ['c2', 'slab_lb', 'impurity', 'vis'])
def test_outcomes_and_brf_within_3_sigma_of_reference(optics_root, name):
    from monte_carlompi_b200 import ssp
    if not os.path.isfile(_stats_file(name)):
        pytest.skip('reference statistics fixture not generated')
    z = np.load(_stats_file(name))
    cfg = ast.literal_eval(str(z['config']))
    n_ref = cfg['n_photon']
    scale = cfg['half_width'] / 2.355
    k_lo, k_hi = ssp.wavelength_grid(cfg['wvl0'], scale)
    rows = ssp.build_table(optics_root[cfg['fixture']], 'mie_sot_ChC90_dns_1317.nc', cfg['rds_snw'], k_lo, k_hi,
                           cfg.get('imp_cnc', 0.0))
    P = engine.make_params(np.pi * cfg['theta_0'] / 180., cfg['tau_tot'], 300., cfg['Lambertian_reflectance'], cfg['wvl0'],
                           scale, k_lo, lambert_bottom=cfg['Lambertian_bottom'], n_theta_bins=cfg['n_theta_bins'])
    n_gpu = 8000000 if name != 'vis' else 1000000
    _, tally, st = _run(P, rows, 20190603, 0, n_gpu, records=False)
    tally = tally.astype(np.int64)
    assert tally[:, 0].sum() == n_gpu

    def check(k_gpu, k_ref, what, z_max=3.0):
        p = (k_gpu + k_ref) / float(n_gpu + n_ref)
        sig = np.sqrt(p * (1 - p) * (1.0 / n_gpu + 1.0 / n_ref))
        zs = (k_gpu / float(n_gpu) - k_ref / float(n_ref)) / max(sig, 1e-300)
        assert abs(zs) <= z_max, (what, k_gpu / float(n_gpu), k_ref / float(n_ref), zs)
        return zs

    # outcome fractions (reflected, diffuse / direct transmitted, absorbed)
    for cond in (1, 2, 3, 4, 5):
        check(tally[:, cond].sum(), z['counts'][:, cond].sum(), 'condition %d' % cond)
    if name == 'slab_lb':
        assert min(z['counts'][:, c].sum() for c in (1, 2, 3, 4)) > 10000       # the fixture exercises every outcome
    if name == 'impurity':
        assert z['counts'][:, 5].sum() > 10000 and tally[:, 5].sum() > 10000
    # wavelength distribution of the drawn photons, per 10 nm bin (Box-Muller on Philox vs numpy's legacy normal)
    kr = int(z['k_first'])
    zs_w = [check(tally[kr - k_lo + j, 0], z['counts'][j, 0], 'wavelength row %d' % j, z_max=4.5)
            for j in range(z['counts'].shape[0])]
    assert np.sum(np.abs(zs_w) > 3.0) <= 1
    # BRF zenith bins of reflected photons (post_processing.py:73-76): 137 bins, each within 3 sigma; with 137
    # simultaneous tests ~0.4 excursions beyond 3 sigma are expected, so allow two, none beyond 4.5
    brf_gpu = tally[:, engine.N_COND:].sum(axis=0)
    brf_ref = z['brf'].sum(axis=0)
    zs = np.array([check(brf_gpu[b], brf_ref[b], 'BRF bin %d' % b, z_max=4.5) for b in range(len(brf_ref))])
    assert np.sum(np.abs(zs) > 3.0) <= 2, zs[np.abs(zs) > 3.0]
    assert abs(zs.mean()) < 0.3 and 0.8 < zs.std() < 1.25       # no systematic angular bias
    # mean number of scatterings per photon
    mean_ref = float(z['n_scat_sum']) / n_ref
    mean_gpu = (st['n_events'] - n_gpu) / float(n_gpu)
    hist = z['n_scat_hist'].astype(np.float64)
    var_ref = (hist * np.arange(len(hist)) ** 2).sum() / n_ref - mean_ref ** 2
    if 'n_scat_sq_sum' in z.files:                               # (the histogram clips at 4095 scatterings)
        var_ref = float(z['n_scat_sq_sum']) / n_ref - mean_ref ** 2
    assert abs(mean_gpu - mean_ref) < 3.5 * np.sqrt(var_ref * (1.0 / n_ref + 1.0 / n_gpu)), (mean_gpu, mean_ref)
    if name == 'vis':
        assert mean_ref > 1500 and z['n_scat_log2_hist'][14:].sum() > 10      # the fixture does hold long walks


@pytest.mark.parametrize('name', ['c2', 'slab_lb', 'impurity', 'vis'])
def test_path_length_statistics_match_reference(optics_root, name):
    # mean photon path length inside the slab (all photons, and reflected ones) against the reference at 10^6 photons
    from monte_carlompi_b200 import ssp
    if not os.path.isfile(_stats_file(name)):
        pytest.skip('reference statistics fixture not generated')
    z = np.load(_stats_file(name))
    cfg = ast.literal_eval(str(z['config']))
    n_ref = cfg['n_photon']
    scale = cfg['half_width'] / 2.355
    k_lo, k_hi = ssp.wavelength_grid(cfg['wvl0'], scale)
    rows = ssp.build_table(optics_root[cfg['fixture']], 'mie_sot_ChC90_dns_1317.nc', cfg['rds_snw'], k_lo, k_hi,
                           cfg.get('imp_cnc', 0.0))
    P = engine.make_params(np.pi * cfg['theta_0'] / 180., cfg['tau_tot'], 300., cfg['Lambertian_reflectance'], cfg['wvl0'],
                           scale, k_lo, lambert_bottom=cfg['Lambertian_bottom'], n_theta_bins=cfg['n_theta_bins'])
    n = 4000000 if name != 'vis' else 400000
    rec, _, _ = _run(P, rows, 424242, 0, n)
    path = rec['path_length'].astype(np.float64)
    if name == 'vis':
        # the distribution of walk lengths itself, octave by octave (chi-square against the reference's 3000 photons)
        h_ref = z['n_scat_log2_hist'].astype(np.float64)
        h_gpu = np.bincount(np.floor(np.log2(np.maximum(rec['n_scat'], 1))).astype(np.int64), minlength=len(h_ref))[:len(h_ref)]
        expect = h_gpu * (n_ref / float(n))
        keep = expect >= 5
        chi2 = ((h_ref[keep] - expect[keep]) ** 2 / expect[keep]).sum()
        assert chi2 < keep.sum() + 4.0 * np.sqrt(2.0 * keep.sum()), (chi2, keep.sum())
    mean_ref = float(z['path_sum']) / n_ref
    var_ref = float(z['path_sq_sum']) / n_ref - mean_ref ** 2
    assert abs(path.mean() - mean_ref) < 3.5 * np.sqrt(var_ref / n_ref + path.var() / n), (path.mean(), mean_ref)
    refl = rec['condition'] == 1
    n_refl_ref = int(z['counts'][:, 1].sum())
    mean_refl_ref = float(z['path_sum_reflected']) / n_refl_ref
    assert abs(path[refl].mean() - mean_refl_ref) < 3.5 * np.sqrt(path[refl].var() * (1.0 / n_refl_ref + 1.0 / refl.sum()))
    # metres: tau / (ext rho) with ext ~ 16.4 m2/kg, rho 300 kg/m3 -> sub-centimetre paths at 1.3 um
    assert name == 'vis' or 1e-3 < path.mean() < 5e-2


def test_known_answers_van_de_hulst():
    # monte_carlo3D.py:1849-1852: tau 2, omega 0.9, g 0.75, mu0 1, black bottom: albedo 0.09739, transmittance 0.66096
    n = 8000000
    P, _ = gpu_util.both_params(0., 2.0, 1.0, 0.5, SIGMA13, 50, False)
    _, tally, _ = _run(P, gpu_util.const_table(0.9, 0.75), 1, 0, n, records=False)
    frac = tally[0, :6].astype(np.float64) / n
    sig = lambda p: 3.0 * np.sqrt(p * (1 - p) / n) + 2e-5        # + the published values' last-digit rounding
    assert abs(frac[1] - 0.09739) < sig(0.09739), frac
    assert abs(frac[2] + frac[3] - 0.66096) < sig(0.66096), frac
    assert abs(frac[3] - np.exp(-2.0)) < sig(np.exp(-2.0)), frac


def test_direct_transmittance_and_lambertian_cosine_law():
    n = 4000000
    # direct beam through a purely absorbing... no: any slab: P(no extinction before the bottom) = exp(-tau / mu0)
    P, _ = gpu_util.both_params(60., 1.0, 1.0, 0.5, 0.0, 50, False)
    _, tally, _ = _run(P, gpu_util.const_table(0.5, 0.75), 2, 0, n, records=False)
    p = np.exp(-1.0 / np.cos(np.pi / 3))
    assert abs(tally[0, 3] / float(n) - p) < 3.5 * np.sqrt(p * (1 - p) / n)
    # non-scattering slab over a white Lambertian bottom: every reflected photon left the bottom with the cosine law
    # (monte_carlo3D.py:1238-1250), so E[cos(theta_n)] = 2/3 weighted by the escape probability exp(-tau / cos)
    P, _ = gpu_util.both_params(0., 0.05, 1.0, 0.5, 0.0, 50, True)
    rec, _, _ = _run(P, gpu_util.const_table(0.0, 0.0), 3, 0, 1000000)
    refl = rec['condition'] == 1
    c = np.cos(rec['theta_n'][refl].astype(np.float64))
    w = np.exp(-0.05 / np.linspace(1e-3, 1, 200001))
    mu = np.linspace(1e-3, 1, 200001)
    expect = (2 * mu * w * mu).sum() / (2 * mu * w).sum()
    assert abs(c.mean() - expect) < 4 * c.std() / np.sqrt(refl.sum())
    assert (rec['n_scat'][refl] == 1).all()                      # the bottom reflection counts as one scattering


# ---------------------------------------------------------------------------------------------- 3. properties
def test_tallies_equal_histograms_of_the_records_full_size():
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    P, _ = gpu_util.both_params(15., 1e6, 0.5, 1.3, SIGMA13, 104, True)
    n = 1000000                                                   # BASELINE.json configs[1]
    rec, tally, st = _run(P, rows, 5, 0, n)
    assert np.array_equal(tally, gpu_util.tally_from_records(rec, len(rows), 137))
    assert st['n_events'] == int(rec['n_scat'].astype(np.int64).sum()) + n
    assert tally[:, 0].sum() == n and tally[:, 1:6].sum() == n
    assert set(np.unique(rec['condition'])) <= {1, 4}             # semi-infinite, no impurity
    refl = rec['condition'] == 1
    assert (rec['theta_n'][refl] < np.pi / 2).all() and (rec['theta_n'] >= 0).all() and (rec['theta_n'] <= np.pi).all()
    assert (rec['phi_n'] >= 0).all() and (rec['phi_n'] <= np.float32(2 * np.pi)).all()
    assert (rec['phi_n'][rec['n_scat'] == 0] == 0).all() and (rec['path_length'] > 0).all()


def test_full_size_properties_1e8_photons():
    # BASELINE.json configs[3]-[4] sizes: tallies only; counts add up, chunking (2 chunks of 2^26 ids) and a range
    # that crosses a multiple of 2^32 in the photon id give the same totals as the pieces
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    P, _ = gpu_util.both_params(15., 1e6, 0.5, 1.3, SIGMA13, 104, True)
    n = 100000000
    begin = 2**32 - 30000000                      # crosses 2^32 inside the range
    _, tally, st = _run(P, rows, 9, begin, n, records=False)
    assert tally[:, 0].sum() == n and tally[:, 1:6].sum() == n and tally[:, 8:].sum() == tally[:, 1].sum()
    assert st['n_photon'] == n and 60 * n < st['n_events'] < 70 * n
    pieces = [(begin, 30000000), (2**32, 70000000)]
    tsum = sum(_run(P, rows, 9, b, c, records=False)[1] for b, c in pieces)
    assert np.array_equal(tsum, tally)
    # an independent 10^7-photon realisation (other seed) agrees on the reflected fraction within binomial noise
    frac = tally[:, 1].sum() / float(n)
    _, t7, _ = _run(P, rows, 10, 0, 10000000, records=False)
    f7 = t7[:, 1].sum() / 1e7
    assert abs(frac - f7) < 4 * np.sqrt(frac * (1 - frac) * (1e-7 + 1e-8))


def test_full_hemisphere_theta_phi_tally_is_histogram2d_of_the_records():
    # the reference stores phi_n but never bins it; with n_phi_bins > 1 the device splits every zenith bin in azimuth
    from monte_carlompi_b200 import post
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    n_t, n_p, n = 30, 24, 600000
    P = engine.make_params(np.pi * 50. / 180., 1e6, 300., 0.5, 1.3, SIGMA13, 104, lambert_bottom=True, n_theta_bins=n_t, n_phi_bins=n_p)
    rec, tally, _ = _run(P, rows, 77, 0, n)
    assert tally.shape == (len(rows), engine.N_COND + n_t * n_p) and P.tally_width == engine.N_COND + n_t * n_p
    refl = rec['condition'] == 1
    want = np.zeros((len(rows), n_t * n_p), np.int64)
    for r in range(len(rows)):
        m = refl & (rec['wvl_row'] == r)
        h2, _, _ = np.histogram2d(rec['theta_n'][m].astype(np.float64), rec['phi_n'][m].astype(np.float64), bins=(n_t, n_p),
                                  range=((0., np.pi / 2), (0., 2 * np.pi)))
        want[r] = h2.astype(np.int64).ravel()
    assert np.array_equal(tally[:, engine.N_COND:].astype(np.int64), want)
    # summing over azimuth gives the zenith-only tally of the same run
    P1 = engine.make_params(np.pi * 50. / 180., 1e6, 300., 0.5, 1.3, SIGMA13, 104, lambert_bottom=True, n_theta_bins=n_t)
    _, t1, _ = _run(P1, rows, 77, 0, n, records=False)
    dropped = int(refl.sum()) - int(want.sum())                 # phi_n == float32(2 pi) lies outside [0, 2 pi]
    assert 0 <= dropped <= 3
    assert np.abs(tally[:, engine.N_COND:].reshape(len(rows), n_t, n_p).sum(axis=2).astype(np.int64)
                  - t1[:, engine.N_COND:].astype(np.int64)).sum() == dropped
    # oblique incidence (50 deg) + forward-peaked phase function: the BRF is brighter in the forward direction (phi ~ 0)
    tm, pm, brf = post.brf2d_from_tally(tally, rows, n_t, n_p)
    fwd = brf[20:28][:, [0, n_p - 1]].mean()
    back = brf[20:28][:, [n_p // 2 - 1, n_p // 2]].mean()
    assert fwd > 1.15 * back


def test_results_do_not_depend_on_range_split_or_launch_shape():
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    P, _ = gpu_util.both_params(15., 6.0, 0.5, 1.3, SIGMA13, 104, True)
    ctx = gpu_util.context()
    n, seed = 300000, 11
    ref, tref, st = ctx.run(P, rows, seed, 0, n)
    try:
        # (a) the id range split like 8 ranks (np.array_split), tallies summed: bit-identical to one range
        from monte_carlompi_b200.parallelize import partition
        parts = [ctx.run(P, rows, seed, b, c) for b, c in partition(n, 8)]
        for col in ref:
            assert np.array_equal(np.concatenate([p[0][col] for p in parts]), ref[col]), col
        assert np.array_equal(sum(p[1] for p in parts), tref)
        # (b) other block sizes / occupancies / refill thresholds
        for bps, bt, thr in ((8, 128, 1), (2, 512, 16), (5, 256, 32), (4, 256, 7)):
            ctx.set_launch(bps, bt, thr)
            rec, t, _ = ctx.run(P, rows, seed, 0, n)
            for col in ref:
                assert np.array_equal(rec[col], ref[col]), (col, bps, bt, thr)
            assert np.array_equal(t, tref)
        # (c) the fused one-thread-per-photon kernel (chosen automatically for short walks) and the persistent
        #     three-kernel path follow the same per-photon stream: bit-identical records
        for path in ('fused', 'persistent'):
            os.environ['MC3D_WALK_PATH'] = path
            try:
                with engine.Context([0]) as c2:
                    rec, t, st2 = c2.run(P, rows, seed, 0, n)
            finally:
                del os.environ['MC3D_WALK_PATH']
            for col in ref:
                assert np.array_equal(rec[col], ref[col]), (col, path)
            assert np.array_equal(t, tref) and st2['n_events'] == st['n_events']
        # (d) drain-phase consolidation (photons handed between the warps of a block through shared memory; switched
        #     on automatically when other calls are in flight): any hand-over threshold, several grid shapes
        for give, bps in (('0', 255), ('8', 255), ('16', 1), ('24', 255), ('31', 2)):
            os.environ['MC3D_DRAIN_GIVE'] = give
            try:
                with engine.Context([0]) as c2:
                    c2.set_launch(bps, 256, 4)
                    rec, t, st2 = c2.run(P, rows, seed, 0, n)
            finally:
                del os.environ['MC3D_DRAIN_GIVE']
            for col in ref:
                assert np.array_equal(rec[col], ref[col]), (col, give)
            assert np.array_equal(t, tref) and st2['n_events'] == st['n_events']
        # (f) how the call ends: the walk kernel draining by itself (latency-oriented groups), or the tail kernel (dense
        #     warps, then helper lanes preparing a lone photon's next eight events): bit-identical
        for mode, bps in ((0, 255), (1, 255), (1, 1), (0, 4), (1, 4)):
            ctx.set_tail_kernel(mode)
            ctx.set_launch(bps, 256, 4)
            rec, t, st2 = ctx.run(P, rows, seed, 0, n)
            for col in ref:
                assert np.array_equal(rec[col], ref[col]), (col, mode, bps)
            assert np.array_equal(t, tref) and st2['n_events'] == st['n_events']
    finally:
        ctx.set_tail_kernel(-1)
        ctx.set_launch(255, 256, 4)      # back to the automatic grid
    # (e) a different seed gives a different realisation
    other, _, _ = ctx.run(P, rows, seed + 1, 0, n)
    assert (other['n_scat'] != ref['n_scat']).mean() > 0.5


@pytest.mark.parametrize('what', ['visible_long_walks', 'impurity', 'lambert_bottom_thin', 'tiny'])
def test_tail_kernel_is_bit_identical(what):
    """A call that runs alone ends in tail_kernel.cu (helper lanes prepare the next eight events of a warp's last
    photons).  Same per-photon stream, same arithmetic: identical records, tallies and event counts, also where walks
    are thousands of events long, where the impurity species is drawn between events (no helpers) and where a
    Lambertian bottom reflects photons back into the walk."""
    if what == 'visible_long_walks':
        rows, n = gpu_util.const_table(0.999989859099, 0.89, ext=6.6), 3000
        P, _ = gpu_util.both_params(15., 1e6, 0.5, 0.5, SIGMA13, 24, True)
    elif what == 'impurity':
        rows, n = gpu_util.fixture_table('spectral', 100, 104, 156, 1e-5), 100000
        P, _ = gpu_util.both_params(30., 3.0, 0.5, 1.3, SIGMA13, 104, True)
    elif what == 'lambert_bottom_thin':
        rows, n = gpu_util.fixture_table('spectral', 100, 64, 116), 100000
        P, _ = gpu_util.both_params(15., 0.3, 1.0, 0.9, SIGMA13, 64, True)
    else:
        rows, n = gpu_util.fixture_table('spectral', 100, 104, 156), 37
        P, _ = gpu_util.both_params(15., 1e6, 0.5, 1.3, SIGMA13, 104, True)
    ctx = gpu_util.context()
    out = {}
    try:
        ctx.set_walk_path('persistent')
        for mode in (0, 1):
            ctx.set_tail_kernel(mode)
            out[mode] = ctx.run(P, rows, 5, 1 << 34, n)
    finally:
        ctx.set_tail_kernel(-1)
        ctx.set_walk_path('auto')
    for col in out[0][0]:
        assert np.array_equal(out[0][0][col], out[1][0][col]), col
    assert np.array_equal(out[0][1], out[1][1]) and out[0][2]['n_events'] == out[1][2]['n_events']
    if what == 'visible_long_walks':
        assert out[1][0]['n_scat'].max() > 20000


@pytest.mark.parametrize('n', [0, 1, 31, 32, 33, 1000])
def test_ragged_and_tiny_ranges(n):
    from oracle import oracle
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    Pe, Po = gpu_util.both_params(15., 3.0, 0.5, 1.3, SIGMA13, 104, True)
    rec, tally, st = _run(Pe, rows, 3, 2**33 + 5, n)             # ids beyond 32 bits exercise the high counter word
    o = oracle.philox(Po, rows, 3, 2**33 + 5, n)
    assert len(rec['condition']) == n and tally[:, 0].sum() == n and st['n_photon'] == n
    if n:
        assert (rec['condition'] == o['condition']).mean() >= 0.99 and (rec['n_scat'] == o['n_scat']).mean() >= 0.99


def test_long_visible_walks():
    # weakly absorbing ice, mean ~2700 events per photon, walks beyond 10^5 events: direction
    # renormalisation and the split path accumulator keep fp32 honest over the whole walk
    from oracle import oracle
    rows = gpu_util.const_table(0.999989859099, 0.89, ext=6.6)
    Pe, Po = gpu_util.both_params(15., 1e6, 0.5, 0.5, 0.0, 50, True)
    n = 20000
    rec, tally, st = _run(Pe, rows, 21, 0, n)
    o = oracle.philox(Po, rows, 21, 0, n, n_threads=os.cpu_count())
    assert rec['n_scat'].max() > 100000
    same = (rec['condition'] == o['condition']) & (rec['n_scat'] == o['n_scat'])
    # chaotic divergence: after thousands of events an fp32 path decorrelates from the fp64 one; only short walks match
    short = o['n_scat'] < 50
    assert same[short].mean() > 0.995
    # ... but the statistics agree: reflected fraction and mean walk length
    pr, po_ = (rec['condition'] == 1).mean(), (o['condition'] == 1).mean()
    assert abs(pr - po_) < 4 * np.sqrt(po_ * (1 - po_) * 2 / n)
    lg, lo = np.log1p(rec['n_scat'].astype(np.float64)), np.log1p(o['n_scat'].astype(np.float64))
    assert abs(lg.mean() - lo.mean()) < 4 * lo.std() * np.sqrt(2.0 / n)
    # path length of long walks: metres = sum(dtau) / (ext rho); compare with n_scat / ext rho to 1 %
    big = rec['n_scat'] > 10000
    ratio = rec['path_length'][big].astype(np.float64) * (6.6 * 300.) / (rec['n_scat'][big] + 1.0)
    assert abs(ratio.mean() - 1.0) < 0.02


def test_packed_records_equal_the_columns():
    # the 16-byte packed record (the default way records come back) against the six-column form, incl. every condition
    rows = gpu_util.fixture_table('spectral', 100, 104, 156, 1e-5)
    P, _ = gpu_util.both_params(30., 3.0, 0.5, 1.3, SIGMA13, 104, True)
    ctx = gpu_util.context()
    n = 300000
    a, ta, _ = ctx.run(P, rows, 17, 5, n)
    b, tb, _ = ctx.run(P, rows, 17, 5, n, records='columns')
    assert set(np.unique(a['condition'])) == {1, 2, 3, 4, 5}
    for col in a:
        assert a[col].dtype == b[col].dtype and np.array_equal(a[col], b[col]), col
    assert np.array_equal(ta, tb)
    big = np.zeros(engine.PACKED_MAX_ROWS + 1, engine.ROW_DTYPE)      # more rows than the packed form can index
    big[:] = rows[0]
    Pb, _ = gpu_util.both_params(30., 3.0, 0.5, 1.3, 0.0, 130 - 256, True)
    with pytest.raises(engine.Mc3dError):
        ctx.run_async(0, Pb, big, 1, 0, 10, engine.RecordBuffers(10), None)
    rec, _, _ = ctx.run(Pb, big, 1, 0, 1000)                            # ctx.run falls back to the columns by itself
    assert (rec['wvl_row'] == 256).all()


def test_argument_validation():
    rows = gpu_util.const_table(0.9, 0.75)
    ctx = gpu_util.context()
    for kw in (dict(tau_tot=-1.0), dict(rho_snw=0.0), dict(theta0_rad=2.0)):
        args = dict(theta0_rad=0.1, tau_tot=2.0, rho_snw=300., r_lambert=1.0, wvl0_um=0.5, sigma_um=0.0, k_first=50)
        args.update(kw)
        with pytest.raises(engine.Mc3dError):
            ctx.run(engine.make_params(**args), rows, 1, 0, 10)
    bad = rows.copy()
    bad['g'] = 1.0
    with pytest.raises(engine.Mc3dError):
        ctx.run(engine.make_params(0.1, 2.0, 300., 1.0, 0.5, 0.0, 50), bad, 1, 0, 10)


def test_async_slots_overlap_and_match_sync():
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    P, _ = gpu_util.both_params(15., 1e6, 0.5, 1.3, SIGMA13, 104, True)
    ctx = gpu_util.context()
    n = 200000
    bufs = [engine.RecordBuffers(n), engine.RecordBuffers(n)]
    tallies = [np.zeros((len(rows), engine.N_COND + 137), np.uint64) for _ in range(2)]
    ctx.run_async(0, P, rows, 9, 0, n, bufs[0], tallies[0])
    ctx.run_async(1, P, rows, 9, n, n, bufs[1], tallies[1])
    s0, s1 = ctx.wait(0), ctx.wait(1)
    ref, tref, _ = ctx.run(P, rows, 9, 0, 2 * n)
    for col in ref:
        assert np.array_equal(np.concatenate([bufs[0].view(n)[col], bufs[1].view(n)[col]]), ref[col]), col
    assert np.array_equal(tallies[0] + tallies[1], tref) and s0['n_events'] + s1['n_events'] > 2 * n
    [b.free() for b in bufs]


def test_every_slot_in_flight_with_ragged_sizes():
    # all MC3D_N_SLOTS slots busy at once (drain-phase consolidation switches on by itself), different photon counts
    # per slot, buffers larger than the call: the concatenation equals one synchronous run over the whole id range
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    P, _ = gpu_util.both_params(15., 6.0, 0.5, 1.3, SIGMA13, 104, True)
    ctx = gpu_util.context()
    counts = [30000 + 977 * s for s in range(engine.N_SLOTS)]
    begins = np.concatenate([[0], np.cumsum(counts)])
    bufs = [engine.RecordBuffers(max(counts) + 123) for _ in counts]
    tallies = [np.zeros((len(rows), P.tally_width), np.uint64) for _ in counts]
    for s, c in enumerate(counts):
        ctx.run_async(s, P, rows, 21, int(begins[s]), c, bufs[s], tallies[s])
    with pytest.raises(engine.Mc3dError):
        ctx.run_async(3, P, rows, 21, 0, 10, bufs[3], tallies[3])           # busy slot
    with pytest.raises(engine.Mc3dError):
        ctx.run_async(engine.N_SLOTS, P, rows, 21, 0, 10, None, None)        # no such slot
    events = sum(ctx.wait(s)['n_events'] for s in range(len(counts)))
    ref, tref, st = ctx.run(P, rows, 21, 0, int(begins[-1]))
    for col in ref:
        assert np.array_equal(np.concatenate([bufs[s].view(c)[col] for s, c in enumerate(counts)]), ref[col]), col
    assert np.array_equal(sum(tallies), tref) and events == st['n_events']
    # input caching off (every call uploads its table again) and a table that changes between calls on one slot
    ctx.set_input_caching(False)
    try:
        again, tagain, _ = ctx.run(P, rows, 21, 0, counts[0])
    finally:
        ctx.set_input_caching(True)
    assert np.array_equal(again['n_scat'], ref['n_scat'][:counts[0]])
    rows2 = rows.copy()
    rows2['ssa_ice'] = 0.5
    dark, _, _ = ctx.run(P, rows2, 21, 0, counts[0])
    assert dark['n_scat'].mean() < 0.2 * again['n_scat'].mean()                # the new table was uploaded, not the cached one
    back, _, _ = ctx.run(P, rows, 21, 0, counts[0])
    assert np.array_equal(back['n_scat'], again['n_scat'])
    [b.free() for b in bufs]
