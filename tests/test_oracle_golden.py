"""The oracle is pinned here: oracle/mc3d_oracle.c (fp64 restatement) against records of the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/make_golden.py under the import shims) and against published known answers."""
import numpy as np
import pytest

import golden_util as gu
from oracle import oracle


def _params(cfg, **kw):
    return oracle.make_params(np.pi * cfg['theta_0'] / 180., cfg['tau_tot'], cfg['rho_snw'],
                              cfg['Lambertian_reflectance'], lambert_bottom=cfg['Lambertian_bottom'],
                              lambert_surface=cfg.get('Lambertian_surface', False), **kw)


@pytest.mark.parametrize('name', gu.CASES)
def test_replay_matches_reference_records(name):
    c = gu.load_case(name)
    out = oracle.replay(_params(c['cfg']), c['wvl'], c['ssa_ice'], c['ssa_imp'], c['g'], c['ext_cff_mss'],
                        c['p_ext_imp'], c['init_draws'], c['offsets'], c['stream'])
    stats = gu.compare_replay(out, c, min_exact=1.0, rtol=1e-9)
    assert stats['n_mismatch'] == 0 and stats['exact_fraction'] == 1.0


def test_golden_cases_cover_every_outcome():
    seen = set()
    for name in gu.CASES:
        seen |= set(np.unique(gu.load_case(name)['golden']['condition']).tolist())
    assert seen == {1, 2, 3, 4, 5}


def test_recorded_stream_accounting():
    # every photon consumes 5 uniforms per scatter plus Lambertian-bottom extras
    for name in ('c1_default', 'slab_tau3_lb', 'slab_tau05_normal'):
        c = gu.load_case(name)
        extras = np.diff(c['offsets']) - 5 * c['golden']['n_scat']
        assert (extras >= 0).all()
        if not c['cfg']['Lambertian_bottom'] or c['cfg']['tau_tot'] > 1e5:
            assert (extras == 0).all()
        else:
            assert extras.sum() > 0


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 with 7 rounds (the production stream) and with 10 rounds
    pi_ctr, pi_key = [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]
    kat = [(7, [0, 0, 0, 0], [0, 0], [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]),
           (7, [0xffffffff] * 4, [0xffffffff] * 2, [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]),
           (7, pi_ctr, pi_key, [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]),
           (10, [0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           (10, [0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           (10, pi_ctr, pi_key, [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for rounds, ctr, key, expect in kat:
        assert oracle.philox4x32(ctr, key, rounds).tolist() == expect


def test_henyey_greenstein_inverse_cdf():
    # reference monte_carlo3D.py:790-800: end points, the g == 0 branch and E[cos(theta)] = g (comment at 1006-1008)
    assert oracle.henyey_greenstein2(0.0, 0.25) == 0.5
    for g in (0.75, 0.89, -0.89, 0.3):
        assert abs(oracle.henyey_greenstein2(g, 0.0) + 1.0) < 1e-12
        assert abs(oracle.henyey_greenstein2(g, 1.0) - 1.0) < 1e-12
        r = (np.arange(200000) + 0.5) / 200000
        mean = np.mean([oracle.henyey_greenstein2(g, x) for x in r[::40]])
        assert abs(mean - g) < 2e-3


def test_histogram_bin_is_numpy_histogram():
    rng = np.random.RandomState(3)
    x = np.concatenate([rng.uniform(0, np.pi / 2, 5000), oracle.theta_edges(137), [0.0, np.pi / 2, 2.0, -1e-9]])
    for nb in (137, 90):
        ref = np.zeros(nb, np.int64)
        for v in x:
            b = oracle.histogram_bin(v, nb)
            if b >= 0:
                ref[b] += 1
        assert np.array_equal(ref, np.histogram(x, bins=nb, range=(0., np.pi / 2))[0])


def test_production_restatement_known_answers():
    # van de Hulst (1980) / Wang et al. (1995): tau 2, omega 0.9, g 0.75, normal incidence, black lower boundary:
    # albedo 0.09739, total transmittance 0.66096 (reference monte_carlo3D.py:1849-1852); direct beam exp(-2)
    rows = np.zeros(1, oracle.ROW_DTYPE)
    rows[0] = (0.5, 0.9, 0.3, 0.75, 16.4, 0.0)
    n = 400000
    P = oracle.make_params(0.0, 2.0, 300., 1.0, 0.5, 0.0, 50, lambert_bottom=False, n_theta_bins=137)
    o = oracle.philox(P, rows, seed=99, begin=0, n=n, n_threads=4)
    frac = np.bincount(o['condition'], minlength=6) / n
    sig = lambda p: 3.5 * np.sqrt(p * (1 - p) / n)
    assert abs(frac[1] - 0.09739) < sig(0.09739)
    assert abs(frac[2] + frac[3] - 0.66096) < sig(0.66096)
    assert abs(frac[3] - np.exp(-2.0)) < sig(np.exp(-2.0))
    t = o['tally'][0]
    assert t[0] == n and t[1:6].sum() == n and t[8:].sum() == t[1]
    assert o['n_events'] == int(o['n_scat'].sum()) + n
    # thread count and range splitting do not change per-photon results
    a = oracle.philox(P, rows, 99, 1000, 5000, n_threads=1)
    b = oracle.philox(P, rows, 99, 0, n, n_threads=3)
    for col in ('condition', 'n_scat', 'theta_n', 'path_length'):
        assert np.array_equal(a[col], b[col][1000:6000])
