"""Replay mode on the GPU (mc3d_replay, fp64) against records of the UNMODIFIED reference.

north_star bar: per-photon condition and n_scat exact for >= 99.99 % of photons (mismatches only from libm-vs-CUDA
ulp branch flips), angles and path_length within 1e-9 relative."""
import numpy as np
import pytest

import golden_util as gu
import gpu_util
from monte_carlompi_b200 import engine

pytestmark = pytest.mark.gpu


def _replay(c):
    cfg = c['cfg']
    P = engine.make_params(np.pi * cfg['theta_0'] / 180., cfg['tau_tot'], cfg['rho_snw'], cfg['Lambertian_reflectance'],
                           cfg['wvl0'], cfg['half_width'] / 2.355, 0, lambert_bottom=cfg['Lambertian_bottom'],
                           lambert_surface=cfg.get('Lambertian_surface', False))
    return gpu_util.context().replay(P, c['wvl'], c['ssa_ice'], c['ssa_imp'], c['g'], c['ext_cff_mss'], c['p_ext_imp'],
                                     c['init_draws'], c['offsets'], c['stream'])


@pytest.mark.parametrize('name', gu.CASES)
def test_replay_reproduces_reference(name):
    c = gu.load_case(name)
    stats = gu.compare_replay(_replay(c), c, min_exact=0.9999, rtol=1e-9)
    assert stats['n_mismatch'] <= max(1, int(1e-4 * stats['n']))


def test_replay_equals_oracle_bit_for_bit_on_discrete_columns():
    from oracle import oracle
    c = gu.load_case('slab_tau3_lb')
    cfg = c['cfg']
    Po = oracle.make_params(np.pi * cfg['theta_0'] / 180., cfg['tau_tot'], cfg['rho_snw'], cfg['Lambertian_reflectance'],
                            lambert_bottom=cfg['Lambertian_bottom'])
    o = oracle.replay(Po, c['wvl'], c['ssa_ice'], c['ssa_imp'], c['g'], c['ext_cff_mss'], c['p_ext_imp'], c['init_draws'],
                      c['offsets'], c['stream'])
    g = _replay(c)
    for col in ('condition', 'n_scat', 'consumed'):
        assert np.array_equal(g[col], o[col]), col


def test_replay_detects_a_truncated_stream():
    c = gu.load_case('slab_tau05_normal')
    off = c['offsets'].copy()
    p = int(np.argmax(np.diff(off) >= 10))           # a photon with at least two scatterings
    off[p + 1:] -= 5                                 # drop its last event from the recorded stream
    stream = np.delete(c['stream'], np.arange(c['offsets'][p + 1] - 5, c['offsets'][p + 1]))
    c2 = dict(c, offsets=off, stream=stream)
    out = _replay(c2)
    assert out['n_mismatch'] >= 1 and out['consumed'][p] != off[p + 1] - off[p]


def test_replay_argument_errors():
    c = gu.load_case('kat_vdh')
    off = c['offsets'].copy()
    off[5], off[6] = off[6], off[5] - 1          # not non-decreasing -> MC3D_ESTREAM from the library
    bad = dict(c, offsets=off)
    with pytest.raises(engine.Mc3dError):
        _replay(bad)
