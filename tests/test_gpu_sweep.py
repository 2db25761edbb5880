"""Single-launch sweeps (mc3d_run_sweep): many cases -- the loops of the reference's driver over wavelength, grain
radius and zenith angle, monte_carlo3D-run.py:60-96, 112-122 -- walked by the same launches over a concatenated SSP
table, the case index in the high bits of the photon id.

  1. per photon against the oracle (fp64 restatement fed the same Philox draws) for cases in the middle of a sweep;
  2. bit-identical to one mc3d_run per case with photon_begin = case << 40 (records, tallies, events), on both kernel
     paths, for any split of the concatenated photon range (how ranks share a sweep), across the 2^26-photon chunks.
"""
import os

import numpy as np
import pytest

import gpu_util
from monte_carlompi_b200 import engine

pytestmark = pytest.mark.gpu

SIGMA = 0.085 / 2.355


def _cases():
    """A small sweep with everything that differs between cases: table, zenith angle, slab depth, boundary options,
    photon count (including an empty case)."""
    t13 = gpu_util.fixture_table('spectral', 100, 104, 156)
    t17 = gpu_util.fixture_table('spectral', 250, 144, 196)
    t21 = gpu_util.fixture_table('spectral', 1000, 184, 236)
    t09 = gpu_util.fixture_table('spectral', 50, 64, 116)
    table = np.concatenate([t13, t17, t21, t09, t13])        # the last block: t13 again for a different slab depth
    rb = np.cumsum([0, len(t13), len(t17), len(t21), len(t09)])
    spec = [  # (theta0, tau_tot, R, lambert_bottom, lambert_surface, wvl0, row block, n_photon)
        (15., 1e6, 0.5, True, False, 1.3, 0, 30000),
        (60., 1e6, 0.5, True, False, 1.3, 0, 20011),          # shares the rows of case 0
        (0., 1e6, 1.0, True, False, 1.7, 1, 40000),
        (30., 1e6, 0.5, True, False, 2.1, 2, 50000),          # two events per photon
        (45., 1e6, 0.5, True, False, 0.9, 3, 0),              # empty
        (45., 1e6, 0.5, True, False, 0.9, 3, 3000),           # ~500 events per photon
        (15., 2.0, 0.7, True, False, 1.3, 4, 25000),          # finite slab over a Lambertian bottom: own copy of the rows
        (40., 1e6, 0.6, False, True, 1.7, 1, 10000),          # Lambertian surface
    ]
    cases, oracle_cases = [], []
    for th, tau, R, lb, ls, wvl0, blk, n in spec:
        k0 = int(round(wvl0 * 100)) - 26
        pe = engine.make_params(np.pi * th / 180., tau, 300., R, wvl0, SIGMA, k0, lambert_bottom=lb, lambert_surface=ls,
                                n_theta_bins=137)
        nrows = len(t13)
        cases.append((pe, int(rb[blk]), nrows, n))
        oracle_cases.append((th, tau, R, lb, ls, wvl0, k0))
    return table, cases, oracle_cases


def _individual(ctx, table, cases, seed):
    out = []
    for c, (pe, rb, nr, n) in enumerate(cases):
        rec, tally, st = ctx.run(pe, table[rb:rb + nr], seed, c << engine.SWEEP_ID_SHIFT, n)
        out.append((rec, tally, st))
    return out


@pytest.mark.parametrize('path', ['auto', 'persistent', 'fused'])
def test_sweep_equals_one_run_per_case(path):
    table, cases, _ = _cases()
    seed = 20190603
    ctx = gpu_util.context()
    try:
        ctx.set_walk_path(path)
        per_case, tally, events, st = ctx.run_sweep(cases, table, seed)
        ctx.set_walk_path('auto')
        ref = _individual(ctx, table, cases, seed)
    finally:
        ctx.set_walk_path('auto')
    expect = np.zeros_like(tally)
    for c, (pe, rb, nr, n) in enumerate(cases):
        rec, t, s = ref[c]
        for col in rec:
            assert np.array_equal(per_case[c][col], rec[col]), (c, col)
        expect[rb:rb + nr] += t
        assert int(events[c]) == int(s['n_events']), c
    assert np.array_equal(tally, expect)
    assert int(st['n_events']) == int(events.sum()) and int(st['n_photon']) == sum(c[3] for c in cases)


def test_sweep_cases_match_the_oracle_per_photon():
    from oracle import oracle
    table, cases, ocases = _cases()
    seed = 7
    per_case, tally, events, _ = gpu_util.context().run_sweep(cases, table, seed)
    for c in (1, 2, 6):                                       # cases away from the start of the id space
        th, tau, R, lb, ls, wvl0, k0 = ocases[c]
        pe, rb, nr, n = cases[c]
        po = oracle.make_params(np.pi * th / 180., tau, 300., R, wvl0, SIGMA, k0, lambert_bottom=lb, lambert_surface=ls,
                                n_theta_bins=137)
        o = oracle.philox(po, table[rb:rb + nr], seed, c << engine.SWEEP_ID_SHIFT, n, n_threads=os.cpu_count())
        rec = per_case[c]
        same = (rec['condition'] == o['condition']) & (rec['n_scat'] == o['n_scat']) & (rec['wvl_row'] == o['wvl_row'])
        assert same.mean() >= 0.9995, (c, same.mean())
        a, b = rec['theta_n'][same].astype(np.float64), o['theta_n'][same]
        assert np.percentile(np.abs(a - b), 99.9) < 2e-4
        a, b = rec['path_length'][same].astype(np.float64), o['path_length'][same]
        assert np.percentile(np.abs(a - b) / np.maximum(b, 1e-6), 99.9) < 2e-3
        assert abs(int(events[c]) - o['n_events']) <= np.abs(rec['n_scat'].astype(np.int64) - o['n_scat'])[~same].sum()


def test_any_split_of_the_photon_range_gives_the_same_sweep():
    """Ranks of a multi-rank context share a sweep by sub-ranges of its concatenated photons (np.array_split
    boundaries, parallelize.py:14-15): records concatenate, tallies and events add up -- bit for bit."""
    from monte_carlompi_b200.parallelize import partition
    table, cases, _ = _cases()
    seed = 99
    ctx = gpu_util.context()
    total = sum(c[3] for c in cases)
    full = np.empty(4 * total, np.uint32)
    t_full = np.zeros((len(table), cases[0][0].tally_width), np.uint64)
    e_full = np.zeros(len(cases), np.uint64)
    ctx.run_sweep_async(0, cases, table, seed, full, t_full, e_full)
    ctx.wait(0)
    for world in (2, 3, 8):
        parts, t_sum, e_sum = [], np.zeros_like(t_full), np.zeros_like(e_full)
        for begin, count in partition(total, world):
            rec = np.empty(4 * max(count, 1), np.uint32)
            t = np.zeros_like(t_full)
            e = np.zeros_like(e_full)
            ctx.run_sweep_async(1, cases, table, seed, rec, t, e, range_begin=begin, range_count=count)
            ctx.wait(1)
            parts.append(rec[:4 * count])
            t_sum += t
            e_sum += e
        assert np.array_equal(np.concatenate(parts), full), world
        assert np.array_equal(t_sum, t_full) and np.array_equal(e_sum, e_full), world


def test_sweep_across_launch_chunks():
    """A launch covers at most 2^26 photons: cases that straddle the chunk boundaries (tallies only, short walks)."""
    t21 = gpu_util.fixture_table('spectral', 1000, 184, 236)
    t25 = gpu_util.fixture_table('spectral', 1000, 224, 276)
    table = np.concatenate([t21, t25])
    mk = lambda th, wvl0: engine.make_params(np.pi * th / 180., 1e6, 300., .5, wvl0, SIGMA, int(round(wvl0 * 100)) - 26,
                                             lambert_bottom=True, n_theta_bins=137)
    n = (1 << 25) + 12345
    cases = [(mk(0., 2.1), 0, 53, n), (mk(60., 2.5), 53, 53, n), (mk(30., 2.1), 0, 53, n)]     # 1.5 chunks in total
    ctx = gpu_util.context()
    _, tally, events, st = ctx.run_sweep(cases, table, 5, records=False)
    expect = np.zeros_like(tally)
    for c, (pe, rb, nr, m) in enumerate(cases):
        _, t, s = ctx.run(pe, table[rb:rb + nr], 5, c << engine.SWEEP_ID_SHIFT, m, records=False)
        expect[rb:rb + nr] += t
        assert int(events[c]) == int(s['n_events'])
    assert np.array_equal(tally, expect)


@pytest.mark.parametrize('path,wvl0', [('fused', 2.1), ('persistent', 1.5), ('auto', 1.7)])
def test_sweep_with_a_table_too_large_for_a_shared_memory_tally(path, wvl0):
    """Five grain sizes x five zenith angles of one wavelength (one iteration of the reference driver's loops,
    monte_carlo3D-run.py:60-96): 265 rows x 145 tally words do not fit in shared memory, so the blocks keep the rows of
    the case they are working on there and re-target as they move through the cases (finalize_window)."""
    k0 = int(round(wvl0 * 100)) - 26
    tables = [gpu_util.fixture_table('spectral', r, k0, k0 + 52) for r in (50, 100, 250, 500, 1000)]
    table = np.concatenate(tables)
    cases = []
    for j in range(5):
        for th in (0., 15., 30., 45., 60.):
            pe = engine.make_params(np.pi * th / 180., 1e6, 300., .5, wvl0, SIGMA, k0, lambert_bottom=True, n_theta_bins=137)
            cases.append((pe, 53 * j, 53, 20000 + 777 * len(cases)))
    ctx = gpu_util.context()
    try:
        ctx.set_walk_path(path)
        per_case, tally, events, st = ctx.run_sweep(cases, table, 11)
    finally:
        ctx.set_walk_path('auto')
    expect = np.zeros_like(tally)
    for c, (pe, rb, nr, n) in enumerate(cases):
        rec, t, s = ctx.run(pe, table[rb:rb + nr], 11, c << engine.SWEEP_ID_SHIFT, n)
        for col in rec:
            assert np.array_equal(per_case[c][col], rec[col]), (c, col)
        expect[rb:rb + nr] += t
        assert int(events[c]) == int(s['n_events']), c
    assert np.array_equal(tally, expect)
    assert tally[:, 0].sum() == sum(c[3] for c in cases)


def test_sweep_rejects_bad_arguments():
    table, cases, _ = _cases()
    ctx = gpu_util.context()
    bad = list(cases)
    pe = cases[0][0]
    other = engine.make_params(pe.theta0_rad, 3.0, pe.rho_snw, pe.r_lambert, pe.wvl0_um, pe.sigma_um, pe.k_first,
                               n_theta_bins=137)
    bad[1] = (other, cases[0][1], cases[0][2], 10)            # shares rows with case 0 but another slab depth
    with pytest.raises(engine.Mc3dError, match='share rows'):
        ctx.run_sweep(bad, table, 1, records=False)
    with pytest.raises(engine.Mc3dError, match='outside the table'):
        ctx.run_sweep([(pe, len(table) - 5, 53, 10)], table, 1, records=False)
    with pytest.raises(engine.Mc3dError, match='n_theta_bins'):
        p90 = engine.make_params(pe.theta0_rad, 1e6, 300., .5, 1.3, SIGMA, 104, n_theta_bins=90)
        ctx.run_sweep([cases[0], (p90, cases[2][1], 53, 10)], table, 1, records=False)
    # the context is still usable
    per_case, _, _, _ = ctx.run_sweep(cases[:2], table, 1)
    assert len(per_case[1]['condition']) == cases[1][3]
