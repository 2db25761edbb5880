"""Small workload touching every kernel and branch, meant to be run under compute-sanitizer (tools/sanitize.sh):
init / walk (impurity on and off; Lambertian bottom and surface; drain by itself and through the tail kernel with its
helper lanes) / fused (three stages over shared-memory queues) / finalize (tallies in shared and in global memory,
one case's rows at a time in sweeps, theta x phi bins, column histograms, packed records) / sweeps / replay."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from monte_carlompi_b200 import engine  # noqa: E402
import gpu_util  # noqa: E402
import golden_util as gu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = engine.Context([0])
sig = 0.085 / 2.355
rows = gpu_util.fixture_table('spectral', 100, 104, 156)
rows_imp = gpu_util.fixture_table('spectral', 100, 104, 156, imp_cnc=1e-5)
vis = gpu_util.const_table(0.999989859099, 0.89, ext=6.6)
total = 0
for name, table, kw, th, tau, wvl0, s, k0 in [
        ('nir semi-infinite', rows, dict(n_theta_bins=137), 15., 1e6, 1.3, sig, 104),
        ('nir theta x phi', rows, dict(n_theta_bins=90, n_phi_bins=36), 15., 6.0, 1.3, sig, 104),
        ('big tally (global atomics)', rows, dict(n_theta_bins=900, n_phi_bins=36), 15., 6.0, 1.3, sig, 104),
        ('impurity slab', rows_imp, dict(n_theta_bins=137), 30., 3.0, 1.3, sig, 104),
        ('thin slab, black bottom', rows, dict(n_theta_bins=10, lambert_bottom=False), 0., 0.5, 1.3, sig, 104),
        ('lambert surface', rows, dict(n_theta_bins=137, lambert_surface=True, lambert_bottom=False), 40., 5.0, 1.3, sig, 104),
        ('visible, long walks', vis, dict(n_theta_bins=137), 15., 10.0, 0.5, 0.0, 50)]:
    p = engine.make_params(np.pi * th / 180., tau, 300., 0.5, wvl0, s, k0, **kw)
    m = n if 'visible' not in name else max(256, n // 20)
    rec, tally, st = ctx.run(p, table, 7, 12345, m)
    assert tally[:, 0].sum() == m and (rec['condition'] >= 1).all() and (rec['condition'] <= 5).all()
    total += st['n_events']
    print('%-28s photons %7d events %9d' % (name, m, st['n_events']))
# both kernel paths and both ways a call ends, on short and on long walks
nir21 = gpu_util.fixture_table('spectral', 1000, 184, 236)
p21 = engine.make_params(np.pi * 30. / 180., 1e6, 300., 0.5, 2.1, sig, 184, n_theta_bins=137)
pv = engine.make_params(np.pi * 15. / 180., 10.0, 300., 0.5, 0.5, 0.0, 50, n_theta_bins=137)
for path in ('fused', 'persistent'):
    ctx.set_walk_path(path)
    for tail in (0, 1):
        ctx.set_tail_kernel(tail)
        for table, prm, m in ((nir21, p21, n), (rows_imp, engine.make_params(np.pi * 30. / 180., 3.0, 300., 0.5, 1.3, sig, 104, n_theta_bins=137), n),
                              (vis, pv, max(64, n // 50))):
            rec, tally, st = ctx.run(prm, table, 11, 1 << 33, m)
            assert tally[:, 0].sum() == m
            total += st['n_events']
        print('path %-10s tail kernel %d ok' % (path, tail))
ctx.set_walk_path('auto')
ctx.set_tail_kernel(-1)
# sweeps: a concatenated table that fits the shared-memory tally, and one that does not (rows of one case at a time on
# the persistent path, outcome counts only on the fused path)
for wvl0, radii in ((1.3, (100, 250)), (2.1, (50, 100, 250, 500, 1000)), (1.5, (50, 100, 250, 500, 1000))):
    k0 = int(round(wvl0 * 100)) - 26
    tabs = [gpu_util.fixture_table('spectral', r, k0, k0 + 52) for r in radii]
    cases = []
    for j in range(len(radii)):
        for th in (0., 60.):
            cases.append((engine.make_params(np.pi * th / 180., 1e6, 300., .5, wvl0, sig, k0, n_theta_bins=137), 53 * j, 53, n // 4 + 17 * len(cases)))
    for path in ('fused', 'persistent'):
        ctx.set_walk_path(path)
        per_case, tally, events, st = ctx.run_sweep(cases, np.concatenate(tabs), 3)
        assert tally[:, 0].sum() == sum(c[3] for c in cases) and int(events.sum()) == st['n_events']
        total += st['n_events']
    print('sweep wvl0 %.1f, %d cases, %d rows ok' % (wvl0, len(cases), 53 * len(radii)))
ctx.set_walk_path('auto')
# async slots + column histograms
p = engine.make_params(np.pi * 15. / 180., 1e6, 300., 0.5, 1.3, sig, 104, n_theta_bins=137)
ctx.set_histograms(200, (0., 500.), 1000, (0., 30.), 100.)
bufs = [engine.RecordBuffers(n) for _ in range(3)]
tallies = [np.zeros((len(rows), p.tally_width), np.uint64) for _ in range(3)]
for s in range(3):
    ctx.run_async(s, p, rows, 9, s * n, n, bufs[s], tallies[s])
for s in range(3):
    ctx.wait(s)
    ns, pl = ctx.histograms(s)
    assert ns.sum() <= n and tallies[s][:, 0].sum() == n
ctx.set_histograms()
# replay (fp64)
case = gu.load_case('slab_tau3_lb')
cfg = case['cfg']
pr = engine.make_params(np.pi * cfg['theta_0'] / 180., cfg['tau_tot'], cfg['rho_snw'], cfg['Lambertian_reflectance'], 0., 0., 0,
                        lambert_bottom=cfg['Lambertian_bottom'])
out = ctx.replay(pr, case['wvl'], case['ssa_ice'], case['ssa_imp'], case['g'], case['ext_cff_mss'], case['p_ext_imp'],
                 case['init_draws'], case['offsets'], case['stream'])
assert out['n_mismatch'] == 0
ctx.close()
print('sanitize workload ok, events', total)
