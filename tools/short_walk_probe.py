"""Short walks: persistent path at several refill thresholds vs the fused kernel (kernel ms for n photons, tallies only).
usage: python tools/short_walk_probe.py [n_photon] [only_case_index]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import gpu_util
from monte_carlompi_b200 import engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
only = int(sys.argv[2]) if len(sys.argv) > 2 else None
ctx = engine.Context([0])
cases = ((1.3, 1000, 0.), (1.7, 50, 0.), (1.7, 250, 30.), (2.1, 50, 0.), (2.1, 1000, 0.))
for ci, (wvl0, r, th) in enumerate(cases):
    if only is not None and ci != only:
        continue
    k0 = int(round(wvl0 * 100)) - 26
    rows = gpu_util.fixture_table('spectral', r, k0, k0 + 52)
    P = engine.make_params(np.pi * th / 180., 1e6, 300., .5, wvl0, 0.085 / 2.355, k0, lambert_bottom=True, n_theta_bins=137)
    line = 'wvl0=%.1f r=%-4d' % (wvl0, r)
    for path, thr in (('fused', 4), ('persistent', 4), ('persistent', 8), ('persistent', 12), ('persistent', 16), ('persistent', 24)):
        ctx.set_walk_path(path)
        ctx.set_launch(0, 0, thr)
        best = None
        for rep in range(3 if only is None else 1):
            _, t, st = ctx.run(P, rows, 20190603, 0, n, records=False)
            best = st['kernel_ms'] if best is None else min(best, st['kernel_ms'])
        line += '  %s%s %.3f' % (path[0], '' if path == 'fused' else thr, best)
    print(line, ' (%.1f events/photon)' % (st['n_events'] / n), flush=True)
