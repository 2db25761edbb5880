"""Short-walk (fused, one kernel) vs persistent (three kernels) path over a sample of config C3: kernel time of 1e7
photons, tallies only and with packed records, and whether the two paths agree bit for bit.
usage: python tools/fused_vs_persistent.py [n_photon]   (needs a GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import gpu_util
from monte_carlompi_b200 import engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
ctx = engine.Context([0])
buf = engine.RecordBuffers(n)
print('%-28s %8s | %-30s | %-30s | %s' % ('case', 'ev/phot', 'tallies only: fused / persist ms', 'packed records: fused / persist ms', 'auto picks'))
for wvl0, r, th in ((1.3, 50, 0.), (1.3, 250, 30.), (1.3, 1000, 0.), (1.7, 50, 0.), (1.7, 250, 30.), (1.7, 1000, 60.), (2.1, 50, 0.),
                    (2.1, 1000, 0.), (2.5, 50, 60.), (2.5, 1000, 0.)):
    k0 = int(round(wvl0 * 100)) - 26
    rows = gpu_util.fixture_table('spectral', r, k0, k0 + 52)
    P = engine.make_params(np.pi * th / 180., 1e6, 300., .5, wvl0, 0.085 / 2.355, k0, lambert_bottom=True, n_theta_bins=137)
    out = {}
    for path in ('fused', 'persistent'):
        ctx.set_walk_path(path)
        for records in (False, True):
            best = None
            for rep in range(3):
                t = np.zeros((len(rows), P.tally_width), np.uint64)
                st = ctx.run_sync(P, rows, 20190603, 0, n, buf if records else None, t)
                best = st['kernel_ms'] if best is None else min(best, st['kernel_ms'])
            out[path, records] = (best, t.copy(), buf.packed(n).copy() if records else None, st['n_events'])
    ctx.set_walk_path('auto')
    t = np.zeros((len(rows), P.tally_width), np.uint64)
    auto = ctx.run_sync(P, rows, 20190603, 0, n, None, t)['walk_path']
    same = np.array_equal(out['fused', True][1], out['persistent', True][1]) and np.array_equal(out['fused', True][2], out['persistent', True][2])
    print('wvl0=%.1f r=%-4d theta0=%-3d %10.1f | %8.3f / %8.3f  (x%.2f)    | %8.3f / %8.3f  (x%.2f)    | %s %s' % (
        wvl0, r, th, out['fused', False][3] / n, out['fused', False][0], out['persistent', False][0],
        out['persistent', False][0] / out['fused', False][0], out['fused', True][0], out['persistent', True][0],
        out['persistent', True][0] / out['fused', True][0], {1: 'fused', 2: 'persistent'}[auto], 'bit-identical' if same else 'DIFFERENT'), flush=True)
