"""Run the production walk a few times (for ncu / quick timing).  usage: profile_walk.py [n_photon] [reps] [kind]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from monte_carlompi_b200 import engine, ssp_fixtures

def table_for(kind, r, k_lo, k_hi):
    wvl, ssa, ext, g = ssp_fixtures.ice_table(kind, r)
    rows = np.zeros(k_hi - k_lo + 1, engine.ROW_DTYPE)
    for j, k in enumerate(range(k_lo, k_hi + 1)):
        w = k / 100.0
        i = int(np.argmin(np.abs(wvl * 1e6 - w)))
        rows[j] = (w, ssa[i], 0.3, g[i], ext[i], 0.0)
    return rows

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else 'spectral'
wvl0 = 0.5 if kind == 'const-vis' else 1.3
k0 = int(round(wvl0 * 100)) - 26
rows = table_for(kind, 100, k0, k0 + 52)
ctx = engine.Context([0])
if len(sys.argv) > 6:
    ctx.set_launch(int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))
P = engine.make_params(np.pi * 15 / 180., 1e6, 300., .5, wvl0, 0.085 / 2.355, k0, lambert_bottom=True, n_theta_bins=137)
for rep in range(reps):
    rec, tally, st = ctx.run(P, rows, 777 + rep, 0, n, records=False)
    print('n', n, 'kernel_ms %.3f' % st['kernel_ms'], 'events/s %.4e' % (st['n_events'] / st['kernel_ms'] * 1e3),
          'photons/s %.4e' % (n / st['kernel_ms'] * 1e3), 'events/photon %.1f' % (st['n_events'] / n))
