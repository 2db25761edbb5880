"""Per-event latency of the end of a launch: a few visible-wavelength photons (thousands of events each) walked alone;
the kernel time divided by the longest walk is the time per event of one photon's dependent chain.
usage: [MC3D_TAIL=0|1] python tools/tail_latency.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import gpu_util
from monte_carlompi_b200 import engine
rows = gpu_util.fixture_table('const-vis', 100, 24, 76)
P = engine.make_params(np.pi * 15 / 180., 1e6, 300., .5, 0.5, 0.085 / 2.355, 24, lambert_bottom=True, n_theta_bins=137)
ctx = engine.Context([0])
ctx.set_walk_path('persistent')
for n in (1, 4, 32, 256, 4096):
    for seed in (1, 2, 3):
        rec, tally, st = ctx.run(P, rows, seed, 0, n)
        ns = rec['n_scat'].astype(np.int64)
        print('n %5d seed %d kernel_ms %8.3f  longest walk %8d events  -> %6.1f ns per event of the longest walk; all walks %9d events'
              % (n, seed, st['kernel_ms'], ns.max() + 1, 1e6 * st['kernel_ms'] / (ns.max() + 1), ns.sum() + n), flush=True)
