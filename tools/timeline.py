"""Per-warp timeline of one isolated walk-kernel launch (debug build: make -C monte_carlompi_b200/csrc timeline).
Prints when warps entered, exhausted the fresh list and left, relative to the first entry.
usage: MC3D_LIB=monte_carlompi_b200/libmc3d_timeline.so python tools/timeline.py [n_photon] [bps block thr] [kind]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from monte_carlompi_b200 import engine
sys.argv = sys.argv[:1] + sys.argv[1:]
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
kind = sys.argv[5] if len(sys.argv) > 5 else 'spectral'
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import gpu_util
wvl0 = 0.5 if kind == 'const-vis' else 1.3
k0 = int(round(wvl0 * 100)) - 26
rows = gpu_util.fixture_table(kind, 100, k0, k0 + 52)
ctx = engine.Context([0])
ctx.set_walk_path('persistent')
if len(sys.argv) > 4:
    ctx.set_launch(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
P = engine.make_params(np.pi * 15 / 180., 1e6, 300., .5, wvl0, 0.085 / 2.355, k0, lambert_bottom=True, n_theta_bins=137)
lib = engine.load_library()
for rep in range(3):
    rec, tally, st = ctx.run(P, rows, 777 + rep, 0, n, records=True)
    n_warps = st['grid_blocks'] * st['block_threads'] // 32
    buf = np.zeros(3 * n_warps, np.uint64)
    assert lib.mc3d_debug_timeline(buf.ctypes.data_as(C.c_void_p), buf.size) == 0
    t = buf.reshape(-1, 3).astype(np.int64)
    t0 = t[:, 0].min()
    enter, exh, leave = [(t[:, k] - t0) / 1e3 for k in range(3)]
    q = lambda a: ' '.join('%7.1f' % v for v in np.percentile(a, [0, 10, 50, 90, 99, 100]))
    print('n %d grid %d kernel_ms %.3f max n_scat %d | us since first entry, percentiles 0/10/50/90/99/100' % (n, st['grid_blocks'], st['kernel_ms'], rec['n_scat'].max()))
    print('   enter     ', q(enter))
    print('   exhausted ', q(exh))
    print('   leave     ', q(leave))
    print('   drain time', q(leave - exh))
