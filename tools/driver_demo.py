"""End-to-end through the drop-in driver surface on a GPU box: default run (n=1e4), then n=1e6 with a timing
breakdown, then BRF from the written file (the reference's post_processing formulas) vs BRF from the GPU tallies."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

work = tempfile.mkdtemp(prefix='mc3d_demo_')
import shutil
shutil.copy(os.path.join(ROOT, 'config.ini'), work)
env = dict(os.environ, PYTHONPATH=ROOT, MC3D_SYNTHETIC_OPTICS='1')
t0 = time.time()
out = subprocess.run([sys.executable, os.path.join(ROOT, 'monte_carlo3D-run.py')], cwd=work, env=env, capture_output=True, text=True)
print('monte_carlo3D-run.py (default, n_photon=10000): %.2f s wall incl. interpreter start' % (time.time() - t0))
print('  stdout:', out.stdout.strip(), '| stderr tail:', out.stderr.strip()[-200:])
path = out.stdout.strip().splitlines()[-1]
print('  file %s: %d lines, header %r' % (os.path.basename(path), sum(1 for _ in open(os.path.join(work, path))), open(os.path.join(work, path)).readline()))

os.chdir(work)
sys.argv = ['monte_carlo3D-run.py']
from monte_carloMPI import monte_carlo3D
from monte_carlompi_b200 import post
mc = monte_carlo3D.MonteCarlo(seed=1)
t0 = time.time(); mc.run(1000, 1.3, 0.085, 100., theta_0=15., Lambertian_reflectance=0.5); print('warm-up run: %.3f s' % (time.time() - t0))
t0 = time.time()
mc.run(1000000, 1.3, 0.085, 100., theta_0=15., Lambertian_bottom=True, Lambertian_reflectance=0.5)
t_all = time.time() - t0
print('MonteCarlo.run(n_photon=1e6): %.3f s total (GPU kernels %.2f ms, library call %.2f ms; the rest is the table lookup and the 102 MB text file)'
      % (t_all, mc.last_stats['kernel_ms'], mc.last_stats['total_ms']))
import pandas as pd
files = sorted(os.listdir(os.path.join('monte_carlo_results', 'sphere')))
f = [x for x in files if '_1000000_' in x][0]
data = pd.read_csv(os.path.join('monte_carlo_results', 'sphere', f), sep=r'\s+', float_precision='round_trip')
mid_f, brf_f = post.brf_from_records(data['condition'].values, data['wvn[um^-1]'].values, data['theta_n'].values, 137)
mid_t, brf_t = post.brf_from_tally(mc.last_tally, mc.last_table)
print('BRF from file (post_processing.py formulas) vs from GPU tallies: max |diff| = %.3e over 137 bins; albedo %.5f vs %.5f'
      % (np.abs(brf_f - brf_t).max(), data[data.condition == 1]['wvn[um^-1]'].sum() / data['wvn[um^-1]'].sum(), post.albedo_from_tally(mc.last_tally, mc.last_table)))
mc.close()
