"""BASELINE.json configs[2] at its stated size: the NIR sweep 0.9-2.5 um (17 wavelengths) x 5 effective radii x 5
incident zenith angles (reference monte_carlo3D-run.py:60-96, 76: the loops a user writes around run()), 10^7 photons
per case = 425 cases, 4.25e9 photons, tallies only.

Two ways, same photons (case c = photon ids (c << 40) + j of one stream), results compared bit for bit:
  per case   one mc3d_run per case (what looping over run() does): kernel time, events/s and issue-roofline fraction
             of every case on its own;
  sweep      mc3d_run_sweep: the 25 cases of a wavelength (5 tables, 5 angles sharing each) in ONE set of launches, up
             to three wavelengths in flight.
usage: python tools/run_c3_grid.py [n_photon_per_case] [per-case: 0|1] [sweeps in flight: 3]      (needs a GPU; synthetic Mie tables)"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from monte_carlompi_b200 import engine, post, ssp, ssp_fixtures

FI_IMP = 'mie_sot_ChC90_dns_1317.nc'
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
per_case = (int(sys.argv[2]) if len(sys.argv) > 2 else 1) != 0
WVLS = [round(0.9 + 0.1 * k, 1) for k in range(17)]
RADII = (50, 100, 250, 500, 1000)
ANGLES = (0., 15., 30., 45., 60.)
SEED = 20190603
root = tempfile.mkdtemp(prefix='mc3d_c3_')
optics = ssp_fixtures.write_optics_dir(os.path.join(root, 'spectral'), 'spectral', RADII)
q = engine.query(0)
peak = q['sm_count'] * 128 * q['sm_clock_khz'] * 1e3 / 111.0
print('device: %s; issue roofline %.3e events/s (N_SM x 128 x f / 111, SURVEY.md 8d)' % (q, peak), flush=True)

ctx = engine.Context([0])
batches = []                      # one per wavelength: (table, [(params, row_begin, n_rows, n)], [(r, theta)])
for wvl0 in WVLS:
    scale = 0.085 / 2.355
    k_lo, k_hi = ssp.wavelength_grid(wvl0, scale)
    tables, cases, labels = [], [], []
    for j, r in enumerate(RADII):
        rows = ssp.build_table(optics, FI_IMP, r, k_lo, k_hi, 0.0, quiet=True)
        tables.append(rows)
        for th in ANGLES:
            P = engine.make_params(np.pi * th / 180., 1e6, 300., .5, wvl0, scale, k_lo, lambert_bottom=True, n_theta_bins=137)
            cases.append((P, j * len(rows), len(rows), n))
            labels.append((r, th))
    batches.append((np.concatenate(tables), cases, labels))

# ---- warm-up (buffers, clocks)
ctx.run_sweep(batches[8][1], batches[8][0], SEED, records=False)

# ---- per case
tallies_single = []
if per_case:
    print('\nper case (one mc3d_run each, %.0e photons, tallies only)' % n)
    print('%-30s %10s %10s %12s %8s %10s %8s' % ('case', 'ev/photon', 'kernel ms', 'events/s', 'frac', 'photons/s', 'albedo'))
    t_wall = time.perf_counter()
    k_ms_total, ev_total = 0.0, 0
    for b, (table, cases, labels) in enumerate(batches):
        acc = np.zeros((len(table), cases[0][0].tally_width), np.uint64)
        for c, ((P, rb, nr, m), (r, th)) in enumerate(zip(cases, labels)):
            _, t, st = ctx.run(P, table[rb:rb + nr], SEED + b, c << engine.SWEEP_ID_SHIFT, m, records=False)
            acc[rb:rb + nr] += t
            k_ms_total += st['kernel_ms']
            ev_total += st['n_events']
            eps = st['n_events'] / st['kernel_ms'] * 1e3
            print('wvl0=%.1f r=%-4d theta0=%-3d %12.2f %10.3f %12.3e %8.3f %10.3e %8.4f  %s' % (
                WVLS[b], r, th, st['n_events'] / m, st['kernel_ms'], eps, eps / peak, m / st['kernel_ms'] * 1e3,
                post.albedo_from_tally(t, table[rb:rb + nr]), {1: 'fused', 2: 'persistent'}.get(st.get('walk_path'), '')), flush=True)
        tallies_single.append(acc)
    t_wall = time.perf_counter() - t_wall
    print('per case total: %d cases, %.3e photons, %.3e events, kernel time %.3f s, wall %.3f s  -> %.3e photons/s, %.3e events/s (kernel)'
          % (len(WVLS) * 25, len(WVLS) * 25 * n, ev_total, k_ms_total / 1e3, t_wall, len(WVLS) * 25 * n / (k_ms_total / 1e3), ev_total / (k_ms_total / 1e3)))

# ---- sweep: one set of launches per wavelength, three wavelengths in flight
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pending = [None] * depth
print('\nsweep (mc3d_run_sweep: 25 cases per launch set, %d in flight)' % depth)
tallies_sweep = [None] * len(batches)
k_ms_sweep, ev_sweep = 0.0, 0
t_wall = time.perf_counter()


def finish(slot):
    global k_ms_sweep, ev_sweep
    b, tally, events = pending[slot]
    st = ctx.wait(slot)
    pending[slot] = None
    tallies_sweep[b] = tally
    k_ms_sweep += st['kernel_ms']
    ev_sweep += st['n_events']
    print('wvl0=%.1f: 25 cases, %.3e events, %.3f ms from first launch to last kernel end (%s), events/case min %.2f max %.2f per photon'
          % (WVLS[b], st['n_events'], st['kernel_ms'], {1: 'fused', 2: 'persistent'}.get(st.get('walk_path'), ''), events.min() / n, events.max() / n), flush=True)


for b, (table, cases, labels) in enumerate(batches):
    slot = b % depth
    if pending[slot] is not None:
        finish(slot)
    tally = np.zeros((len(table), cases[0][0].tally_width), np.uint64)
    events = np.zeros(len(cases), np.uint64)
    ctx.run_sweep_async(slot, cases, table, SEED + b, None, tally, events)
    pending[slot] = (b, tally, events)
for b in range(len(batches) - depth, len(batches)):
    if pending[b % depth] is not None:
        finish(b % depth)
t_wall = time.perf_counter() - t_wall
print('sweep total: %d cases, %.3e photons, %.3e events, wall %.3f s -> %.3e photons/s, %.3e events/s (%.3f of the issue roofline, wall clock)'
      % (len(WVLS) * 25, len(WVLS) * 25 * n, ev_sweep, t_wall, len(WVLS) * 25 * n / t_wall, ev_sweep / t_wall, ev_sweep / t_wall / peak))
if per_case:
    same = all(np.array_equal(a, b) for a, b in zip(tallies_single, tallies_sweep))
    print('sweep tallies == sum of the per-case tallies, all 17 wavelengths: %s' % ('bit-identical' if same else 'DIFFERENT'))
    if not same:
        sys.exit(1)
