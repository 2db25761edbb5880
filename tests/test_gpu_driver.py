"""The drop-in driver surface end to end on a GPU: MonteCarlo().run(...) -> output file, read back the way
post_processing.py does."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(run_dir, optics_root, kind='spectral', **kw):
    from monte_carloMPI import monte_carlo3D
    return monte_carlo3D.MonteCarlo(optics_dir=optics_root[kind], output_dir=str(run_dir / 'monte_carlo_results'),
                                    devices=[0], **kw)


def test_default_run_writes_the_reference_file(run_dir, optics_root, capsys):
    pd = pytest.importorskip('pandas')
    mc = _model(run_dir, optics_root, seed=20190603)
    # the reference driver's defaults (monte_carlo3D-run.py:9-11, 18, 21, 48, 54, 84)
    mc.run(10000, 1.3, 0.085, 100., theta_0=15., Lambertian_bottom=True, Lambertian_reflectance=0.5)
    path = capsys.readouterr().out.strip().splitlines()[-1]                   # print(output_file), monte_carlo3D.py:1648
    assert path == os.path.join(str(run_dir / 'monte_carlo_results'), 'sphere', '1.3_0.085_100.0_10000_14.999999999999998_HG.txt')
    lines = open(path).read().splitlines()
    assert lines[0] == 'condition wvn[um^-1] theta_n phi_n n_scat path_length[m], snow_depth[m]' and len(lines) == 10001
    data = pd.read_csv(path, sep=r'\s+', float_precision='round_trip')        # post_processing.py:38
    rec = mc.last_records
    assert np.array_equal(data['condition'].values, rec['condition'])
    assert np.array_equal(data['n_scat'].values, rec['n_scat'])
    assert np.array_equal(data['theta_n'].values, rec['theta_n'].astype(np.float64))
    wvl = np.round(1.0 / data['wvn[um^-1]'].values, 2)
    assert abs(wvl.mean() - 1.3) < 4 * (0.085 / 2.355) / 100 and 0.03 < wvl.std() < 0.042
    assert np.allclose(data['snow_depth[m]'].values * data['wvn[um^-1]'].values * 0 + data['snow_depth[m]'].values,
                       1e6 / (mc.last_table['ext_cff_mss'][rec['wvl_row']] * 300.))
    # albedo the reference's way (calculate_albedo, monte_carlo3D.py:1659-1671)
    q_up = data[data.condition == 1]['wvn[um^-1]'].sum() / data['wvn[um^-1]'].sum()
    assert abs(q_up - mc.calculate_albedo()) < 1e-12 and 0.35 < q_up < 0.55
    # same seed -> same file; a second run gets the `_1` suffix (monte_carlo3D.py:135-141)
    mc.run(10000, 1.3, 0.085, 100., theta_0=15., Lambertian_bottom=True, Lambertian_reflectance=0.5)
    path2 = capsys.readouterr().out.strip().splitlines()[-1]
    assert path2 == path[:-4] + '_1.txt' and open(path2).read() == open(path).read()
    mc.close()


def test_known_answer_through_the_driver(run_dir, optics_root):
    from monte_carloMPI import monte_carlo3D
    mc = monte_carlo3D.MonteCarlo(tau_tot=2.0, imp_cnc=0, optics_dir=optics_root['const-kat'], output_dir=str(run_dir / 'o'),
                                  devices=[0], seed=3)
    mc.ssa_ice = 0.9                          # the reference's test() (monte_carlo3D.py:1849-1866)
    mc.g = 0.75
    n = 2000000
    mc.run(n, 0.5, 0.085, 100, test=True, Lambertian_bottom=False, write_output=False)
    albedo = mc.calculate_albedo()
    cond = mc.last_records['condition']
    trans = ((cond == 2) | (cond == 3)).mean()
    assert abs(albedo - 0.09739) < 3.5 * np.sqrt(0.09739 * 0.9 / n) + 2e-4
    assert abs(trans - 0.66096) < 3.5 * np.sqrt(0.66 * 0.34 / n)
    assert mc.last_tally[:, 0].sum() == n
    mc.close()


REFERENCE_RUN_SCRIPT_SHA256 = '02006f9a1e66265e10c5b0af5ae8d80f145d9f12d4ce69436047d49b6997c766'


def test_unmodified_reference_run_script_drives_the_package(run_dir, optics_root):
    """The reference's own user script (monte_carlo3D-run.py:4, 104-110: ``from monte_carloMPI import monte_carlo3D``,
    ``MonteCarlo().run(...)``), byte for byte as staged from the reference by build() into the git-ignored
    oracle/_ref/, run as a user would -- ``python monte_carlo3D-run.py`` in a directory with config.ini -- with this
    repository first on the module path: the import resolves to the B200 package, the run happens on the GPU, and the
    file it prints is the reference's (name, header, 10^4 rows, statistics of the driver-default golden case)."""
    import hashlib
    import subprocess
    import golden_util as gu
    script = os.path.join(ROOT, 'oracle', '_ref', 'reference', 'monte_carlo3D-run.py')
    if not os.path.isfile(script):
        pytest.skip('oracle/_ref/reference/monte_carlo3D-run.py not staged (build() stages it where /root/reference exists)')
    assert hashlib.sha256(open(script, 'rb').read()).hexdigest() == REFERENCE_RUN_SCRIPT_SHA256
    # a user keeps the script in the working directory, next to config.ini (python puts the script's own directory
    # first on sys.path: left in oracle/_ref/reference it would import the reference package staged beside it)
    import shutil
    shutil.copyfile(script, str(run_dir / 'monte_carlo3D-run.py'))
    script = str(run_dir / 'monte_carlo3D-run.py')
    assert hashlib.sha256(open(script, 'rb').read()).hexdigest() == REFERENCE_RUN_SCRIPT_SHA256
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''), CUDA_VISIBLE_DEVICES='0')
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, script, '--optics_dir', optics_root['spectral']], cwd=str(run_dir), env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    path = out.stdout.strip().splitlines()[-1]                                # print(output_file), monte_carlo3D.py:1648
    assert path == os.path.join('monte_carlo_results', 'sphere', '1.3_0.085_100.0_10000_14.999999999999998_HG.txt')
    lines = open(str(run_dir / path)).read().splitlines()
    assert lines[0] == 'condition wvn[um^-1] theta_n phi_n n_scat path_length[m], snow_depth[m]' and len(lines) == 10001
    cond = np.array([int(l.split()[0]) for l in lines[1:]])
    n_scat = np.array([int(l.split()[4]) for l in lines[1:]])
    gold = gu.load_case('c1_full_10k')['golden']                           # the same script run on the reference
    n = len(cond)
    for c in (1, 4):
        f, g = (cond == c).mean(), (gold['condition'] == c).mean()
        assert abs(f - g) < 4 * np.sqrt(2 * g * (1 - g) / n) + 1e-3, (c, f, g)
    assert abs(n_scat.mean() - gold['n_scat'].mean()) < 4 * np.sqrt(2) * gold['n_scat'].std() / np.sqrt(n)


def test_sweep_equals_case_by_case_runs(run_dir, optics_root, capsys):
    # reference monte_carlo3D-run.py:60-96: wavelengths x grain sizes, one run() each; here one launch per batch of cases
    # (mc3d_run_sweep), case c being photons (c << 40) + j of the sweep's stream
    grid = [(1.3, 0.085, 50, 15.), (1.3, 0.085, 100, 15.), (1.55, 0.130, 250, 0.), (1.55, 0.130, 500, 30.),
            (0.9, 0.085, 1000, 60.), (2.2, 0.085, 100, 45.), (1.0, 0.085, 250, 15.), (1.3, 0.085, 1000, 15.),
            (1.8, 0.26, 100, 15.), (1.3, 1e-12, 100, 15.), (1.3, 0.085, 100, 60.)]      # the last shares case 1's rows
    cases = [dict(n_photon=30000 + 1000 * k, wvl0=w, half_width=hw, rds_snw=r, theta_0=th, Lambertian_bottom=True,
                  Lambertian_reflectance=0.5) for k, (w, hw, r, th) in enumerate(grid)]
    # an aspherical habit in the same sweep (needs HG=True): other table loader, other output directory
    cases.insert(3, dict(n_photon=25000, wvl0=1.55, half_width=0.130, rds_snw=120, theta_0=30.,
                         shape='solid hexagonal column', roughness='smooth'))
    cases[6]['seed'] = 77                                    # a case with its own stream (starts a new batch)
    a = _model(run_dir, optics_root, tau_tot=8.0, HG=True)
    a.output_dir = str(run_dir / 'sweep')
    paths = a.run_sweep(cases, seed=100, write_output='both')
    assert a.last_seed == 100
    assert len(paths) == len(cases) and len(set(paths)) == len(cases)
    assert os.path.join('sweep', 'solid_column', 'Rough000') in paths[3] and os.path.join('sweep', 'sphere') in paths[4]
    results = a.run_sweep(cases, seed=100, write_output=False)
    a.close()
    from monte_carlompi_b200 import output
    b = _model(run_dir, optics_root, tau_tot=8.0, HG=True)
    b.output_dir = str(run_dir / 'single')
    for k, (c, p) in enumerate(zip(cases, paths)):
        kw = {key: v for key, v in c.items() if key not in ('n_photon', 'wvl0', 'half_width', 'rds_snw')}
        kw.setdefault('seed', 100)
        b.run(c['n_photon'], c['wvl0'], c['half_width'], c['rds_snw'], first_photon_id=k << 40, **kw)
        q = capsys.readouterr().out.strip().splitlines()[-1]
        assert os.path.basename(p) == os.path.basename(q)
        assert open(p).read() == open(q).read()
        rec, tally, table = results[k]
        assert np.array_equal(tally, b.last_tally), k      # also for the two cases that share rows in one launch
        assert np.array_equal(table, b.last_table)
        for col in rec:
            assert np.array_equal(rec[col], b.last_records[col]), (k, col)
        z = np.load(output.sidecar_path(p))
        assert np.array_equal(z['tally'], b.last_tally) and np.array_equal(z['n_scat'], rec['n_scat'])
    b.close()


def test_aspherical_habit_with_hg(run_dir, optics_root, capsys):
    # run(shape='droxtal') under --HG: same walk, SSPs from the habit's isca.dat (monte_carlo3D.py:1529-1545), output
    # under <output_dir>/<shape_dir>/<roughness_dir>/ with the size class' effective radius in the name (:105-131)
    import golden_util as gu
    case = gu.load_case('aspherical_hg')
    cfg = case['cfg']
    mc = _model(run_dir, optics_root, tau_tot=cfg['tau_tot'], imp_cnc=cfg['imp_cnc'], HG=True, seed=4)
    n = 400000
    mc.run(n, cfg['wvl0'], cfg['half_width'], cfg['rds_snw'], theta_0=cfg['theta_0'], shape=cfg['shape'],
           roughness=cfg['roughness'], Lambertian_bottom=True, Lambertian_reflectance=0.5)
    path = capsys.readouterr().out.strip().splitlines()[-1]
    assert path == os.path.join(str(run_dir / 'monte_carlo_results'), 'droxtal', 'Rough003',
                                '1.3_0.26_75.0_%d_29.999999999999996_HG.txt' % n)
    assert mc.snow_effective_radius == 75.0
    # rows == what the reference derived; outcome fractions within 4 sigma of the reference's 1500-photon golden
    rows_k = case['rows_k']
    mine = mc.last_table[rows_k - _k_first(mc)]       # table row r <-> drawn wavelength (k_first + r) / 100 um
    for col in case['rows'].dtype.names:
        assert np.array_equal(mine[col], case['rows'][col]), col
    gold = np.bincount(case['golden']['condition'], minlength=6)[1:] / cfg['n_photon']
    got = np.bincount(mc.last_records['condition'], minlength=6)[1:] / n
    sigma = np.sqrt(np.maximum(got * (1 - got), 1e-6) / cfg['n_photon'])
    assert (np.abs(got - gold) < 4 * sigma + 1e-3).all(), (got, gold)
    # the file's wvn column holds library wavelengths (multiples of 0.05 um), like the reference's
    first = open(path).read().splitlines()[1:200]
    wv = np.array([1.0 / float(l.split()[1]) for l in first])
    assert np.allclose(wv * 20, np.round(wv * 20), atol=1e-9)
    mc.close()


def _k_first(mc):
    from monte_carlompi_b200 import ssp
    return ssp.wavelength_grid(mc.wvl0, 0.26 / 2.355)[0]


def test_binary_sidecar_of_a_run(run_dir, optics_root, capsys):
    from monte_carlompi_b200 import output
    mc = _model(run_dir, optics_root, seed=8, tau_tot=5.0)
    mc.run(20000, 1.3, 0.085, 100., theta_0=15., Lambertian_reflectance=0.5, write_output='both')
    txt = capsys.readouterr().out.strip().splitlines()[-1]
    a, b = output.load_run(txt), output.load_run(output.sidecar_path(txt))
    for c in a.columns:
        assert np.array_equal(a[c].values, b[c].values), c
    z = np.load(output.sidecar_path(txt))
    assert np.array_equal(z['tally'], mc.last_tally) and z['tally'][:, 0].sum() == 20000
    mc.run(20000, 1.3, 0.085, 100., theta_0=15., Lambertian_reflectance=0.5, write_output='binary')
    npz = capsys.readouterr().out.strip().splitlines()[-1]
    assert npz.endswith('_HG_1.npz') and not os.path.exists(npz[:-4] + '.txt')      # same de-duplicated run name
    c = output.load_run(npz)
    assert np.array_equal(c['n_scat'].values, a['n_scat'].values)
    mc.close()
