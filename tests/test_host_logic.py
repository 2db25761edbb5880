"""CPU tests of the host side: SSP lookup, output naming / writing, work partition, rendezvous, config surface."""
import contextlib
import io
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import golden_util as gu
from monte_carlompi_b200 import output, parallelize, ssp


# ---- SSP lookup: bit-identical to the arrays the reference derives (monte_carlo3D.py:498-777, 1575-1588) ------
@pytest.mark.parametrize('name,radius,imp_cnc', [('c1_default', 100, 0.0), ('edge_of_table', 100, 0.0),
                                                  ('impurity', 100, 1e-5), ('slab_tau3_black', 250, 0.0)])
def test_ssp_table_matches_reference(optics_root, name, radius, imp_cnc):
    rows = gu.load_case(name)['rows']
    k = np.rint(rows['wvl_um'] * 100).astype(int)
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        full = ssp.build_table(optics_root['spectral'], 'mie_sot_ChC90_dns_1317.nc', radius, k.min(), k.max(), imp_cnc)
    mine = full[k - k.min()]
    for col in rows.dtype.names:
        assert np.array_equal(mine[col], rows[col]), col
    # out-of-table wavelengths take the nearest row and warn like the reference (monte_carlo3D.py:536-537)
    assert ('using nearest value instead' in err.getvalue()) == (name == 'edge_of_table')


def test_aspherical_hg_table_matches_reference(optics_root):
    # get_aspherical_SSPs with --HG (monte_carlo3D.py:173-266, 316-336, 396-415): rows the reference derived, per
    # drawn wavelength, from the same synthetic isca.dat (oracle/make_golden.py: make_aspherical_case)
    case = gu.load_case('aspherical_hg')
    cfg, rows, rows_k = case['cfg'], case['rows'], case['rows_k']
    full, r_eff = ssp.build_table_aspherical(optics_root['spectral'], 'mie_sot_ChC90_dns_1317.nc', cfg['shape'],
                                             cfg['roughness'], cfg['wvl0'], cfg['rds_snw'], rows_k.min(), rows_k.max(),
                                             cfg['imp_cnc'], cfg['rho_ice'], quiet=True)
    mine = full[rows_k - rows_k.min()]
    for col in rows.dtype.names:
        assert np.array_equal(mine[col], rows[col]), col
    assert r_eff == 75.0                                           # nearest size class to the requested 80 um
    # wavelengths are replaced by the library's nearest ones (0.05 um steps here), not interpolated
    assert set(np.round(rows['wvl_um'] * 100).astype(int) % 5) == {0} and len(np.unique(rows['wvl_um'])) < len(rows)
    # ... and the effective radius of the size class goes into the file name (monte_carlo3D.py:118, 131)
    name = output.run_name(cfg['wvl0'], cfg['half_width'], r_eff, cfg['n_photon'], np.pi * cfg['theta_0'] / 180.)
    assert name == case['file_name'] == '1.3_0.26_75.0_1500_29.999999999999996_HG.txt'


def test_aspherical_library_errors(optics_root):
    args = (optics_root['spectral'], 'mie_sot_ChC90_dns_1317.nc')
    with pytest.raises(ValueError, match='not a wavelength'):      # the reference needs wvl0 to be a library member
        ssp.build_table_aspherical(*args, 'droxtal', 'moderately rough', 1.31, 80., 120, 140, 0.0, 917.)
    with pytest.raises(ValueError, match='outside'):
        ssp.aspherical_dirs('droxtal', 'smooth', 15.8)
    with pytest.raises(ValueError, match='unknown shape'):
        ssp.aspherical_dirs('cube', 'smooth', 1.3)
    assert ssp.aspherical_dirs('8-element column aggregate', 'severely rough', 20.0) == ('16.4-99.0', 'column_8elements', 'Rough050')


def test_setup_output_nested_dirs_for_aspherical(tmp_path):
    p = output.setup_output(str(tmp_path / 'res'), 1.3, 0.085, 75.0, 10, 0.0, shape_dir=('droxtal', 'Rough003'))
    assert p == os.path.join(str(tmp_path / 'res'), 'droxtal', 'Rough003', '1.3_0.085_75.0_10_0.0_HG.txt')
    assert os.path.isdir(os.path.dirname(p))


def test_ssp_nearest_row_outside_table(optics_root):
    t = ssp.read_table(ssp.ice_file(optics_root['spectral'], 100), ('wvl', 'ss_alb'))
    lo = ssp.nearest_pair_interp(t['wvl'], {'a': t['ss_alb']}, [0.30, 0.305, 0.31, 4.995, 5.0, 9.0])['a']
    assert lo[0] == t['ss_alb'][0] and lo[1] == t['ss_alb'][0]
    assert lo[2] == 0.5 * t['ss_alb'][0] + 0.5 * t['ss_alb'][1] or abs(lo[2] - t['ss_alb'][:2].mean()) < 1e-15
    assert lo[3] == t['ss_alb'][-1] and lo[4] == t['ss_alb'][-1] and lo[5] == t['ss_alb'][-1]


def test_test_hook_overrides(optics_root):
    rows = ssp.build_table(optics_root['const-kat'], 'mie_sot_ChC90_dns_1317.nc', 100, 40, 60, 0.0,
                           overrides={'ssa_ice': 0.9, 'g': 0.75}, quiet=True)
    assert (rows['ssa_ice'] == 0.9).all() and (rows['g'] == 0.75).all() and (rows['p_ext_imp'] == 0).all()


def test_wavelength_grid_covers_seven_sigma():
    k_lo, k_hi = ssp.wavelength_grid(1.3, 0.085 / 2.355)
    assert k_lo <= 130 - 26 and k_hi >= 130 + 26 and k_hi - k_lo < 256
    k_lo, k_hi = ssp.wavelength_grid(1.3, 1e-15 / 2.355)          # monochromatic: a handful of rows around 1.30
    assert k_lo <= 130 <= k_hi and k_hi - k_lo <= 4
    assert ssp.wavelength_grid(0.05, 0.26 / 2.355)[0] == 1      # wavelengths stay positive


# ---- output: file name and bytes (monte_carlo3D.py:96-143, 1621-1648) ------------------------------------------
def test_output_bytes_match_reference():
    z = np.load(os.path.join(gu.GOLDEN_DIR, 'text_default.npz'))
    body = output.format_lines(z['condition'], z['wvn'], z['theta_n'], z['phi_n'], z['n_scat'], z['path_length'],
                               z['snow_depth'])
    assert output.HEADER + body == str(z['text'])
    assert output.run_name(1.3, 0.085, 100., 40, np.pi * 15. / 180.) == str(z['name'])


def test_native_writer_is_byte_identical_to_python_repr(tmp_path):
    from monte_carlompi_b200 import engine
    rng = np.random.RandomState(1)
    vals = np.concatenate([rng.uniform(0, 7, 5000), 10.0 ** rng.uniform(-12, 20, 5000),
                           rng.standard_normal(2000).astype(np.float32).astype(np.float64),
                           [0.0, -0.0, 1.0, 1e16, 1e-4, 9.999e-5, 123456789012345680.0, 1e22, 5e-324, 1.7976931348623157e308,
                            0.1, 100.0, 1e15, 9999999999999998.0, 1e-5, 0.7692307692307693, float('inf'), -float('inf')]])
    for v in vals:
        assert engine.py_repr(v) == repr(float(v)), v
    assert engine.py_repr(float('nan')) == 'nan'
    n = 70001                                   # more than one 65536-line work item, ragged tail
    rec = dict(condition=rng.randint(1, 6, n).astype(np.uint8), wvl_row=rng.randint(0, 53, n).astype(np.int16),
               theta_n=rng.uniform(0, np.pi, n).astype(np.float32), phi_n=rng.uniform(0, 6.28, n).astype(np.float32),
               n_scat=rng.randint(0, 500000, n).astype(np.uint32), path_length=rng.exponential(.01, n).astype(np.float32))
    rec['phi_n'][::7] = 0.0                     # unscattered photons print 0.0
    wvn, depth = 1 / np.round(rng.uniform(1.0, 1.6, 53), 2), rng.uniform(100, 300, 53)
    path = output.write_run(str(tmp_path / 'native.txt'), rec, wvn, depth)
    want = output.HEADER + output.format_lines(rec['condition'], wvn[rec['wvl_row']], rec['theta_n'], rec['phi_n'],
                                               rec['n_scat'], rec['path_length'], depth[rec['wvl_row']])
    assert open(path).read() == want
    z = np.load(os.path.join(gu.GOLDEN_DIR, 'text_default.npz'))   # and against the reference's own bytes
    assert engine.py_repr(float(z['theta_n'][3])) in str(z['text'])


def test_run_name_theta_round_trip_quirk():
    name = lambda deg: output.run_name(1.3, 0.085, 100., 10000, np.pi * deg / 180.)
    assert name(15.) == '1.3_0.085_100.0_10000_14.999999999999998_HG.txt'
    assert name(60.).split('_')[4] == '59.99999999999999'
    assert name(0.).split('_')[4] == '0.0' and name(45.).split('_')[4] == '45.0'


def test_setup_output_dedup_suffix(tmp_path):
    d = str(tmp_path / 'out')
    p0 = output.setup_output(d, 1.3, 0.085, 100., 10, 0.0)
    open(p0, 'w').close()
    p1 = output.setup_output(d, 1.3, 0.085, 100., 10, 0.0)
    open(p1, 'w').close()
    p2 = output.setup_output(d, 1.3, 0.085, 100., 10, 0.0)
    assert os.path.dirname(p0).endswith(os.path.join('out', 'sphere'))
    assert p1 == p0[:-4] + '_1.txt' and p2 == p0[:-4] + '_2.txt'


def test_written_file_reads_back_like_post_processing(tmp_path):
    pd = pytest.importorskip('pandas')
    n = 1000
    rng = np.random.RandomState(0)
    cols = dict(condition=rng.randint(1, 6, n), wvn=1 / np.round(rng.normal(1.3, .04, n), 2),
                theta_n=rng.uniform(0, np.pi, n).astype(np.float32), phi_n=rng.uniform(0, 6.28, n).astype(np.float32),
                n_scat=rng.randint(0, 5000, n), path_length=rng.exponential(.01, n).astype(np.float32),
                snow_depth=np.full(n, 201.8))
    path = output.write_records(str(tmp_path / 'f.txt'), **cols)
    data = pd.read_csv(path, sep=r'\s+', float_precision='round_trip')    # post_processing.py:38 (exact parser)
    assert list(data.columns) == ['condition', 'wvn[um^-1]', 'theta_n', 'phi_n', 'n_scat', 'path_length[m],',
                                  'snow_depth[m]']
    assert np.array_equal(data['condition'].values, cols['condition'])
    assert np.array_equal(data['theta_n'].values, cols['theta_n'].astype(np.float64))      # exact round trip
    assert np.array_equal(data['wvn[um^-1]'].values, cols['wvn'])


def test_binary_sidecar_reads_back_like_the_text_file(tmp_path):
    pd = pytest.importorskip('pandas')
    n = 2000
    rng = np.random.RandomState(1)
    rec = dict(condition=rng.randint(1, 6, n).astype(np.uint8), wvl_row=rng.randint(0, 53, n).astype(np.int16),
               theta_n=rng.uniform(0, np.pi, n).astype(np.float32), phi_n=rng.uniform(0, 6.28, n).astype(np.float32),
               n_scat=rng.randint(0, 5000, n).astype(np.uint32), path_length=rng.exponential(.01, n).astype(np.float32))
    wvn, depth = 1. / (np.arange(104, 157) / 100.), 1e6 / (np.linspace(16, 17, 53) * 300.)
    txt = output.write_run(str(tmp_path / 'r.txt'), rec, wvn, depth)
    npz = output.write_sidecar(txt, rec, wvn, depth, tally=np.zeros((53, 145), np.uint64))
    assert npz == str(tmp_path / 'r.npz')
    a, b = output.load_run(txt), output.load_run(npz)
    assert list(a.columns) == list(b.columns) == list(output.COLUMNS)
    for c in a.columns:
        assert a[c].dtype == b[c].dtype and np.array_equal(a[c].values, b[c].values), c
    assert os.path.getsize(npz) - 53 * 145 * 8 < 0.3 * os.path.getsize(txt)      # 19 B vs ~100 B per photon


# ---- partition / ranks (parallelize.py:14-15) -----------------------------------------------------------------
@pytest.mark.parametrize('n,parts', [(10, 3), (1000000, 8), (7, 8), (0, 4), (33, 1), (10**9, 8)])
def test_partition_is_array_split(n, parts):
    got = parallelize.partition(n, parts)
    if n <= 10**6:
        ref = np.array_split(np.arange(n), parts)
        assert [c for _, c in got] == [len(r) for r in ref]
        assert [b for b, c in got if c] == [int(r[0]) for r in ref if len(r)]
    assert sum(c for _, c in got) == n and got[0][0] == 0
    assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(parts - 1))


def test_detect_ranks():
    assert parallelize.detect_ranks({}) == (0, 1, 0)
    assert parallelize.detect_ranks({'RANK': '3', 'WORLD_SIZE': '8', 'LOCAL_RANK': '3'}) == (3, 8, 3)
    assert parallelize.detect_ranks({'OMPI_COMM_WORLD_RANK': '1', 'OMPI_COMM_WORLD_SIZE': '2'}) == (1, 2, 1)


def _rdzv_worker(rank, root, q):
    r = parallelize.FileRendezvous(rank, 2, token='t', root=root, timeout=30)
    if rank == 0:
        r.put('nccl_id', bytes(range(128)))
    blob = r.get('nccl_id')
    r.barrier('b')
    q.put((rank, blob == bytes(range(128))))


def _allgather_worker(rank, root, q):
    par = parallelize.Parallel(100, environ={'RANK': str(rank), 'WORLD_SIZE': '2', 'LOCAL_RANK': str(rank)})
    par.rendezvous(root=root, token='ag', timeout=30)                                       # what open() sets up
    ext = np.array([3. + rank, 40. - rank, 0.5 * (rank + 1), 9. + rank])                  # per-rank extrema
    parts = [np.frombuffer(b, np.float64) for b in par.allgather_bytes(ext.tobytes())]
    again = par.allgather_bytes(b'x%d' % rank)                                             # tags do not collide
    if rank == 1:
        import time
        time.sleep(1.0)                        # rank 0 reaches close() first and must wait for rank 1, not vice versa
    t0 = __import__('time').time()
    par.close()
    assert __import__('time').time() - t0 < 10.0
    q.put((rank, [p.tolist() for p in parts], again))


def test_allgather_bytes_two_processes(tmp_path):
    # MonteCarlo.histograms combines the per-rank extrema this way (one process per GPU)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_allgather_worker, args=(r, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=60) for _ in ps)
    [p.join(30) for p in ps]
    want = [[3., 40., 0.5, 9.], [4., 39., 1.0, 10.]]
    assert got == [(0, want, [b'x0', b'x1']), (1, want, [b'x0', b'x1'])]
    assert not [d for d in os.listdir(str(tmp_path)) if d.startswith('mc3d_rdzv_')]       # rank 0 cleaned up after both left
    assert parallelize.Parallel(5, environ={}).allgather_bytes(b'solo') == [b'solo']


def _two_instances_worker(rank, root, q):
    """Two Parallel objects one after the other in the same job (the driver's multiple_wavelengths loop, or one
    MonteCarlo per run): each gets its own rendezvous directory, so the second never reads the first one's NCCL id,
    and closing the first -- at different times on the two ranks -- never removes a file of the second."""
    import time
    env = {'RANK': str(rank), 'WORLD_SIZE': '2', 'LOCAL_RANK': str(rank)}
    got = []
    for k in range(2):
        par = parallelize.Parallel(1000 + k, environ=env)
        rdzv = par.rendezvous(root=root, token='job', timeout=30)
        if rank == 0:
            time.sleep(0.3 * k)                                         # rank 1 runs ahead into instance 2
            rdzv.put('nccl_id', b'id-of-instance-%d' % k)
        got.append(rdzv.get('nccl_id'))
        got.append(par.broadcast_seed(1234567890123456789 + k if rank == 0 else 7))
        cols = {'condition': np.full(3 + rank, rank, np.uint8), 'theta_n': np.arange(3 + rank, dtype=np.float32) + 10 * rank}
        ans = par.answer_and_reduce(cols, lambda parts: {c: np.concatenate([p[c] for p in parts]) for c in parts[0]})
        got.append(None if ans is None else {c: v.tolist() for c, v in ans.items()})
        if rank == 0:
            time.sleep(0.3)                                             # rank 0 closes late
        par.close()
    q.put((rank, got))


def test_two_parallel_instances_in_one_job(tmp_path):
    shm_before = set(f for f in os.listdir('/dev/shm') if f.startswith('mc3d_'))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_two_instances_worker, args=(r, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in ps]
    got = dict(q.get(timeout=90) for _ in ps)
    [p.join(30) for p in ps]
    gathered = {'condition': [0, 0, 0, 1, 1, 1, 1], 'theta_n': [0., 1., 2., 10., 11., 12., 13.]}
    for rank in (0, 1):
        ans = gathered if rank == 0 else None
        assert got[rank] == [b'id-of-instance-0', 1234567890123456789, ans, b'id-of-instance-1', 1234567890123456790, ans]
    assert not [d for d in os.listdir(str(tmp_path)) if d.startswith('mc3d_rdzv_')]
    assert set(f for f in os.listdir('/dev/shm') if f.startswith('mc3d_')) <= shm_before                                        # segments unlinked by their owners


def test_file_rendezvous_two_processes(tmp_path):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_rdzv_worker, args=(r, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=60) for _ in ps)
    [p.join(30) for p in ps]
    assert got == [(0, True), (1, True)]


def test_parallel_shape_and_gather_order():
    par = parallelize.Parallel(10, environ={})
    assert (par.rank, par.size, par.working_set) == (0, 1, (0, 10))
    ans = par.answer_and_reduce({'x': np.arange(3)}, lambda parts: np.concatenate([p['x'] for p in parts]))
    assert ans.tolist() == [0, 1, 2]
    par2 = parallelize.Parallel(10, environ={'RANK': '1', 'WORLD_SIZE': '4', 'LOCAL_RANK': '1'})
    assert par2.working_set == (3, 3) and par2.devices == [1]


# ---- driver surface (monte_carlo3D.py:42-94, 1778-1843) --------------------------------------------------------
def test_drop_in_import_path_and_config(run_dir, monkeypatch):
    from monte_carloMPI import monte_carlo3D                      # reference monte_carlo3D-run.py:4
    import monte_carlompi_b200.monte_carlo3D as impl
    assert monte_carlo3D is impl
    mc = monte_carlo3D.MonteCarlo()
    assert (mc.tau_tot, mc.imp_cnc, mc.rho_snw, mc.rho_ice) == (1000000.0, 0.0, 300.0, 917.0)
    assert (mc.output_dir, mc.optics_dir, mc.fi_imp) == ('monte_carlo_results', 'inputdata', 'mie_sot_ChC90_dns_1317.nc')
    assert mc.HG is False and mc.phase_functions is False and mc.flg_3D == 999
    monkeypatch.setattr(sys, 'argv', ['x', '--tau_tot', '10', '--optics_dir', 'elsewhere'])
    mc = monte_carlo3D.MonteCarlo(rho_snw=200.)                   # kwargs beat flags beat config.ini
    assert (mc.tau_tot, mc.optics_dir, mc.rho_snw) == (10.0, 'elsewhere', 200.)
    import inspect
    sig = inspect.signature(monte_carlo3D.MonteCarlo.run)
    ref_args = ['self', 'n_photon', 'wvl0', 'half_width', 'rds_snw', 'theta_0', 'stokes_params', 'shape', 'roughness',
                'test', 'debug', 'Lambertian_surface', 'Lambertian_bottom', 'Lambertian_reflectance']
    assert list(sig.parameters)[:len(ref_args)] == ref_args      # monte_carlo3D.py:1492-1496
    assert sig.parameters['theta_0'].default == 0. and sig.parameters['Lambertian_bottom'].default is True
    assert sig.parameters['Lambertian_reflectance'].default == 1.


def test_out_of_scope_modes_raise(run_dir):
    from monte_carloMPI import monte_carlo3D
    mc = monte_carlo3D.MonteCarlo()
    with pytest.raises(NotImplementedError):                       # aspherical without --HG = full phase matrix
        mc.run(10, 1.3, 0.085, 100., shape='droxtal')
    with pytest.raises(NotImplementedError):
        mc.run(10, 1.3, 0.085, 100., debug=True)


def test_test_hook_presets_survive_run_attribute_overwrite(run_dir):
    from monte_carloMPI import monte_carlo3D
    mc = monte_carlo3D.MonteCarlo(tau_tot=2.0)
    mc.ssa_ice = 0.9
    mc.g = 0.75
    mc.g = np.array([0.1, 0.2])            # what run() does with the per-wavelength arrays
    assert mc._test_overrides() == {'ssa_ice': 0.9, 'g': 0.75}


# ---- BRF / albedo from tallies == the reference's per-photon formulas (post_processing.py:73-81) ---------------
def test_brf_and_albedo_from_tallies_match_per_photon_formulas():
    import gpu_util_cpu as gc
    from monte_carlompi_b200 import engine, post
    assert post.calculate_bins() == 137 and post.calculate_bins(2., 175.) == 68
    rng = np.random.RandomState(5)
    n, n_rows, nb = 200000, 7, 137
    table = np.zeros(n_rows, engine.ROW_DTYPE)
    table['wvl_um'] = 1.27 + 0.01 * np.arange(n_rows)
    rec = dict(wvl_row=rng.randint(0, n_rows, n).astype(np.int16), condition=rng.choice([1, 2, 3, 4, 5], n, p=[.4, .1, .05, .4, .05]).astype(np.uint8),
               theta_n=np.arccos(np.sqrt(rng.uniform(0, 1, n))).astype(np.float32))
    tally = gc.tally_from_records(rec, n_rows, nb)
    wvn = (1. / table['wvl_um'])[rec['wvl_row']]
    mid_t, brf_t = post.brf_from_tally(tally, table)
    mid_r, brf_r = post.brf_from_records(rec['condition'], wvn, rec['theta_n'], nb)
    assert np.array_equal(mid_t, mid_r) and np.allclose(brf_t, brf_r, rtol=1e-12, atol=0)
    q_up = wvn[rec['condition'] == 1].sum() / wvn.sum()
    assert abs(post.albedo_from_tally(tally, table) - q_up) < 1e-13
    fr = post.outcome_fractions(tally)
    assert abs(fr[1] - (rec['condition'] == 1).mean()) < 1e-15 and abs(sum(fr.values()) - 1) < 1e-12
    # a Lambertian reflector has BRF == albedo in every bin (cos-law sampling above): check the normalisation
    assert abs(np.median(brf_t) - q_up) < 0.02


def test_packed_records_unpack_on_the_host():
    # include/mc3d.h: word 0 = n_scat << 9 | row, words 1..3 = float bits of theta, phi, path with the three condition
    # bits in their sign bits; mc3d_unpack_records is host code (no GPU needed)
    from monte_carlompi_b200 import engine
    rng = np.random.RandomState(5)
    n = 200000
    cond = rng.randint(1, 6, n).astype(np.uint8)
    row = rng.randint(0, 512, n).astype(np.int16)
    n_scat = rng.randint(0, engine.PACKED_NSCAT_MAX + 1, n).astype(np.uint32)
    theta = rng.uniform(0, np.pi, n).astype(np.float32)
    phi = rng.uniform(0, 2 * np.pi, n).astype(np.float32)
    phi[::7] = 0.0
    path = rng.exponential(0.01, n).astype(np.float32)
    packed = np.empty((n, 4), np.uint32)
    packed[:, 0] = (n_scat << 9) | row.astype(np.uint32)
    packed[:, 1] = theta.view(np.uint32) | ((cond.astype(np.uint32) & 1) << 31)
    packed[:, 2] = phi.view(np.uint32) | (((cond.astype(np.uint32) >> 1) & 1) << 31)
    packed[:, 3] = path.view(np.uint32) | (((cond.astype(np.uint32) >> 2) & 1) << 31)
    for threads in (1, 0):
        out = engine.unpack_records(packed, n, n_threads=threads)
        for name, want in (('condition', cond), ('wvl_row', row), ('n_scat', n_scat), ('theta_n', theta), ('phi_n', phi),
                           ('path_length', path)):
            assert out[name].dtype == want.dtype and np.array_equal(out[name], want), name
    assert engine.unpack_records(np.zeros(0, np.uint32))['condition'].size == 0


def test_run_sweep_rejects_too_many_cases(run_dir, monkeypatch):
    # the case index lives in the high bits of the photon id and the engine numbers at most 1024 cases per call
    from monte_carloMPI import monte_carlo3D
    from monte_carlompi_b200 import engine
    mc = monte_carlo3D.MonteCarlo()
    cases = [dict(n_photon=10, wvl0=1.3, half_width=0.085, rds_snw=100.)] * (engine.SWEEP_MAX_CASES + 1)
    with pytest.raises(ValueError, match='at most'):
        mc.run_sweep(cases)
    assert mc.run_sweep([]) == []
