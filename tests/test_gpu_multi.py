"""Multi-GPU paths (need >= 2 visible devices; skipped otherwise): one process driving several GPUs with the NCCL
reduce inside mc3d_run, and one context per GPU (the torchrun layout) with mc3d_reduce_tally.  Results must be
bit-identical to a single-GPU run: photon results depend only on (seed, photon id), tallies are integers."""
import os
import threading

import numpy as np
import pytest

import gpu_util
from monte_carlompi_b200 import engine

pytestmark = pytest.mark.gpu

SIGMA13 = 0.085 / 2.355


def _need(n):
    if engine.device_count() < n:
        pytest.skip('needs %d GPUs' % n)


def _workload():
    rows = gpu_util.fixture_table('spectral', 100, 104, 156)
    P, _ = gpu_util.both_params(15., 6.0, 0.5, 1.3, SIGMA13, 104, True)
    return rows, P


@pytest.mark.parametrize('n_dev', [2, 4, 8])
def test_one_process_many_gpus_equals_one_gpu(n_dev):
    _need(n_dev)
    rows, P = _workload()
    n, seed = 1000003, 77                        # not divisible: exercises the array_split remainders
    ref, tref, sref = gpu_util.context().run(P, rows, seed, 10, n)
    with engine.Context(list(range(n_dev))) as ctx:
        rec, tally, st = ctx.run(P, rows, seed, 10, n)
    assert st['n_devices'] == n_dev and st['n_events'] == sref['n_events']
    for col in ref:
        assert np.array_equal(rec[col], ref[col]), col
    assert np.array_equal(tally, tref)           # reduced over the devices by ncclReduce(sum, uint64)


def test_one_context_per_gpu_with_nccl_reduce():
    _need(2)
    from monte_carlompi_b200.parallelize import partition
    rows, P = _workload()
    n, seed, world = 400001, 5, 2
    ref, tref, _ = gpu_util.context().run(P, rows, seed, 0, n)
    nccl_id = engine.nccl_unique_id()
    out = [None] * world

    def rank_main(rank):
        ctx = engine.Context(rank=rank, world_size=world, nccl_id=nccl_id, device=rank)
        begin, count = partition(n, world)[rank]
        rec, tally, _ = ctx.run(P, rows, seed, begin, count)
        ctx.reduce_tally(tally, root=0)
        out[rank] = (rec, tally)
        ctx.close()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join(300) for t in threads]
    assert all(o is not None for o in out)
    for col in ref:
        assert np.array_equal(np.concatenate([out[r][0][col] for r in range(world)]), ref[col]), col
    assert np.array_equal(out[0][1], tref)


def test_histograms_over_two_gpus_equal_one_gpu():
    _need(2)
    rows, P = _workload()
    n, seed = 300001, 9
    spec = dict(n_scat_bins=200, n_scat_range=(0., 400.), path_bins=1000, path_range=(0., 25.), path_scale=100.)
    one = gpu_util.context()
    one.set_histograms(**spec)
    try:
        one.run(P, rows, seed, 0, n, records=False, tally=False)
        want, ext = one.histograms(0), one.extrema(0)
    finally:
        one.set_histograms()
    with engine.Context([0, 1]) as ctx:
        ctx.set_histograms(**spec)
        ctx.run(P, rows, seed, 0, n, records=False, tally=False)
        got = ctx.histograms(0)
        assert ctx.extrema(0) == ext
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and want[0].sum() > 0


_RANK_SCRIPT = r'''
import os, sys, shutil, json
import numpy as np
root, work, optics = sys.argv[1:4]
sys.path.insert(0, root)
os.chdir(work)
sys.argv = ['monte_carlo3D-run.py']
from monte_carloMPI import monte_carlo3D
mc = monte_carlo3D.MonteCarlo(optics_dir=optics, output_dir=os.path.join(work, 'out'), seed=31, tau_tot=6.0)
mc.run(200001, 1.3, 0.085, 100., theta_0=15., Lambertian_reflectance=0.5)
h = mc.histograms(200001, 1.3, 0.085, 100., theta_0=15., Lambertian_reflectance=0.5, n_scat_bins=50, path_length_bins=80)
if h is not None:
    np.savez(os.path.join(work, 'hist.npz'), ns=h['n_scat'][0], pl=h['path_length_cm'][0], tally=mc.last_tally)
mc.close()
# a second instance in the same job (its own rendezvous), no seed given: rank 0's is shared with the other ranks;
# a sweep of three cases (two of them sharing their rows) over the ranks
mc2 = monte_carlo3D.MonteCarlo(optics_dir=optics, output_dir=os.path.join(work, 'out2'), tau_tot=6.0)
cases = [dict(n_photon=50001, wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=15., Lambertian_reflectance=0.5),
         dict(n_photon=30000, wvl0=1.55, half_width=0.130, rds_snw=250., theta_0=0., Lambertian_reflectance=0.5),
         dict(n_photon=20002, wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=60., Lambertian_reflectance=0.5)]
res = mc2.run_sweep(cases, write_output=False)
rank = int(os.environ.get('RANK', 0))
with open(os.path.join(work, 'seed.%d' % rank), 'w') as f:
    f.write(str(mc2.last_seed))
if res[0] is not None:
    np.savez(os.path.join(work, 'sweep.npz'), seed=np.uint64(mc2.last_seed),
             **{'n_scat_%d' % k: r[0]['n_scat'] for k, r in enumerate(res)}, **{'tally_%d' % k: r[1] for k, r in enumerate(res)})
mc2.close()
'''


def test_driver_one_process_per_gpu_equals_one_process(tmp_path, optics_root):
    # `torchrun --nproc-per-node 2 monte_carlo3D-run.py` layout: every rank runs the script, rank 0 writes the file
    # (reference: mpirun -np N, README.md:45); ranks find each other through the file rendezvous of parallelize.py
    _need(2)
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'rank_main.py'
    script.write_text(_RANK_SCRIPT)
    outs = {}
    for world in (1, 2):
        work = tmp_path / ('w%d' % world)
        work.mkdir()
        shutil.copy(os.path.join(root, 'config.ini'), str(work / 'config.ini'))
        procs = []
        for rank in range(world):
            env = dict(os.environ, MC3D_RDZV_DIR=str(work), MASTER_PORT=str(29000 + world))
            if world > 1:
                env.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
            else:
                env.update(CUDA_VISIBLE_DEVICES='0')
                for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
                    env.pop(k, None)
            procs.append(subprocess.Popen([sys.executable, str(script), root, str(work), optics_root['spectral']], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        res = [p.communicate(timeout=600) for p in procs]
        assert all(p.returncode == 0 for p in procs), res
        paths = [l for l in res[0][0].splitlines() if l.endswith('.txt')]
        assert len(paths) == 1 and all('.txt' not in r[0] for r in res[1:])        # only rank 0 writes / prints
        outs[world] = (open(paths[0]).read(), np.load(str(work / 'hist.npz')))
    assert outs[1][0] == outs[2][0]                                                # byte-identical output file
    for k in ('ns', 'pl', 'tally'):
        assert np.array_equal(outs[1][1][k], outs[2][1][k]), k
    # the second instance: both ranks walked rank 0's seed, and the two-rank sweep is what one process computes under it
    w2 = tmp_path / 'w2'
    assert open(str(w2 / 'seed.0')).read() == open(str(w2 / 'seed.1')).read()
    z = np.load(str(w2 / 'sweep.npz'))
    sys.path.insert(0, root)
    from monte_carloMPI import monte_carlo3D
    cwd, argv = os.getcwd(), sys.argv
    os.chdir(str(tmp_path / 'w1'))
    sys.argv = ['monte_carlo3D-run.py']                       # MonteCarlo() reads flags from the command line
    try:
        mc = monte_carlo3D.MonteCarlo(optics_dir=optics_root['spectral'], output_dir=str(tmp_path / 'o3'), tau_tot=6.0, devices=[0])
        cases = [dict(n_photon=50001, wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=15., Lambertian_reflectance=0.5),
                 dict(n_photon=30000, wvl0=1.55, half_width=0.130, rds_snw=250., theta_0=0., Lambertian_reflectance=0.5),
                 dict(n_photon=20002, wvl0=1.3, half_width=0.085, rds_snw=100., theta_0=60., Lambertian_reflectance=0.5)]
        res = mc.run_sweep(cases, write_output=False, seed=int(z['seed']))
        mc.close()
    finally:
        os.chdir(cwd)
        sys.argv = argv
    for k, r in enumerate(res):
        assert np.array_equal(z['n_scat_%d' % k], r[0]['n_scat']), k
        assert np.array_equal(z['tally_%d' % k], r[1]), k
