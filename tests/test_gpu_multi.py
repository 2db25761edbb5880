"""Multi-GPU paths (need >= 2 visible devices; skipped otherwise): one process driving several GPUs with the NCCL
reduce inside mc3d_run, and one context per GPU (the torchrun layout) with mc3d_reduce_tally.  Results must be
bit-identical to a single-GPU run: photon results depend only on (seed, photon id), tallies are integers."""
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
