"""world_size-2 host logic on CPU (gloo): photon-id ranges per rank, one reduce of the integer tallies, records
concatenated in rank order.  The walk itself is stood in for by the oracle's production-mode restatement, which
draws the same Philox stream keyed on (seed, global photon id) -- so the layout-invariance the GPU path promises
(bit-identical results for any rank count) is checked end to end without a GPU."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist           # noqa: E402
import torch.multiprocessing as tmp        # noqa: E402

N_PHOTON = 20001
SEED = 4242


def _table():
    from oracle import oracle
    rows = np.zeros(5, oracle.ROW_DTYPE)
    for j in range(5):
        rows[j] = (1.28 + 0.01 * j, 0.97 + 0.004 * j, 0.3, 0.89, 16.4, 0.0)
    return rows


def _params():
    from oracle import oracle
    return oracle.make_params(np.pi * 15. / 180., 5.0, 300., 0.5, 1.3, 0.01, 128, lambert_bottom=True, n_theta_bins=30)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from monte_carlompi_b200 import parallelize
    from oracle import oracle
    dist.init_process_group('gloo', rank=rank, world_size=world)
    par = parallelize.Parallel(N_PHOTON)                   # picks rank / size up from the environment
    assert (par.rank, par.size) == (rank, world)
    begin, count = par.working_set
    o = oracle.philox(_params(), _table(), SEED, begin, count, n_threads=1)
    tally = torch.from_numpy(o['tally'].astype(np.int64))
    dist.reduce(tally, dst=0, op=dist.ReduceOp.SUM)        # the path's single collective
    gathered = [None] * world
    dist.gather_object({k: o[k] for k in ('condition', 'n_scat', 'theta_n')}, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        rec = {k: np.concatenate([g[k] for g in gathered]) for k in gathered[0]}
        np.savez(os.path.join(out_dir, 'result.npz'), tally=tally.numpy(), **rec)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    from oracle import oracle
    port = 29500 + (os.getpid() % 2000)
    tmp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(str(tmp_path / 'result.npz'))
    one = oracle.philox(_params(), _table(), SEED, 0, N_PHOTON, n_threads=1)
    assert np.array_equal(z['tally'], one['tally'].astype(np.int64))
    for k in ('condition', 'n_scat', 'theta_n'):
        assert np.array_equal(z[k], one[k]), k
    assert z['tally'][:, 0].sum() == N_PHOTON
