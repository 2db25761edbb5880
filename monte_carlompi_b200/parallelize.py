"""Work partition over GPUs -- the B200 counterpart of the reference's monte_carloMPI/parallelize.py.

The reference's ``Parallel`` scatters ``np.array_split(wvls, size)`` over MPI ranks and gathers the per-photon
answers to rank 0 (parallelize.py:14-38).  Here the unit of work is a photon-id range and the ranks are GPUs:

  * default (plain ``python monte_carlo3D-run.py``): ONE process drives every visible GPU; libmc3d splits the id
    range with the same ``array_split`` boundaries and sums the tallies with one ncclReduce;
  * one process per GPU (``torchrun`` / ``mpirun -np N``, detected from RANK / WORLD_SIZE or the OpenMPI
    variables): each rank walks its own sub-range on GPU ``LOCAL_RANK``; rank 0 creates the NCCL id and passes it
    through a file in ``MC3D_RDZV_DIR`` (default: the system temp dir; all ranks are on one node); records travel to
    rank 0 the same way, tallies through ``mc3d_reduce_tally``.

Photon results depend only on (seed, photon id), so every layout gives bit-identical output.
"""
import os
import tempfile
import time

import numpy as np


def partition(n, parts):
    """(begin, count) of each of ``parts`` chunks of ``range(n)`` with np.array_split boundaries
    (parallelize.py:14-15): the first ``n % parts`` chunks hold one extra element."""
    q, r = divmod(int(n), int(parts))
    out = []
    begin = 0
    for k in range(parts):
        cnt = q + (1 if k < r else 0)
        out.append((begin, cnt))
        begin += cnt
    return out


def detect_ranks(environ=None):
    """(rank, world_size, local_rank) from torchrun- or OpenMPI-style environment variables; (0, 1, 0) otherwise."""
    env = os.environ if environ is None else environ
    for r, w, l in (('RANK', 'WORLD_SIZE', 'LOCAL_RANK'),
                    ('OMPI_COMM_WORLD_RANK', 'OMPI_COMM_WORLD_SIZE', 'OMPI_COMM_WORLD_LOCAL_RANK'),
                    ('PMI_RANK', 'PMI_SIZE', 'MPI_LOCALRANKID')):
        if r in env and w in env:
            return int(env[r]), int(env[w]), int(env.get(l, env[r]))
    return 0, 1, 0


class FileRendezvous(object):
    """Tiny single-node exchange: rank 0 publishes byte blobs under a job-unique directory, others poll for them."""

    def __init__(self, rank, world_size, token=None, root=None, timeout=300.0):
        self.rank, self.world_size, self.timeout = rank, world_size, timeout
        env = os.environ
        token = token or '_'.join(str(x) for x in (env.get('TORCHELASTIC_RUN_ID', 'job'),
                                                    env.get('MASTER_PORT', env.get('OMPI_MCA_orte_hnp_uri', '0')),
                                                    os.getppid()))
        token = ''.join(c if c.isalnum() else '_' for c in token)[:96]
        self.dir = os.path.join(root or env.get('MC3D_RDZV_DIR', tempfile.gettempdir()), 'mc3d_rdzv_' + token)
        os.makedirs(self.dir, exist_ok=True)

    def put(self, name, blob):
        tmp = os.path.join(self.dir, '.%s.%d.tmp' % (name, os.getpid()))
        with open(tmp, 'wb') as f:
            f.write(blob)
        os.replace(tmp, os.path.join(self.dir, name))

    def get(self, name):
        path = os.path.join(self.dir, name)
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > self.timeout:
                raise TimeoutError('rendezvous: %s did not appear within %.0f s' % (path, self.timeout))
            time.sleep(0.01)
        with open(path, 'rb') as f:
            return f.read()

    def barrier(self, name):
        self.put('%s.%d' % (name, self.rank), b'1')
        for r in range(self.world_size):
            self.get('%s.%d' % (name, r))

    def cleanup(self):
        if self.rank == 0:
            for fn in os.listdir(self.dir):
                try:
                    os.remove(os.path.join(self.dir, fn))
                except OSError:
                    pass
            try:
                os.rmdir(self.dir)
            except OSError:
                pass


class Parallel(object):
    """Same shape as the reference's ``Parallel`` (``size``, ``rank``, ``working_set``, ``answer_and_reduce``) with
    photon-id ranges as the data set.

    ``data_set``: the number of photons (or any sized sequence, whose length is used).  ``working_set`` is this
    rank's ``(photon_begin, n_photon)``.
    """

    def __init__(self, data_set, devices=None, environ=None, engine_module=None):
        n = int(data_set) if np.isscalar(data_set) else len(data_set)
        self.n_total = n
        self.rank, self.size, self.local_rank = detect_ranks(environ)
        self._engine = engine_module
        self._rdzv = None
        self._call = 0
        if self.size > 1:
            self.working_set = partition(n, self.size)[self.rank]
            self.devices = [self.local_rank if devices is None else devices[self.local_rank]]
        else:
            self.working_set = (0, n)
            self.devices = devices
        self.context = None

    def _map(self, n):
        """Re-partition for a new photon count (reference: Parallel._map, parallelize.py:28-38)."""
        self.n_total = int(n)
        self.working_set = partition(n, self.size)[self.rank] if self.size > 1 else (0, int(n))
        return self.working_set

    # ---- context creation ------------------------------------------------------------------------------------
    def open(self):
        """Create the libmc3d context for this layout (lazily; needs a GPU)."""
        if self.context is not None:
            return self.context
        eng = self._engine
        if eng is None:
            from . import engine as eng
        if self.size > 1:
            self._rdzv = FileRendezvous(self.rank, self.size)
            if self.rank == 0:
                self._rdzv.put('nccl_id', eng.nccl_unique_id())
            nccl_id = self._rdzv.get('nccl_id')
            self.context = eng.Context(rank=self.rank, world_size=self.size, nccl_id=nccl_id,
                                       device=self.devices[0])
        else:
            devices = self.devices
            if devices is None:
                devices = list(range(max(1, eng.device_count())))
            self.devices = devices
            self.context = eng.Context(devices=devices)
        return self.context

    def close(self):
        if self.context is not None:
            self.context.close()
            self.context = None
        if self._rdzv is not None:
            # every rank checks out; only rank 0 waits (for all of them) and then removes the directory -- nobody
            # waits on rank 0, so it can never delete a file another rank is still polling for
            self._rdzv.put('close.%d' % self.rank, b'1')
            if self.rank == 0:
                for r in range(1, self.size):
                    self._rdzv.get('close.%d' % r)
                self._rdzv.cleanup()
            self._rdzv = None

    # ---- gather ----------------------------------------------------------------------------------------------
    def allgather_bytes(self, blob):
        """Every rank's ``blob`` on every rank, in rank order (small control data, e.g. histogram extrema)."""
        if self.size == 1:
            return [blob]
        self._call += 1
        tag = 'allgather%d' % self._call
        self._rdzv.put('%s.%d' % (tag, self.rank), blob)
        return [self._rdzv.get('%s.%d' % (tag, r)) for r in range(self.size)]

    def answer_and_reduce(self, answer, reduce_work_fn):
        """Reference semantics (parallelize.py:17-26, 40-41): gather every rank's ``answer`` to rank 0 and return
        ``reduce_work_fn(list_of_answers)`` there, ``None`` elsewhere.  ``answer`` is a dict of numpy columns."""
        if self.size == 1:
            return reduce_work_fn([answer])
        self._call += 1
        tag = 'answer%d' % self._call
        if self.rank != 0:
            import io
            buf = io.BytesIO()
            np.savez(buf, **answer)
            self._rdzv.put('%s.%d' % (tag, self.rank), buf.getvalue())
            return None
        import io
        answers = [answer]
        for r in range(1, self.size):
            z = np.load(io.BytesIO(self._rdzv.get('%s.%d' % (tag, r))))
            answers.append({k: z[k] for k in z.files})
        return reduce_work_fn(answers)
