"""Work partition over GPUs -- the B200 counterpart of the reference's monte_carloMPI/parallelize.py.

The reference's ``Parallel`` scatters ``np.array_split(wvls, size)`` over MPI ranks and gathers the per-photon
answers to rank 0 (parallelize.py:14-38).  Here the unit of work is a photon-id range and the ranks are GPUs:

  * default (plain ``python monte_carlo3D-run.py``): ONE process drives every visible GPU; libmc3d splits the id
    range with the same ``array_split`` boundaries and sums the tallies with one ncclReduce;
  * one process per GPU (``torchrun`` / ``mpirun -np N``, detected from RANK / WORLD_SIZE or the OpenMPI
    variables): each rank walks its own sub-range on GPU ``LOCAL_RANK``; rank 0 creates the NCCL id (and the seed,
    when none was given) and passes it through a file in ``MC3D_RDZV_DIR`` (default: the system temp dir; all ranks
    are on one node); records travel to rank 0 through POSIX shared memory, tallies through ``mc3d_reduce_tally``.

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
    """Tiny single-node exchange of small control blobs (NCCL id, seed, shared-memory segment names): a rank publishes
    byte blobs under a directory unique to (job, Parallel instance), the others poll for them.  Bulk data (records)
    never goes through files -- see ``Parallel.answer_and_reduce``."""

    def __init__(self, rank, world_size, token=None, root=None, timeout=300.0, instance=0):
        self.rank, self.world_size, self.timeout = rank, world_size, timeout
        env = os.environ
        token = token or '_'.join(str(x) for x in (env.get('TORCHELASTIC_RUN_ID', 'job'),
                                                    env.get('MASTER_PORT', env.get('OMPI_MCA_orte_hnp_uri', '0')),
                                                    os.getppid()))
        token = ''.join(c if c.isalnum() else '_' for c in token)[:96]
        self.token = '%s_i%d' % (token, instance)
        self.dir = os.path.join(root or env.get('MC3D_RDZV_DIR', tempfile.gettempdir()), 'mc3d_rdzv_' + self.token)
        os.makedirs(self.dir, exist_ok=True)
        self._mine = []

    def put(self, name, blob):
        tmp = os.path.join(self.dir, '.%s.%d.tmp' % (name, os.getpid()))
        with open(tmp, 'wb') as f:
            f.write(blob)
        os.replace(tmp, os.path.join(self.dir, name))
        self._mine.append(name)

    def get(self, name):
        path = os.path.join(self.dir, name)
        t0 = time.time()
        pause = 5e-5
        while not os.path.exists(path):
            if time.time() - t0 > self.timeout:
                raise TimeoutError('rendezvous: %s did not appear within %.0f s' % (path, self.timeout))
            time.sleep(pause)
            pause = min(2 * pause, 2e-3)
        with open(path, 'rb') as f:
            return f.read()

    def barrier(self, name):
        self.put('%s.%d' % (name, self.rank), b'1')
        for r in range(self.world_size):
            self.get('%s.%d' % (name, r))

    def cleanup(self):
        """Remove this rank's own files (never another rank's); the directory goes with its last file."""
        for fn in self._mine:
            try:
                os.remove(os.path.join(self.dir, fn))
            except OSError:
                pass
        self._mine = []
        try:
            os.rmdir(self.dir)
        except OSError:
            pass                                    # another rank still has files here; the last one out removes it


class Parallel(object):
    """Same shape as the reference's ``Parallel`` (``size``, ``rank``, ``working_set``, ``answer_and_reduce``) with
    photon-id ranges as the data set.

    ``data_set``: the number of photons (or any sized sequence, whose length is used).  ``working_set`` is this
    rank's ``(photon_begin, n_photon)``.
    """

    _instances = 0          # rendezvous directories are numbered per process; every rank constructs its Parallel
                            # objects in the same order, so instance k on one rank meets instance k on the others

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

    # ---- rendezvous / context creation ---------------------------------------------------------------------------
    def rendezvous(self, root=None, token=None, timeout=300.0):
        """This instance's control channel (created on first use; ``None`` for a single process)."""
        if self.size > 1 and self._rdzv is None:
            Parallel._instances += 1
            self._rdzv = FileRendezvous(self.rank, self.size, token=token, root=root, timeout=timeout,
                                        instance=Parallel._instances)
        return self._rdzv

    def open(self):
        """Create the libmc3d context for this layout (lazily; needs a GPU)."""
        if self.context is not None:
            return self.context
        eng = self._engine
        if eng is None:
            from . import engine as eng
        if self.size > 1:
            rdzv = self.rendezvous()
            if self.rank == 0:
                rdzv.put('nccl_id', eng.nccl_unique_id())
            nccl_id = rdzv.get('nccl_id')
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
        rdzv, self._rdzv = self._rdzv, None
        if rdzv is None:
            return
        # A real barrier first: nobody removes anything while another rank may still be polling for it.  Afterwards a
        # rank removes the files it wrote -- except its barrier file, which others may still be reading -- and leaves
        # a 'gone' note written after its own last read; rank 0 collects the notes and removes the rest.
        rdzv.barrier('close')
        if self.rank != 0:
            keep = 'close.%d' % self.rank
            rdzv._mine = [f for f in rdzv._mine if f != keep]
            rdzv.cleanup()
            rdzv.put('gone.%d' % self.rank, b'1')
            return
        for r in range(1, self.size):
            rdzv.get('gone.%d' % r)
            rdzv._mine += ['gone.%d' % r, 'close.%d' % r]
        rdzv.cleanup()

    # ---- small control data --------------------------------------------------------------------------------------
    def allgather_bytes(self, blob):
        """Every rank's ``blob`` on every rank, in rank order (small control data, e.g. histogram extrema)."""
        if self.size == 1:
            return [blob]
        self._call += 1
        tag = 'allgather%d' % self._call
        rdzv = self.rendezvous()
        rdzv.put('%s.%d' % (tag, self.rank), blob)
        return [rdzv.get('%s.%d' % (tag, r)) for r in range(self.size)]

    def broadcast_bytes(self, blob, root=0):
        """``blob`` of rank ``root`` on every rank."""
        if self.size == 1:
            return blob
        self._call += 1
        name = 'bcast%d' % self._call
        rdzv = self.rendezvous()
        if self.rank == root:
            rdzv.put(name, blob)
            return blob
        return rdzv.get(name)

    def broadcast_seed(self, seed):
        """Rank 0's seed on every rank: a run whose seed came from OS entropy walks ONE random stream whatever the
        layout, and rank 0's ``last_seed`` reproduces it."""
        return int.from_bytes(self.broadcast_bytes(int(seed).to_bytes(8, 'little')), 'little')

    # ---- gather --------------------------------------------------------------------------------------------------
    def answer_and_reduce(self, answer, reduce_work_fn):
        """Reference semantics (parallelize.py:17-26, 40-41): gather every rank's ``answer`` to rank 0 and return
        ``reduce_work_fn(list_of_answers)`` there, ``None`` elsewhere.  ``answer`` is a dict of numpy columns.

        The columns travel through POSIX shared memory (all ranks are on one node): a rank lays its columns out in a
        segment of its own and publishes the segment's name and layout; rank 0 maps it, hands ``reduce_work_fn``
        views of it (no deserialisation; the one copy is the concatenation ``reduce_work_fn`` does anyway), and
        acknowledges, after which the owner unlinks it."""
        if self.size == 1:
            return reduce_work_fn([answer])
        import json
        import mmap
        self._call += 1
        tag = 'answer%d' % self._call
        rdzv = self.rendezvous()
        if self.rank != 0:
            cols = {k: np.ascontiguousarray(v) for k, v in answer.items()}
            layout, at = [], 0
            for k, v in cols.items():
                at = -(-at // 64) * 64
                layout.append((k, v.dtype.str, list(v.shape), at))
                at += v.nbytes
            path = os.path.join(_shm_dir(rdzv.dir), 'mc3d_%s_%s.%d' % (rdzv.token, tag, self.rank))
            fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            try:
                os.ftruncate(fd, max(at, 1))
                with mmap.mmap(fd, max(at, 1)) as m:
                    for (k, _, _, off), v in zip(layout, cols.values()):
                        np.frombuffer(m, np.uint8, v.nbytes, off)[:] = v.reshape(-1).view(np.uint8)
                rdzv.put('%s.%d' % (tag, self.rank), json.dumps({'path': path, 'bytes': max(at, 1), 'layout': layout}).encode())
                rdzv.get('%s.ack%d' % (tag, self.rank))
            finally:
                os.close(fd)
                os.unlink(path)
            return None
        answers, maps = [answer], []
        try:
            for r in range(1, self.size):
                meta = json.loads(rdzv.get('%s.%d' % (tag, r)).decode())
                fd = os.open(meta['path'], os.O_RDONLY)
                try:
                    m = mmap.mmap(fd, meta['bytes'], prot=mmap.PROT_READ)
                finally:
                    os.close(fd)
                maps.append(m)
                answers.append({k: np.ndarray(tuple(shape), np.dtype(dt), m, off) for k, dt, shape, off in meta['layout']})
            out = reduce_work_fn(answers)
            if any(np.shares_memory(v, a) for v in _arrays_of(out) for part in answers[1:] for a in part.values()):
                out = _deep_copy_arrays(out)        # e.g. a reduce function that returns one of its inputs
        finally:
            del answers
            for r, m in enumerate(maps, 1):
                try:
                    m.close()
                except BufferError:                 # a view escaped; the mapping lives until it is collected
                    pass
                rdzv.put('%s.ack%d' % (tag, r), b'1')
        return out


def _shm_dir(fallback):
    """Where the record segments live: the node's shared-memory file system (what shm_open uses), else ``fallback``."""
    return '/dev/shm' if os.path.isdir('/dev/shm') and os.access('/dev/shm', os.W_OK) else fallback


def _arrays_of(x):
    if isinstance(x, np.ndarray):
        return [x]
    if isinstance(x, dict):
        return [a for v in x.values() for a in _arrays_of(v)]
    if isinstance(x, (list, tuple)):
        return [a for v in x for a in _arrays_of(v)]
    return []


def _deep_copy_arrays(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _deep_copy_arrays(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_deep_copy_arrays(v) for v in x)
    return x
