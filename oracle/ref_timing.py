"""TEST / BENCH INFRASTRUCTURE ONLY -- times the reference's own Python implementation of the path on host cores.

``mpirun -np K monte_carlo3D-run.py`` (reference README.md:45) is emulated by K independent single-rank processes,
each running the UNMODIFIED ``MonteCarlo.run`` (reference monte_carlo3D.py:1492-1657: wavelength draw, SSP lookup,
initial_pdfs, the photon loop 1613-1616, the rank-0 text file) under the import shims of ``oracle/ref_shim.py`` with its
own ``np.random`` seed -- the reference's only inter-rank traffic is one scatter before and one gather after the loop
(parallelize.py:19, 36), and real MPI ranks are independently OS-seeded because the reference never calls ``seed``.
No MPI exists in the image (no ``mpirun``, no ``mpi4py``).

The reference sources are read from ``/root/reference`` in the build container, or from the git-ignored staging copy
``oracle/_ref/reference`` that ``__graft_entry__.build()`` makes so that they travel to the GPU box.  Never imported by
the product package.
"""
import multiprocessing as mp
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

_state = {}


def _init_worker(cfg):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from monte_carlompi_b200 import ssp_fixtures
    optics = os.path.join(tempfile.mkdtemp(prefix='mc3d_reftime_optics_'), cfg['fixture'])
    ssp_fixtures.write_optics_dir(optics, cfg['fixture'], (cfg['rds_snw'],))
    _state['optics'] = optics
    _state['cfg'] = cfg
    from oracle import ref_shim
    ref_shim.import_reference()          # pay the import once per worker, outside the timed calls


def _work(args):
    seed, n = args
    from oracle import ref_shim
    cfg = _state['cfg']
    t0 = time.perf_counter()
    r = ref_shim.run_reference(n, cfg['wvl0'], cfg['half_width'], cfg['rds_snw'], cfg['theta_0'], seed=seed,
                               optics_dir=_state['optics'], model_kwargs=dict(tau_tot=cfg['tau_tot'], rho_snw=cfg['rho_snw']),
                               run_kwargs=dict(Lambertian_bottom=cfg['lambert_bottom'], Lambertian_reflectance=cfg['r_lambert']),
                               record=False, keep_text=True)
    dt = time.perf_counter() - t0
    return n, int(r['n_scat'].sum()) + n, dt


class ReferencePool(object):
    """K worker processes, each holding the imported reference module and a synthetic optics directory."""

    def __init__(self, cfg, n_procs=None):
        from oracle import ref_shim
        if not ref_shim.reference_available():
            raise RuntimeError('reference sources not found (looked in %s)' % ref_shim.REFERENCE_ROOT)
        self.n_procs = int(n_procs or os.cpu_count())
        self._pool = mp.get_context('spawn').Pool(self.n_procs, initializer=_init_worker, initargs=(dict(cfg),))
        self._next_seed = 9000
        self._pool.map(_work, [(1, 4)] * self.n_procs, chunksize=1)      # touch every worker (imports, file caches)

    def run(self, photons_per_proc):
        """One emulated ``mpirun -np K`` run of K x photons_per_proc photon packets.  Returns (photons, events, wall s,
        mean per-process run() seconds)."""
        jobs = [(self._next_seed + k, int(photons_per_proc)) for k in range(self.n_procs)]
        self._next_seed += self.n_procs
        t0 = time.perf_counter()
        out = self._pool.map(_work, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(o[0] for o in out), sum(o[1] for o in out), wall, sum(o[2] for o in out) / len(out)

    def close(self):
        self._pool.close()
        self._pool.join()
