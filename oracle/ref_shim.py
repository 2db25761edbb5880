"""TEST INFRASTRUCTURE ONLY -- runs the *unmodified* reference (/root/reference) in this container.

The reference cannot be imported as shipped here: matplotlib and mpi4py are absent, Python 3.12 dropped
``configparser.SafeConfigParser`` and the Mie tarball is missing.  This module installs the
four stubs that make ``from monte_carloMPI import monte_carlo3D`` work against the read-only sources where they
lie, runs ``MonteCarlo.run`` with a seeded ``np.random`` and captures

  * the raw per-photon answer tuples (before text formatting, reference monte_carlo3D.py:1614-1618),
  * the per-photon SSP arrays the reference derived (monte_carlo3D.py:1575-1588, 1612),
  * every uniform the walk consumed, in consumption order, with per-photon offsets (for replay mode).

``oracle/make_golden.py`` uses it to write the committed fixtures under ``tests/golden/``; ``oracle/ref_timing.py``
uses it to time the reference's Python path as the CPU baseline of ``bench.py`` (on the GPU box from the staged,
git-ignored copy ``oracle/_ref/reference`` -- ``/root/reference`` does not exist there; GPU tests and ``smoke()`` never
read either).  Never imported by the product package.
"""
import configparser
import contextlib
import io
import os
import shutil
import sys
import tempfile
import types

import numpy as np

# /root/reference in the build container; on the GPU box (where it does not exist) the git-ignored staging copy that
# __graft_entry__.build() made under oracle/_ref/reference (three .py files + config.ini, never committed)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'reference')
REFERENCE_ROOT = os.environ.get('MC3D_REFERENCE_ROOT') or \
    ('/root/reference' if os.path.isfile('/root/reference/monte_carloMPI/monte_carlo3D.py') else STAGED_ROOT)


def stage_reference():
    """Copy the reference's path sources into oracle/_ref/reference (git-ignored, not gpurun-ignored) so that the CPU
    baseline leg of bench.py can time the unmodified reference on the GPU box.  No-op when /root/reference is absent."""
    src = '/root/reference'
    if not os.path.isfile(os.path.join(src, 'monte_carloMPI', 'monte_carlo3D.py')):
        return False
    os.makedirs(os.path.join(STAGED_ROOT, 'monte_carloMPI'), exist_ok=True)
    for rel in ('monte_carloMPI/__init__.py', 'monte_carloMPI/monte_carlo3D.py', 'monte_carloMPI/parallelize.py',
                'monte_carloMPI/config.ini'):
        shutil.copyfile(os.path.join(src, rel), os.path.join(STAGED_ROOT, rel))
    shutil.copyfile(os.path.join(src, 'monte_carloMPI', 'config.ini'), os.path.join(STAGED_ROOT, 'config.ini'))
    # the reference's user script, for the test that runs it UNMODIFIED on top of the B200 package
    shutil.copyfile(os.path.join(src, 'monte_carlo3D-run.py'), os.path.join(STAGED_ROOT, 'monte_carlo3D-run.py'))
    return True


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'monte_carloMPI', 'monte_carlo3D.py'))


class _FakeComm(object):
    """mpi4py.MPI.COMM_WORLD for a single rank (reference parallelize.py:8-10, 19, 36)."""
    size = 1
    rank = 0

    def scatter(self, chunks, root=0):
        return chunks[0]

    def gather(self, answer, root=0):
        return [answer]


def _install_stubs():
    if not hasattr(configparser, 'SafeConfigParser'):
        configparser.SafeConfigParser = configparser.ConfigParser
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['mpl_toolkits'].mplot3d = sys.modules['mpl_toolkits.mplot3d']
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    if 'mpi4py' not in sys.modules:
        mpi4py = types.ModuleType('mpi4py')
        mpi = types.ModuleType('mpi4py.MPI')
        mpi.COMM_WORLD = _FakeComm()
        mpi4py.MPI = mpi
        sys.modules['mpi4py'] = mpi4py
        sys.modules['mpi4py.MPI'] = mpi


def import_reference():
    """Import the reference's monte_carlo3D module from REFERENCE_ROOT (read-only, nothing copied)."""
    if not reference_available():
        raise RuntimeError('reference sources not found under %s' % REFERENCE_ROOT)
    _install_stubs()
    # the drop-in package in this repo has the same import name; make sure we get the reference's
    for name in [m for m in sys.modules if m == 'monte_carloMPI' or m.startswith('monte_carloMPI.')]:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from monte_carloMPI import monte_carlo3D  # noqa: the reference's own module
    finally:
        sys.path.remove(REFERENCE_ROOT)
    assert monte_carlo3D.__file__.startswith(REFERENCE_ROOT), monte_carlo3D.__file__
    ref_modules = {m: sys.modules.pop(m) for m in list(sys.modules)
                   if m == 'monte_carloMPI' or m.startswith('monte_carloMPI.')}
    del ref_modules
    return monte_carlo3D


class Recorder(object):
    """Wraps np.random.rand / uniform / normal (the only entry points the path uses: monte_carlo3D.py:915, 921,
    1014, 1020, 1023, 1036-1038, 1245-1246, 1422, 1453, 1519) and logs raw uniforms in consumption order."""

    def __init__(self):
        self.values = []
        self.kinds = []  # 0 = rand, 1 = uniform(0, pi/2)
        self.normal_calls = 0

    def install(self):
        self._orig = (np.random.rand, np.random.uniform, np.random.normal)
        orig_rand, orig_uniform, orig_normal = self._orig

        def rand(*shape):
            out = orig_rand(*shape)
            self.values.extend(np.atleast_1d(out).ravel().tolist())
            self.kinds.extend([0] * int(np.size(out)))
            return out

        def uniform(low=0.0, high=1.0, size=None):
            assert size is None and low == 0
            # legacy RandomState.uniform = low + (high - low) * random_sample(); log the raw sample
            out = orig_rand()
            self.values.append(float(out))
            self.kinds.append(1)
            return low + (high - low) * out

        def normal(*a, **k):
            self.normal_calls += 1
            return orig_normal(*a, **k)

        np.random.rand, np.random.uniform, np.random.normal = rand, uniform, normal

    def uninstall(self):
        np.random.rand, np.random.uniform, np.random.normal = self._orig


def run_reference(n_photon, wvl0, half_width, rds_snw, theta_0, seed, optics_dir, model_kwargs=None,
                  run_kwargs=None, preset=None, record=True, keep_text=False):
    """Run the reference's MonteCarlo.run single-rank with ``np.random.seed(seed)``.

    ``preset``: dict of attributes set on the instance before ``run(test=True)`` (monte_carlo3D.py:1553-1573).
    Returns a dict with raw answers, per-photon SSP arrays and (if ``record``) the consumed random stream.
    """
    mc3d = import_reference()
    model_kwargs = dict(model_kwargs or {})
    run_kwargs = dict(run_kwargs or {})
    work = tempfile.mkdtemp(prefix='mc3d_ref_')
    cwd = os.getcwd()
    argv = sys.argv
    captured = {}
    rec = Recorder()
    try:
        shutil.copy(os.path.join(REFERENCE_ROOT, 'config.ini'), os.path.join(work, 'config.ini'))
        os.chdir(work)
        sys.argv = ['monte_carlo3D-run.py']
        mc = mc3d.MonteCarlo(optics_dir=optics_dir, output_dir=os.path.join(work, 'out'), **model_kwargs)
        for k, v in (preset or {}).items():
            setattr(mc, k, v)
        if preset:
            run_kwargs['test'] = True

        starts = []
        orig_walk = mc.monte_carlo3D

        def walk(wvl):
            starts.append(len(rec.values))
            return orig_walk(wvl)
        mc.monte_carlo3D = walk

        orig_reduce = mc3d.Parallel.answer_and_reduce

        def reduce(self, answer, fn):
            captured['answers'] = answer
            captured['working_set'] = np.array(self.working_set)
            return orig_reduce(self, answer, fn) if keep_text else None
        mc3d.Parallel.answer_and_reduce = reduce

        np.random.seed(seed)
        if record:
            rec.install()
        try:
            with contextlib.redirect_stdout(io.StringIO()) as out, contextlib.redirect_stderr(io.StringIO()):
                mc.run(n_photon, wvl0, half_width, rds_snw, theta_0=theta_0, **run_kwargs)
        finally:
            if record:
                rec.uninstall()
            mc3d.Parallel.answer_and_reduce = orig_reduce
        text = None
        if keep_text:
            path = out.getvalue().strip().splitlines()[-1]
            with open(path) as f:
                text = (os.path.basename(path), f.read())
    finally:
        os.chdir(cwd)
        sys.argv = argv
        shutil.rmtree(work, ignore_errors=True)

    ans = captured['answers']
    n = len(ans)
    res = {
        'condition': np.array([a[0] for a in ans], dtype=np.int32),
        'wvn': np.array([a[1] for a in ans], dtype=np.float64),
        'theta_n': np.array([a[2] for a in ans], dtype=np.float64),
        'phi_n': np.array([a[3] for a in ans], dtype=np.float64),
        'n_scat': np.array([a[4] for a in ans], dtype=np.int64),
        'path_length': np.array([np.float64(a[5]) for a in ans], dtype=np.float64),
        'snow_depth': np.array([a[6] for a in ans], dtype=np.float64),
        'wvl': captured['working_set'].astype(np.float64),
        'ssa_ice': np.array(mc.ssa_ice, dtype=np.float64) * np.ones(n),
        'ssa_imp': np.array(mc.ssa_imp, dtype=np.float64) * np.ones(n),
        'g': np.array(mc.g, dtype=np.float64) * np.ones(n),
        'ext_cff_mss': np.array(mc.ext_cff_mss, dtype=np.float64) * np.ones(n),
        'P_ext_imp': np.array(mc.P_ext_imp, dtype=np.float64) * np.ones(n),
        'text': text,
    }
    if record:
        vals = np.array(rec.values, dtype=np.float64)
        kinds = np.array(rec.kinds, dtype=np.uint8)
        assert rec.normal_calls == 1
        assert starts[0] == 3 * n, (starts[0], n)  # initial_pdfs drew 3 per photon before the first walk
        offsets = np.array(starts + [len(vals)], dtype=np.int64) - 3 * n
        res['init_draws'] = vals[:3 * n].copy()
        res['stream'] = vals[3 * n:].copy()
        res['stream_kinds'] = kinds[3 * n:].copy()
        res['offsets'] = offsets
        # recorder self-check: every photon consumed 5 per scatter plus Lambert-bottom extras
        extras = np.diff(offsets) - 5 * res['n_scat']
        assert (extras >= 0).all()
        res['extras'] = extras
    return res
