import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box')


def _n_gpus():
    try:
        from monte_carlompi_b200 import engine
        return engine.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # GPU tests never fall back to anything: without a device they are reported as skipped, not passed
    if _n_gpus() > 0:
        return
    skip = pytest.mark.skip(reason='no CUDA device visible (GPU tests call libmc3d.so through the C ABI)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def optics_root(tmp_path_factory):
    """Synthetic Mie tables in the reference's directory layout, one optics_dir per fixture family."""
    from monte_carlompi_b200 import ssp_fixtures
    root = tmp_path_factory.mktemp('optics')
    out = {}
    for kind in ('spectral', 'const-kat', 'const-vis', 'const-nir'):
        out[kind] = ssp_fixtures.write_optics_dir(str(root / kind), kind, (50, 100, 250, 500, 1000))
    # aspherical habits (isca.dat libraries, used with --HG)
    ssp_fixtures.write_isca(out['spectral'], 'droxtal', 'Rough003')
    ssp_fixtures.write_isca(out['spectral'], 'solid_column', 'Rough000')
    return out


@pytest.fixture()
def run_dir(tmp_path, monkeypatch):
    """A working directory with config.ini (MonteCarlo() reads it from the cwd) and a clean argv."""
    import shutil
    shutil.copy(os.path.join(ROOT, 'config.ini'), str(tmp_path / 'config.ini'))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, 'argv', ['monte_carlo3D-run.py'])
    return tmp_path
