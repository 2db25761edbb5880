"""The C-ABI library loads without a GPU and exports exactly what include/mc3d.h declares; the ctypes mirrors of
its structs have the C layout.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from monte_carlompi_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mc3d.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mc3d_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = _declared()
    assert set(names) == set(engine.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.mc3d_abi_version() == engine.ABI_VERSION


def test_header_is_plain_c_and_struct_layouts_match(tmp_path):
    prog = tmp_path / 'layout.c'
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mc3d.h"\nint main(void){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mc3d_params), sizeof(mc3d_ssp_row),'
                    'sizeof(mc3d_records), sizeof(mc3d_records_f64), sizeof(mc3d_stats),'
                    'offsetof(mc3d_params, k_first), offsetof(mc3d_params, n_phi_bins), offsetof(mc3d_stats, kernel_ms),'
                    'sizeof(mc3d_hist_spec), offsetof(mc3d_hist_spec, path_scale), sizeof(mc3d_extrema),'
                    'offsetof(mc3d_extrema, path_min));'
                    'return 0;}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'), str(prog), '-o', str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(engine.Params), engine.ROW_DTYPE.itemsize, C.sizeof(engine.Records), C.sizeof(engine.RecordsF64),
            C.sizeof(engine.Stats), engine.Params.k_first.offset, engine.Params.n_phi_bins.offset,
            engine.Stats.kernel_ms.offset, C.sizeof(engine.HistSpec), engine.HistSpec.path_scale.offset,
            C.sizeof(engine.Extrema), engine.Extrema.path_min.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device():
    if engine.device_count() > 0:
        pytest.skip('a GPU is visible')
    with pytest.raises(engine.Mc3dError) as e:
        engine.Context([0])
    assert 'no CUDA device' in str(e.value)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, 'monte_carlompi_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'oracle' not in text.replace('oracle/mc3d_oracle.c, which is how production mode is checked', ''), fn


def test_records_layout_is_packed_and_aligned():
    # host code, no GPU: column offsets of the block that comes back with a single copy
    off, total = engine.records_layout(1000)
    assert off == [0, 1024, 3072, 7168, 11264, 15360] and total == 19456
    off0, total0 = engine.records_layout(0)
    assert off0 == [0] * 6 and total0 == 0
    for n in (1, 255, 256, 257, 10**6, 2**26):
        off, total = engine.records_layout(n)
        sizes = [n * s for s in (1, 2, 4, 4, 4, 4)]
        assert all(o % 256 == 0 for o in off) and total % 256 == 0
        assert all(off[c] + sizes[c] <= off[c + 1] for c in range(5)) and off[5] + sizes[5] <= total


def _build_c_example(tmp_path):
    exe = tmp_path / 'mc3d_example'
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'examples', 'mc3d_example.c'), '-L', os.path.join(ROOT, 'monte_carlompi_b200'),
                           '-l:libmc3d.so', '-Wl,-rpath,' + os.path.join(ROOT, 'monte_carlompi_b200'), '-lm', '-o', str(exe)])
    return str(exe)


def test_plain_c_program_links_and_fails_loudly_without_a_device(tmp_path):
    # the boundary is a C ABI: a C99 program binds it with nothing but include/mc3d.h
    engine.load_library()
    exe = _build_c_example(tmp_path)
    if engine.device_count() > 0:
        pytest.skip('a GPU is visible (the run itself is tests/test_abi.py::test_plain_c_program_known_answer)')
    r = subprocess.run([exe, '1000'], capture_output=True, text=True)
    assert r.returncode == 2 and 'no CUDA device' in r.stderr


@pytest.mark.gpu
def test_plain_c_program_known_answer(tmp_path):
    # van de Hulst / Wang et al. known answer (monte_carlo3D.py:1849-1866) through the C ABI from C
    r = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert 'albedo 0.09' in r.stdout
