"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libmc3d_oracle.so (the fp64 CPU restatement).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package.  See the header of mc3d_oracle.c for the reference lines it follows and how it is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libmc3d_oracle.so')
_lib = None

N_COND = 8


class Params(C.Structure):
    _fields_ = [('theta0_rad', C.c_double), ('tau_tot', C.c_double), ('rho_snw', C.c_double),
                ('r_lambert', C.c_double), ('wvl0_um', C.c_double), ('sigma_um', C.c_double),
                ('k_first', C.c_int32), ('flags', C.c_uint32), ('n_theta_bins', C.c_int32),
                ('n_phi_bins', C.c_int32)]   # unused by the oracle: its tallies are zenith-only


ROW_DTYPE = np.dtype([('wvl_um', 'f8'), ('ssa_ice', 'f8'), ('ssa_imp', 'f8'), ('g', 'f8'),
                      ('ext_cff_mss', 'f8'), ('p_ext_imp', 'f8')])


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not os.path.isfile(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, 'mc3d_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, '-B'], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_henyey_greenstein2.restype = C.c_double
        _lib.oracle_henyey_greenstein2.argtypes = [C.c_double, C.c_double]
        _lib.oracle_histogram_bin.argtypes = [C.c_double, C.c_int, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(theta0_rad, tau_tot, rho_snw, r_lambert, wvl0_um=0.0, sigma_um=0.0, k_first=0,
                lambert_bottom=True, lambert_surface=False, n_theta_bins=0):
    return Params(float(theta0_rad), float(tau_tot), float(rho_snw), float(r_lambert), float(wvl0_um),
                  float(sigma_um), int(k_first), (1 if lambert_bottom else 0) | (2 if lambert_surface else 0),
                  int(n_theta_bins), 0)


def replay(params, wvl, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp, init_draws, offsets, stream):
    """fp64 walk over the reference's recorded stream.  Returns dict of per-photon arrays + n_mismatch."""
    n = len(wvl)
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    wvl, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp = map(f8, (wvl, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp))
    init_draws, stream = f8(init_draws), f8(stream)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    out = {'condition': np.zeros(n, np.int32), 'wvn': np.zeros(n), 'theta_n': np.zeros(n), 'phi_n': np.zeros(n),
           'n_scat': np.zeros(n, np.int64), 'path_length': np.zeros(n), 'snow_depth': np.zeros(n),
           'consumed': np.zeros(n, np.int64)}
    mism = lib().oracle_replay(C.byref(params), C.c_int64(n), _p(wvl), _p(ssa_ice), _p(ssa_imp), _p(g),
                               _p(ext_cff_mss), _p(p_ext_imp), _p(init_draws), _p(offsets), _p(stream),
                               _p(out['condition']), _p(out['wvn']), _p(out['theta_n']), _p(out['phi_n']),
                               _p(out['n_scat']), _p(out['path_length']), _p(out['snow_depth']),
                               _p(out['consumed']))
    out['n_mismatch'] = int(mism)
    return out


def theta_edges(n_theta_bins):
    return np.linspace(0.0, np.pi / 2, n_theta_bins + 1)


def philox(params, table, seed, begin, n, n_threads=1, fp32_angles=True, records=True, tally=True):
    """Production-mode restatement: same Philox draws as the CUDA kernel, reference arithmetic in fp64."""
    table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
    n_rows = len(table)
    out = {}
    if records:
        out = {'condition': np.zeros(n, np.int32), 'wvl_row': np.zeros(n, np.int16), 'theta_n': np.zeros(n),
               'phi_n': np.zeros(n), 'n_scat': np.zeros(n, np.int64), 'path_length': np.zeros(n)}
    t = np.zeros((n_rows, N_COND + params.n_theta_bins), np.uint64) if tally else None
    edges = theta_edges(params.n_theta_bins) if params.n_theta_bins > 0 else np.zeros(1)
    ev = C.c_uint64(0)
    rc = lib().oracle_philox(C.byref(params), _p(table), C.c_int(n_rows), C.c_uint64(seed), C.c_uint64(begin),
                             C.c_uint64(n), C.c_int(n_threads), C.c_int(1 if fp32_angles else 0), _p(edges),
                             _p(out.get('condition')), _p(out.get('wvl_row')), _p(out.get('theta_n')),
                             _p(out.get('phi_n')), _p(out.get('n_scat')), _p(out.get('path_length')), _p(t),
                             C.byref(ev))
    assert rc == 0
    out['tally'] = t
    out['n_events'] = int(ev.value)
    return out


def philox4x32(ctr, key, rounds=7):
    """Philox4x32-R block (the production stream uses R = 7; Random123's known answers exist for R = 7 and 10)."""
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32(C.c_int(int(rounds)), _p(ctr), _p(key), _p(out))
    return out


def henyey_greenstein2(g, r):
    return lib().oracle_henyey_greenstein2(float(g), float(r))


def histogram_bin(x, n_bins):
    edges = theta_edges(n_bins)
    return lib().oracle_histogram_bin(float(x), int(n_bins), _p(edges))
