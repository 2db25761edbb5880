"""Helpers shared by the tests: golden fixtures (tests/golden/*.npz, written by oracle/make_golden.py from the
unmodified reference) and regeneration of the reference's random stream from the frozen legacy RandomState."""
import ast
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = ('c1_default', 'slab_tau3_lb', 'slab_tau05_normal', 'slab_tau3_black', 'impurity', 'kat_vdh', 'vis_debug',
         'isotropic', 'edge_of_table', 'vis_long', 'lambert_surface', 'c1_full_10k', 'aspherical_hg')


def regenerate_stream(seed, n_photon, wvl0, half_width, n_walk_draws):
    """Reference draw order: normal(size=n) -> 3 uniforms per photon -> walk draws."""
    rs = np.random.RandomState(seed)
    wvls = np.around(rs.normal(loc=wvl0, scale=half_width / 2.355, size=(n_photon)), decimals=2)
    init = rs.random_sample(3 * n_photon)
    stream = rs.random_sample(int(n_walk_draws))
    return wvls, init, stream


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    cfg = ast.literal_eval(str(z['config']))
    rows = z['rows']
    k = z['wvl_k'].astype(np.int64)
    # aspherical habits: the drawn wavelength only selects the row; the row's wavelength is the library's nearest one
    aspherical = 'rows_k' in z.files
    rows_k = z['rows_k'].astype(np.int64) if aspherical else np.rint(rows['wvl_um'] * 100).astype(np.int64)
    idx = np.searchsorted(rows_k, k)
    wvls, init, stream = regenerate_stream(cfg['seed'], cfg['n_photon'], cfg['wvl0'], cfg['half_width'],
                                           z['offsets'][-1])
    assert np.array_equal(np.rint(wvls * 100).astype(np.int64), k), 'legacy RandomState stream changed'
    case = dict(cfg=cfg, rows=rows, offsets=z['offsets'].astype(np.int64), init_draws=init, stream=stream,
                wvl=wvls, golden={c: z[c] for c in ('condition', 'theta_n', 'phi_n', 'n_scat', 'path_length')})
    for col in ('ssa_ice', 'ssa_imp', 'g', 'ext_cff_mss', 'p_ext_imp'):
        case[col] = rows[col][idx]
    if aspherical:
        case['wvl'] = rows['wvl_um'][idx]
        case['rows_k'] = rows_k
        case['file_name'] = str(z['file_name'])
    case['golden']['wvn'] = 1.0 / case['wvl']
    case['golden']['snow_depth'] = cfg['tau_tot'] / (case['ext_cff_mss'] * cfg['rho_snw'])
    return case


def compare_replay(out, case, min_exact=0.9999, rtol=1e-9):
    """north_star bar for replay mode: condition and n_scat exact for >= 99.99 % of photons, angles and path
    length within 1e-9 relative on the photons whose discrete outcome matches.  Returns a dict of stats."""
    gold = case['golden']
    n = len(gold['condition'])
    same = (out['condition'] == gold['condition']) & (out['n_scat'] == gold['n_scat'])
    stats = {'n': n, 'exact_fraction': float(same.mean()), 'n_mismatch': int(out['n_mismatch'])}
    assert same.mean() >= min_exact, stats
    assert np.array_equal(out['consumed'][same], np.diff(case['offsets'])[same])
    for col in ('wvn', 'theta_n', 'phi_n', 'path_length', 'snow_depth'):
        a, b = out[col][same], np.asarray(gold[col], dtype=np.float64)[same]
        # denominators below 1e-9 (e.g. the ~1e-19 m rounding residue of a path that cancels exactly in
        # Lambertian_surface mode) are compared absolutely
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-9)
        rel[(a == b)] = 0.0
        stats['maxrel_' + col] = float(rel.max()) if len(rel) else 0.0
        assert (rel <= rtol).all(), (col, stats)
    return stats
