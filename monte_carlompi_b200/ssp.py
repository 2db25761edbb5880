"""Per-wavelength single-scattering properties (SSP) for the walk: host-side input preparation.

Restates, per *distinct* rounded wavelength instead of per photon, what the reference does at
monte_carloMPI/monte_carlo3D.py:498-658 (get_optical_properties), 661-777 (get_impurity_optics) and 1575-1588, 1612
(extinction mix, impurity probability, snow depth).  The result is the table uploaded to the GPU
(``mc3d_ssp_row`` in include/mc3d.h): row r holds wavelength (k_first + r) / 100 um.

Lookup rule of the reference, reproduced exactly:
  * the two table rows nearest to the wavelength are selected (argsort of |wvl - wvl_in|, :2);
  * if the wavelength lies inside their bracket: linear interpolation between the two (np.interp, which is what
    scipy.interpolate.interp1d(kind='linear') dispatches to for float64 data)
  * otherwise interp1d raises ValueError and the reference falls back to the nearest row's value, printing
    'error: exception raised while interpolating <name>, using nearest value instead' on stderr.
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file

from .engine import ROW_DTYPE


def read_table(path, names):
    """Read 1-D variables ``names`` from a NetCDF-3 file (monte_carlo3D.py:509-514, 667-671)."""
    f = netcdf_file(path, 'r', mmap=False)
    try:
        return {n: np.array(f.variables[n].data) for n in names}
    finally:
        f.close()


def _interp_two_points(x, y, x_new):
    """What scipy.interpolate.interp1d(x, y)(x_new) returns for two points with x[0] < x[1], same IEEE operations:
    native-endian float64 data go through np.interp; anything else (e.g. the big-endian arrays scipy's NetCDF
    reader hands out, which is the case for the reference) through de Boor's two-term form
    (scipy/interpolate/_interpolate.py, interp1d._call_linear_np / _call_linear)."""
    native = (np.dtype(np.float64), np.dtype(int))
    if x.dtype in native and y.dtype in native:
        return np.interp(x_new, x, y)
    return ((x_new - x[0]) / (x[1] - x[0])) * y[1] + ((x[1] - x_new) / (x[1] - x[0])) * y[0]


def nearest_pair_interp(wvl_in, columns, wvls_um, warn_names=None):
    """Evaluate ``columns`` (dict name -> array over the table) at each wavelength of ``wvls_um`` [um].

    ``wvl_in`` is the table's wavelength axis in metres.  Returns dict name -> float64 array."""
    wvls_um = np.atleast_1d(np.asarray(wvls_um, dtype=np.float64))
    out = {n: np.empty(wvls_um.shape, dtype=np.float64) for n in columns}
    for j, w_um in enumerate(wvls_um):
        wvl = w_um * 1e-6                                           # monte_carlo3D.py:519, 594
        idx = np.argsort(np.absolute(wvl - wvl_in))[:2]             # monte_carlo3D.py:522, 596
        x = wvl_in[idx]
        if not x[0] < x[1]:                                         # monte_carlo3D.py:528-535
            idx = idx[::-1]
            x = x[::-1]
        inside = bool(x[0] <= wvl <= x[1]) and bool(x[0] < x[1])    # interp1d(bounds_error=True)
        for n, col in columns.items():
            y = col[idx]
            if inside:
                out[n][j] = _interp_two_points(x, y, wvl)
            else:
                # ValueError path: value of the nearest row (idx_wvl[0] in the reference's ordering)
                nearest = np.argsort(np.absolute(wvl - wvl_in))[0]
                out[n][j] = col[nearest]
                if warn_names is not None:
                    sys.stderr.write('error: exception raised while interpolating %s, using nearest value '
                                     'instead\n' % warn_names.get(n, n))
    return out


def ice_file(optics_dir, rds_snw):
    return os.path.join(optics_dir, 'mie', 'snicar', 'ice_wrn_%04d.nc' % rds_snw)   # monte_carlo3D.py:507-508


def impurity_file(optics_dir, fi_imp):
    return os.path.join(optics_dir, 'mie', 'snicar', fi_imp)                        # monte_carlo3D.py:666


# run(shape=...) / run(roughness=...) -> directory names of the aspherical library (monte_carlo3D.py:185-210)
SHAPE_DIRS = {'solid hexagonal column': 'solid_column', 'hexagonal plate': 'plate',
              'hollow hexagonal column': 'hollow_column', 'droxtal': 'droxtal',
              'hollow bullet rosette': 'hollow_bullet_rosette', 'solid bullet rosette': 'solid_bullet_rosette',
              '8-element column aggregate': 'column_8elements', '5-element plate aggregate': 'plate_5elements',
              '10-element plate aggregate': 'plate_10elements'}
ROUGHNESS_DIRS = {'smooth': 'Rough000', 'moderately rough': 'Rough003', 'severely rough': 'Rough050'}


def aspherical_dirs(shape, roughness, wvl0):
    """(band, shape_dir, roughness_dir) of the aspherical library for a run (monte_carlo3D.py:180-210, 1529-1545)."""
    if 0.2 <= wvl0 <= 15.25:
        band = '0.2-15.25'
    elif 16.4 <= wvl0 <= 99.0:
        band = '16.4-99.0'
    else:
        raise ValueError('wvl0 = %r um is outside the aspherical library (0.2-15.25 and 16.4-99.0 um)' % (wvl0,))
    if shape not in SHAPE_DIRS:
        raise ValueError('unknown shape %r; one of %s' % (shape, sorted(SHAPE_DIRS)))
    if roughness not in ROUGHNESS_DIRS:
        raise ValueError('unknown roughness %r; one of %s' % (roughness, sorted(ROUGHNESS_DIRS)))
    return band, SHAPE_DIRS[shape], ROUGHNESS_DIRS[roughness]


def read_isca(path):
    """Columns of ``isca.dat`` (monte_carlo3D.py:216-246): wvl [um], max_dim, volume, G, Q_ext, ssa, asm."""
    cols = [[] for _ in range(7)]
    with open(path, 'r') as f:
        for line in f:
            parts = line.split()
            for j in range(7):
                cols[j].append(float(parts[j]))
    return [np.array(c) for c in cols]


def build_table_aspherical(optics_dir, fi_imp, shape, roughness, wvl0, rds_snw, k_lo, k_hi, imp_cnc, rho_ice,
                           overrides=None, quiet=False):
    """SSP rows for aspherical grains under ``--HG`` (get_aspherical_SSPs, monte_carlo3D.py:173-266, 316-336,
    396-415): the size class whose effective radius RE = 3 V / (4 G) is nearest to ``rds_snw`` at wavelength
    ``wvl0`` (which must be a wavelength of the library), and for each photon wavelength the NEAREST library
    wavelength -- no interpolation -- which also replaces the photon's wavelength (monte_carlo3D.py:398-400, 1536),
    so ``rows['wvl_um']`` holds library wavelengths.  Returns (rows, snow_effective_radius)."""
    overrides = overrides or {}
    band, shape_dir, roughness_dir = aspherical_dirs(shape, roughness, wvl0)
    path = os.path.join(optics_dir, 'ice_optics', band, shape_dir, roughness_dir, 'isca.dat')
    wvl_in, _, volume_in, G_in, Q_ext_in, ssa_in, asm_in = read_isca(path)
    wvl0_idxs = np.where(wvl_in == wvl0)                            # monte_carlo3D.py:249
    if len(wvl0_idxs[0]) == 0:
        raise ValueError('wvl0 = %r um is not a wavelength of %s (the reference requires an exact member)'
                         % (wvl0, path))
    RE = (3. / 4.) * (volume_in / G_in)                             # monte_carlo3D.py:252
    idx_RE = np.argsort(np.absolute(rds_snw - RE[wvl0_idxs]))[0]
    snow_effective_radius = RE[wvl0_idxs][idx_RE]
    valid = np.where(RE == snow_effective_radius)
    wvl_in, volume_in, G_in, Q_ext_in, ssa_in, asm_in = (a[valid] for a in (wvl_in, volume_in, G_in, Q_ext_in,
                                                                           ssa_in, asm_in))
    k = np.arange(k_lo, k_hi + 1)
    wvls = k / 100.0
    lib_wvl = np.empty(len(k))
    vals = {n: np.empty(len(k)) for n in ('ssa_ice', 'ext_cff_mss_ice', 'g')}
    for j, wvl in enumerate(wvls):
        i0 = np.argsort(np.absolute(wvl - wvl_in))[0]              # monte_carlo3D.py:331-332
        lib_wvl[j] = wvl_in[i0]
        vals['ssa_ice'][j] = ssa_in[i0]
        vals['ext_cff_mss_ice'][j] = (1e6 * G_in[i0] * Q_ext_in[i0]) / (rho_ice * volume_in[i0])   # :409-412
        vals['g'][j] = asm_in[i0]
    imp = read_table(impurity_file(optics_dir, fi_imp), ('wvl', 'ss_alb', 'ext_cff_mss'))
    imp_v = nearest_pair_interp(imp['wvl'], {'ssa_imp': imp['ss_alb'], 'ext_cff_mss_imp': imp['ext_cff_mss']},
                                lib_wvl, None if quiet else {})
    vals.update(imp_v)
    for name in ('ssa_ice', 'ext_cff_mss_ice', 'g', 'ssa_imp', 'ext_cff_mss_imp'):
        if name in overrides and overrides[name] is not None:
            vals[name] = overrides[name] * np.ones(len(k))
    return derive_rows(lib_wvl, vals, imp_cnc), snow_effective_radius


def wavelength_grid(wvl0, sigma, n_sigma=7.0):
    """Integer grid k (wavelength = k / 100 um) covering wvl0 +- n_sigma sigma.

    The device draws Box-Muller normals from 32-bit uniforms, |z| <= sqrt(-2 ln 2^-33) = 6.76, so 7 sigma covers
    every wavelength np.around(normal(wvl0, sigma), 2) (monte_carlo3D.py:1519-1520) can produce here."""
    k_lo = int(np.floor((wvl0 - n_sigma * sigma) * 100.0)) - 1
    k_hi = int(np.ceil((wvl0 + n_sigma * sigma) * 100.0)) + 1
    k_lo = max(k_lo, 1)
    k_hi = max(k_hi, k_lo)
    return k_lo, k_hi


def build_table(optics_dir, fi_imp, rds_snw, k_lo, k_hi, imp_cnc, overrides=None, quiet=False):
    """SSP rows for wavelengths k / 100 um, k = k_lo .. k_hi.

    ``overrides``: the reference's ``test=True`` hook (monte_carlo3D.py:1553-1573) -- constants replacing
    ssa_ice / ext_cff_mss_ice / g / ssa_imp / ext_cff_mss_imp when present."""
    overrides = overrides or {}
    k = np.arange(k_lo, k_hi + 1)
    wvls = k / 100.0   # == np.around(x, 2) for every x that rounds to k: rint(100 x) / 100
    warn = None if quiet else {}
    ice = read_table(ice_file(optics_dir, rds_snw), ('wvl', 'ss_alb', 'ext_cff_mss', 'asm_prm'))
    ice_v = nearest_pair_interp(ice['wvl'], {'ssa_ice': ice['ss_alb'], 'ext_cff_mss_ice': ice['ext_cff_mss'],
                                             'g': ice['asm_prm']}, wvls, warn)
    imp = read_table(impurity_file(optics_dir, fi_imp), ('wvl', 'ss_alb', 'ext_cff_mss'))
    imp_v = nearest_pair_interp(imp['wvl'], {'ssa_imp': imp['ss_alb'], 'ext_cff_mss_imp': imp['ext_cff_mss']},
                                wvls, warn)
    vals = dict(ice_v)
    vals.update(imp_v)
    for name in ('ssa_ice', 'ext_cff_mss_ice', 'g', 'ssa_imp', 'ext_cff_mss_imp'):
        if name in overrides and overrides[name] is not None:
            vals[name] = overrides[name] * np.ones(len(wvls))       # monte_carlo3D.py:1554-1573
    return derive_rows(wvls, vals, imp_cnc)


def derive_rows(wvls, vals, imp_cnc):
    """monte_carlo3D.py:1575-1588: combined extinction and impurity-extinction probability."""
    ext_ice, ext_imp = vals['ext_cff_mss_ice'], vals['ext_cff_mss_imp']
    rows = np.zeros(len(wvls), dtype=ROW_DTYPE)
    rows['wvl_um'] = wvls
    rows['ssa_ice'] = vals['ssa_ice']
    rows['ssa_imp'] = vals['ssa_imp']
    rows['g'] = vals['g']
    rows['ext_cff_mss'] = ext_ice * (1 - imp_cnc) + ext_imp * imp_cnc
    rows['p_ext_imp'] = (imp_cnc * ext_imp) / (imp_cnc * ext_imp + (1 - imp_cnc) * ext_ice)
    return rows


def snow_depth(rows, tau_tot, rho_snw):
    """monte_carlo3D.py:1612."""
    return tau_tot / (rows['ext_cff_mss'] * rho_snw)
