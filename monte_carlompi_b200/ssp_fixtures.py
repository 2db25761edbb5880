"""Synthetic Mie single-scattering-property (SSP) tables in the reference's on-disk format.

The reference reads NetCDF-3 files ``<optics_dir>/mie/snicar/ice_wrn_%04d.nc`` (variables ``wvl`` [m],
``ss_alb``, ``ext_cff_mss`` [m2/kg], ``asm_prm``; reference monte_carloMPI/monte_carlo3D.py:507-514) and
``<optics_dir>/mie/snicar/<fi_imp>`` (``wvl``, ``ss_alb``, ``ext_cff_mss``; monte_carlo3D.py:666-671).
The Bohren & Huffman tarball that ships those files is absent from the reference checkout
(/root/reference/.MISSING_LARGE_BLOBS), so benches and tests use the deterministic tables written here, on the
SNICAR wavelength grid 0.305 ... 4.995 um, step 0.01 um (470 rows).  If the real tarball is unpacked into
``inputdata/`` the drivers read it unchanged; nothing in this module is on the product path.

Families:
  * ``const-nir``  ssa 0.992, g 0.89, ext 16.4        (lambda independent, NIR-like walk, ~66 events/photon)
  * ``const-kat``  ssa 0.9,   g 0.75, ext 16.4        (van de Hulst / Wang 1995 known-answer case)
  * ``const-vis``  ssa 0.999989859099, g 0.89, ext 6.6 (visible, long-tailed walks; monte_carlo3D.py:1872-1878)
  * ``spectral``   smooth analytic lambda dependence scaled with the effective radius

``write_isca`` writes the text table the reference reads for aspherical grains with ``--HG``
(``<optics_dir>/ice_optics/<band>/<shape>/<roughness>/isca.dat``, seven columns per line: wavelength [um], maximum
dimension [um], volume [um3], projected area [um2], Q_ext, single-scattering albedo, asymmetry factor;
monte_carlo3D.py:216-237) -- the Yang et al. (2013) library it points to is not in the archive either.
"""
import os

import numpy as np
from scipy.io import netcdf_file

RHO_ICE = 917.0

# log-linear anchors of the ice co-albedo (1 - ss_alb) for r_eff = 100 um
_COALB_ANCHORS_UM = np.array([0.305, 0.5, 0.8, 1.0, 1.3, 1.5, 1.8, 2.0, 2.5, 3.0, 5.0])
_COALB_ANCHORS = np.array([3e-7, 1e-6, 1e-4, 7e-4, 8e-3, 1e-1, 5e-2, 3.5e-1, 2.5e-1, 4.7e-1, 4.7e-1])


def snicar_grid_m():
    """470-row SNICAR wavelength grid in metres (0.305 ... 4.995 um)."""
    return (np.arange(470) * 0.01 + 0.305) * 1e-6


def ice_table(kind, rds_snw_um):
    """Return (wvl[m], ss_alb, ext_cff_mss, asm_prm) float64 arrays for one grain radius."""
    wvl = snicar_grid_m()
    um = wvl * 1e6
    n = wvl.size
    if kind == 'const-nir':
        ssa, ext, g = np.full(n, 0.992), np.full(n, 16.4), np.full(n, 0.89)
    elif kind == 'const-kat':
        ssa, ext, g = np.full(n, 0.9), np.full(n, 16.4), np.full(n, 0.75)
    elif kind == 'const-vis':
        ssa, ext, g = np.full(n, 0.999989859099), np.full(n, 6.6), np.full(n, 0.89)
    elif kind == 'spectral':
        coalb = np.exp(np.interp(np.log(um), np.log(_COALB_ANCHORS_UM), np.log(_COALB_ANCHORS)))
        coalb = np.minimum(coalb * (float(rds_snw_um) / 100.0), 0.47)
        ssa = 1.0 - coalb
        # geometric-optics limit Q_ext = 2: ext = 3 Q_ext / (4 rho_ice r)
        ext = np.full(n, 3.0 * 2.0 / (4.0 * RHO_ICE * float(rds_snw_um) * 1e-6)) * (1.0 + 0.01 * np.sin(um))
        g = np.clip(0.89 + 0.02 * (um - 1.3), 0.85, 0.97)
    else:
        raise ValueError('unknown SSP fixture family %r' % (kind,))
    return wvl, ssa, ext, g


def impurity_table():
    """Black-carbon-like impurity, lambda independent (monte_carlo3D.py:1881-1884)."""
    wvl = snicar_grid_m()
    return wvl, np.full(wvl.size, 0.30), np.full(wvl.size, 12000.0)


def _write(path, columns):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    f = netcdf_file(path, 'w')
    n = len(next(iter(columns.values())))
    f.createDimension('wvl', n)
    for name, data in columns.items():
        v = f.createVariable(name, 'd', ('wvl',))
        v[:] = np.asarray(data, dtype=np.float64)
    f.close()


def write_optics_dir(optics_dir, kind='spectral', radii_um=(100,), fi_imp='mie_sot_ChC90_dns_1317.nc'):
    """Write ice tables for ``radii_um`` plus the impurity table under ``optics_dir``; return optics_dir."""
    snicar = os.path.join(optics_dir, 'mie', 'snicar')
    for r in radii_um:
        wvl, ssa, ext, g = ice_table(kind, r)
        _write(os.path.join(snicar, 'ice_wrn_%04d.nc' % r),
               {'wvl': wvl, 'ss_alb': ssa, 'ext_cff_mss': ext, 'asm_prm': g})
    wvl, ssa, ext = impurity_table()
    _write(os.path.join(snicar, fi_imp), {'wvl': wvl, 'ss_alb': ssa, 'ext_cff_mss': ext})
    return optics_dir


# (volume / D^3, projected area / D^2) of the synthetic aspherical habits, D = maximum dimension
_HABIT_GEOMETRY = {'droxtal': (0.30, 0.60), 'solid_column': (0.12, 0.38), 'plate': (0.05, 0.45),
                   'hollow_column': (0.09, 0.38), 'hollow_bullet_rosette': (0.04, 0.30),
                   'solid_bullet_rosette': (0.06, 0.30), 'column_8elements': (0.03, 0.22),
                   'plate_5elements': (0.02, 0.25), 'plate_10elements': (0.015, 0.24)}

ISCA_MAX_DIMS_UM = (2., 5., 10., 20., 50., 100., 200., 400., 700., 1000., 2000., 4000.)


def isca_wavelengths_um(far_ir=False):
    """Two-decimal wavelengths of the synthetic library (so that wvl0 values like 1.3 are table members)."""
    if far_ir:
        return np.round(np.arange(16.4, 99.01, 2.0), 2)
    return np.round(np.concatenate([np.arange(0.2, 3.0, 0.05), np.arange(3.0, 15.26, 0.25)]), 2)


def write_isca(optics_dir, shape_dir='droxtal', roughness_dir='Rough000', far_ir=False):
    """Write a synthetic ``isca.dat`` for one habit / roughness; returns its path."""
    cv, cg = _HABIT_GEOMETRY[shape_dir]
    rough = {'Rough000': 0.0, 'Rough003': 0.01, 'Rough050': 0.03}[roughness_dir]
    band = '16.4-99.0' if far_ir else '0.2-15.25'
    path = os.path.join(optics_dir, 'ice_optics', band, shape_dir, roughness_dir, 'isca.dat')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as f:
        for w in isca_wavelengths_um(far_ir):
            coalb100 = np.exp(np.interp(np.log(min(w, 5.0)), np.log(_COALB_ANCHORS_UM), np.log(_COALB_ANCHORS)))
            for d in ISCA_MAX_DIMS_UM:
                vol, area = cv * d ** 3, cg * d ** 2
                r_eff = 0.75 * vol / area
                ssa = 1.0 - min(coalb100 * r_eff / 100.0, 0.47)
                q_ext = 2.0 + 0.5 / (1.0 + r_eff / w)
                g = min(0.97, 0.74 + 0.02 * (w - 1.3) + 0.03 * np.log10(d) - rough)
                f.write('%8.2f %10.2f %14.6E %14.6E %12.6E %12.6E %12.6E\n' % (w, d, vol, area, q_ext, ssa, g))
    return path
