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
