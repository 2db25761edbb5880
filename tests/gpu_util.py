"""Shared helpers of the GPU tests (everything goes through the C ABI of libmc3d.so via monte_carlompi_b200.engine)."""
import numpy as np

from monte_carlompi_b200 import engine, ssp_fixtures

_ctx = None


def context():
    global _ctx
    if _ctx is None:
        _ctx = engine.Context([0])
    return _ctx


def fixture_table(kind, radius, k_lo, k_hi, imp_cnc=0.0, ssa_imp=0.30, ext_imp=12000.0):
    """SSP rows k_lo..k_hi (wavelength k/100 um) straight from the synthetic tables (nearest row; no file I/O)."""
    wvl, ssa, ext, g = ssp_fixtures.ice_table(kind, radius)
    rows = np.zeros(k_hi - k_lo + 1, engine.ROW_DTYPE)
    for j, k in enumerate(range(k_lo, k_hi + 1)):
        i = int(np.argmin(np.abs(wvl * 1e6 - k / 100.0)))
        ext_mix = ext[i] * (1 - imp_cnc) + ext_imp * imp_cnc
        p_imp = (imp_cnc * ext_imp) / (imp_cnc * ext_imp + (1 - imp_cnc) * ext[i])
        rows[j] = (k / 100.0, ssa[i], ssa_imp, g[i], ext_mix, p_imp)
    return rows


def const_table(ssa, g, ext=16.4, k=50, p_ext_imp=0.0, ssa_imp=0.3):
    rows = np.zeros(1, engine.ROW_DTYPE)
    rows[0] = (k / 100.0, ssa, ssa_imp, g, ext, p_ext_imp)
    return rows


def both_params(theta0_deg, tau_tot, r_lambert, wvl0, sigma, k_first, lambert_bottom, n_theta_bins=137, rho_snw=300.):
    """The same scalars for the CUDA path (engine.Params) and for the oracle (oracle.Params)."""
    from oracle import oracle
    th = np.pi * theta0_deg / 180.
    pe = engine.make_params(th, tau_tot, rho_snw, r_lambert, wvl0, sigma, k_first, lambert_bottom=lambert_bottom,
                            n_theta_bins=n_theta_bins)
    po = oracle.make_params(th, tau_tot, rho_snw, r_lambert, wvl0, sigma, k_first, lambert_bottom=lambert_bottom,
                            n_theta_bins=n_theta_bins)
    return pe, po


from gpu_util_cpu import tally_from_records  # noqa: E402,F401
