"""BRF and albedo straight from the GPU tallies -- no CSV round trip.

The reference's post_processing.py re-reads the per-photon text file with pandas and histograms it
(post_processing.py:46-100, 435-444).  libmc3d already returns, per wavelength row, the outcome counts and the
zenith histogram of the reflected photons with numpy's own binning, so the same quantities follow from a few
vector operations on ~60 KB; the wavenumber weights the reference applies per photon are applied per row here
(every photon of a row has the same wvn), in fp64.  ``*_from_records`` are the reference's formulas on per-photon
columns, kept for cross-checking (tests) and for files written by either implementation.
"""
import numpy as np

from .engine import N_COND


def calculate_bins(active_area=1., d_dome=175.):
    """Number of zenith bins that mimics a photodiode of ``active_area`` mm in a dome of ``d_dome`` mm
    (post_processing.py:435-444): int(pi d_dome / (4 active_area)) = 137 by default."""
    return int((np.pi * d_dome) / (4. * active_area))


def _wvn(table):
    return 1. / np.asarray(table['wvl_um'], dtype=np.float64)


def brf_from_tally(tally, table):
    """(midpoints [rad], brf) exactly as MonteCarloData.brf() computes them (post_processing.py:73-81):
    h = wvn-weighted histogram of theta_n for condition == 1 over (0, pi/2); Q_down = sum of wvn over all photons;
    brf = h / (Q_down w), w = sin cos / sum(sin cos) at the bin midpoints."""
    tally = np.asarray(tally)
    n_bins = tally.shape[1] - N_COND      # zenith-only tally (n_phi_bins <= 1); see brf2d_from_tally otherwise
    wvn = _wvn(table)
    q_down = float((tally[:, 0].astype(np.float64) * wvn).sum())
    h = (tally[:, N_COND:].astype(np.float64) * wvn[:, None]).sum(axis=0)
    edges = np.linspace(0., np.pi / 2, n_bins + 1)
    midpoints = (np.diff(edges) / 2.) + edges[:-1]
    weights = np.sin(midpoints) * np.cos(midpoints) / np.sum(np.sin(midpoints) * np.cos(midpoints))
    return midpoints, h / (q_down * weights)


def brf2d_from_tally(tally, table, n_theta_bins, n_phi_bins):
    """Full-hemisphere BRF(theta, phi) from a tally taken with ``n_phi_bins`` azimuth bins: the same normalisation
    as ``brf_from_tally`` with the zenith weight of a bin shared equally by its azimuth bins, so a Lambertian
    reflector gives its reflectance in every (theta, phi) bin.  Returns (theta midpoints, phi midpoints, brf[theta, phi])."""
    tally = np.asarray(tally)
    wvn = _wvn(table)
    q_down = float((tally[:, 0].astype(np.float64) * wvn).sum())
    h = (tally[:, N_COND:].astype(np.float64) * wvn[:, None]).sum(axis=0).reshape(n_theta_bins, n_phi_bins)
    te = np.linspace(0., np.pi / 2, n_theta_bins + 1)
    pe = np.linspace(0., 2 * np.pi, n_phi_bins + 1)
    tm, pm = (np.diff(te) / 2.) + te[:-1], (np.diff(pe) / 2.) + pe[:-1]
    w = np.sin(tm) * np.cos(tm) / np.sum(np.sin(tm) * np.cos(tm)) / n_phi_bins
    return tm, pm, h / (q_down * w[:, None])


def brf_from_records(condition, wvn, theta_n, n_bins):
    """The reference's own computation on per-photon columns (post_processing.py:73-81)."""
    condition, wvn, theta_n = np.asarray(condition), np.asarray(wvn, dtype=np.float64), np.asarray(theta_n, dtype=np.float64)
    q_down = wvn.sum()
    refl = condition == 1
    h = np.histogram(theta_n[refl], bins=n_bins, range=(0., np.pi / 2), weights=wvn[refl])
    midpoints = (np.diff(h[1]) / 2.) + h[1][:-1]
    weights = np.sin(midpoints) * np.cos(midpoints) / np.sum(np.sin(midpoints) * np.cos(midpoints))
    return midpoints, h[0] / (q_down * weights)


def albedo_from_tally(tally, table):
    """Black-sky albedo Q_up / Q_down weighted by wavenumber (monte_carlo3D.py:1659-1671, post_processing.py:266-275)."""
    wvn = _wvn(table)
    t = np.asarray(tally).astype(np.float64)
    return float((t[:, 1] * wvn).sum() / (t[:, 0] * wvn).sum())


def non_absorbed_fraction_from_tally(tally, table):
    """(Q_up + Q_diffuse + Q_direct) / Q_down (post_processing.py:270-275)."""
    wvn = _wvn(table)
    t = np.asarray(tally).astype(np.float64)
    return float(((t[:, 1] + t[:, 2] + t[:, 3]) * wvn).sum() / (t[:, 0] * wvn).sum())


def outcome_fractions(tally):
    """Unweighted fractions by condition: {1: reflected, 2: diffuse transmitted, 3: direct transmitted, 4: absorbed by
    ice, 5: absorbed by impurity} (README.md:62-66; 5 from monte_carlo3D.py:1465-1466)."""
    t = np.asarray(tally).astype(np.float64)
    n = t[:, 0].sum()
    return {c: float(t[:, c].sum() / n) for c in range(1, 6)}


def spectral_albedo_from_tally(tally, table):
    """Per-wavelength-row albedo (count ratio; within a row every photon has the same wvn) and the row wavelengths."""
    t = np.asarray(tally).astype(np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        alb = np.where(t[:, 0] > 0, t[:, 1] / t[:, 0], np.nan)
    return np.asarray(table['wvl_um'], dtype=np.float64), alb
