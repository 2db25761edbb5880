"""numpy-only helpers shared by CPU and GPU tests."""
import numpy as np

from monte_carlompi_b200 import engine


def tally_from_records(rec, n_rows, n_theta_bins):
    """What the device tally must equal: counts by (row, condition) and np.histogram of float64(theta) of the
    reflected photons (post_processing.py:73-76)."""
    t = np.zeros((n_rows, engine.N_COND + n_theta_bins), np.uint64)
    row = rec['wvl_row'].astype(np.int64)
    cond = rec['condition'].astype(np.int64)
    np.add.at(t, (row, np.zeros_like(row)), 1)
    np.add.at(t, (row, cond), 1)
    if n_theta_bins:
        m = cond == 1
        edges = np.linspace(0., np.pi / 2, n_theta_bins + 1)
        th = rec['theta_n'][m].astype(np.float64)
        b = np.clip(np.searchsorted(edges, th, side='right') - 1, 0, n_theta_bins - 1)
        # cross-check the vectorised binning against numpy's own histogram
        assert np.array_equal(np.bincount(b, minlength=n_theta_bins),
                              np.histogram(th, bins=n_theta_bins, range=(0., np.pi / 2))[0])
        np.add.at(t, (row[m], engine.N_COND + b), 1)
    return t
