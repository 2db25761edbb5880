"""Output file naming and writing -- byte-compatible with the reference.

Reference: MonteCarlo.setup_output (monte_carloMPI/monte_carlo3D.py:96-143) and the writer block of run()
(monte_carlo3D.py:1621-1648).  post_processing.py:24-38 parses the file name's first three '_' fields and reads
the body with pandas.read_csv(delim_whitespace=True), so both are kept exactly:

    <output_dir>/sphere/<wvl0>_<half_width>_<rds_snw>_<n_photon>_<theta0_deg>_HG[_N].txt
    condition wvn[um^-1] theta_n phi_n n_scat path_length[m], snow_depth[m]          <- header, comma included
    %d %r %r %r %d %r %r                                                              <- one line per photon

``%r`` of a float is Python's shortest round-trip repr (the form numpy < 2 scalars print; the ``np.float64(...)``
wrapper the unpinned reference would emit under numpy >= 2 is an environment artefact and is not imitated).
"""
import os

import numpy as np

HEADER = 'condition wvn[um^-1] theta_n phi_n n_scat path_length[m], snow_depth[m]\n'


def theta0_deg_for_name(theta_0_rad):
    """The reference formats np.rad2deg(self.theta_0) with self.theta_0 = pi * theta_0 / 180 (monte_carlo3D.py:1508,
    119): the round trip is part of the file name (15 deg -> 14.999999999999998)."""
    return np.rad2deg(theta_0_rad)


def run_name(wvl0, half_width, rds_snw, n_photon, theta_0_rad, appendix='HG', suffix=None):
    theta0_deg = theta0_deg_for_name(theta_0_rad)
    if suffix is None:
        return '%s_%s_%s_%s_%s_%s.txt' % (wvl0, half_width, rds_snw, n_photon, _plain(theta0_deg), appendix)
    return '%s_%s_%s_%s_%s_%s_%d.txt' % (wvl0, half_width, rds_snw, n_photon, _plain(theta0_deg), appendix, suffix)


def _plain(x):
    """'%s' of a numpy float64 prints the bare number under every numpy version."""
    return float(x)


def setup_output(output_dir, wvl0, half_width, rds_snw, n_photon, theta_0_rad, shape_dir='sphere'):
    """Create <output_dir>/<shape_dir>/ on demand and return a path that does not exist yet (``_N`` de-duplication
    suffix, monte_carlo3D.py:135-141).  ``shape_dir`` is 'sphere', or a (shape_dir, roughness_dir) pair for the
    aspherical habits (monte_carlo3D.py:108-117)."""
    if not os.path.isdir(output_dir):
        os.mkdir(output_dir)
    save_dir = output_dir
    for part in ((shape_dir,) if isinstance(shape_dir, str) else tuple(shape_dir)):
        save_dir = os.path.join(save_dir, part)
        if not os.path.isdir(save_dir):
            os.mkdir(save_dir)
    path = os.path.join(save_dir, run_name(wvl0, half_width, rds_snw, n_photon, theta_0_rad))
    i = 0
    while os.path.isfile(path) or os.path.isfile(sidecar_path(path)):   # a binary-only run occupies its name too
        i += 1
        path = os.path.join(save_dir, run_name(wvl0, half_width, rds_snw, n_photon, theta_0_rad, suffix=i))
    return path


def format_lines(condition, wvn, theta_n, phi_n, n_scat, path_length, snow_depth):
    """Body of the output file as one string; all arguments are equal-length sequences.  Floats are widened to
    float64 first (the reference's columns are float64), so a float32 theta prints as the double it equals."""
    cols = (np.asarray(condition).astype(np.int64).tolist(),
            np.asarray(wvn, dtype=np.float64).tolist(),
            np.asarray(theta_n, dtype=np.float64).tolist(),
            np.asarray(phi_n, dtype=np.float64).tolist(),
            np.asarray(n_scat).astype(np.int64).tolist(),
            np.asarray(path_length, dtype=np.float64).tolist(),
            np.asarray(snow_depth, dtype=np.float64).tolist())
    return ''.join(['%d %r %r %r %d %r %r\n' % row for row in zip(*cols)])


def write_run(path, records, wvn_by_row, snow_depth_by_row):
    """Header + one line per photon from the compact record columns of libmc3d (``wvl_row`` indexes the per-row
    ``wvn`` / ``snow_depth`` tables).  Lines are formatted by the native multithreaded writer
    (csrc/format_records.cpp), byte-identical to ``format_lines`` and ~20x faster."""
    from . import engine
    with open(path, 'w') as f:
        f.write(HEADER)
    engine.write_records_text(path, records, wvn_by_row, snow_depth_by_row, append=True)
    return path


def write_records(path, condition, wvn, theta_n, phi_n, n_scat, path_length, snow_depth, chunk=1 << 18):
    """Write header + one line per photon, in photon order (monte_carlo3D.py:1623-1636)."""
    n = len(condition)
    with open(path, 'w') as f:
        f.write(HEADER)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            f.write(format_lines(condition[lo:hi], wvn[lo:hi], theta_n[lo:hi], phi_n[lo:hi], n_scat[lo:hi],
                                 path_length[lo:hi], snow_depth[lo:hi]))
    return path


# ---- optional binary sidecar ------------------------------------------------------------
COLUMNS = ('condition', 'wvn[um^-1]', 'theta_n', 'phi_n', 'n_scat', 'path_length[m],', 'snow_depth[m]')


def sidecar_path(path):
    """<run>.txt -> <run>.npz"""
    return (path[:-4] if path.endswith('.txt') else path) + '.npz'


def write_sidecar(path, records, wvn_by_row, snow_depth_by_row, tally=None, table=None):
    """Write the run as a compressed-free .npz next to (or instead of) the text file: the compact record columns
    of libmc3d plus the per-row ``wvn`` / ``snow_depth`` tables (19 bytes per photon instead of ~100 bytes of text),
    and optionally the GPU tallies and the SSP table.  ``load_run`` gives back exactly what
    ``pd.read_csv(<run>.txt, delim_whitespace=True)`` (post_processing.py:38) would."""
    out = sidecar_path(path)
    extra = {}
    if tally is not None:
        extra['tally'] = np.asarray(tally)
    if table is not None:
        extra['table'] = np.asarray(table)
    np.savez(out, wvn_by_row=np.asarray(wvn_by_row, dtype=np.float64),
             snow_depth_by_row=np.asarray(snow_depth_by_row, dtype=np.float64),
             **{k: np.asarray(v) for k, v in records.items()}, **extra)
    return out


def load_run(path):
    """Read a run written by this package -- the reference's text file or the binary sidecar -- into a pandas
    DataFrame with the reference's column names (note the comma the reference's header carries after
    ``path_length[m]``).  Both sources give identical frames: float32 columns widen to the doubles the text holds."""
    import pandas as pd
    if path.endswith('.npz'):
        z = np.load(path)
        rows = z['wvl_row'].astype(np.int64)
        return pd.DataFrame({'condition': z['condition'].astype(np.int64), 'wvn[um^-1]': z['wvn_by_row'][rows],
                             'theta_n': z['theta_n'].astype(np.float64), 'phi_n': z['phi_n'].astype(np.float64),
                             'n_scat': z['n_scat'].astype(np.int64), 'path_length[m],': z['path_length'].astype(np.float64),
                             'snow_depth[m]': z['snow_depth_by_row'][rows]}, columns=list(COLUMNS))
    return pd.read_csv(path, sep=r'\s+', float_precision='round_trip')
