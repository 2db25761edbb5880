#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference (run in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

    python oracle/make_golden.py            # regenerate every case

Each case runs the reference's MonteCarlo.run single-rank under oracle/ref_shim.py with np.random.seed(seed) and
stores the per-photon answer columns, the de-duplicated SSP rows the reference derived, and the per-photon
offsets into the random stream.  The stream itself is NOT stored: numpy's legacy MT19937 RandomState is frozen,
so tests regenerate it with ``regenerate_stream`` below (checked here against what the reference consumed).
"""
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from monte_carlompi_b200 import ssp_fixtures  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# name -> (n_photon, wvl0, half_width, rds_snw, theta_0_deg, seed, fixture kind, model_kwargs, run_kwargs, preset)
CASES = {
    # BASELINE.json configs[0] shape (driver defaults, monte_carlo3D-run.py:6-18, 48, 54, 84), fewer photons
    'c1_default': (2000, 1.3, 0.085, 100., 15., 20190603, 'spectral', {}, dict(Lambertian_bottom=True,
                   Lambertian_reflectance=0.5), None),
    # BASELINE.json configs[0] at its full size: the driver default, n_photon = 10000
    'c1_full_10k': (10000, 1.3, 0.085, 100., 15., 424242, 'spectral', {}, dict(Lambertian_bottom=True,
                    Lambertian_reflectance=0.5), None),
    # finite slabs: Lambertian bottom reflections, direct / diffuse transmission
    'slab_tau3_lb': (1500, 1.3, 0.085, 100., 15., 11, 'spectral', dict(tau_tot=3.0), dict(Lambertian_bottom=True,
                     Lambertian_reflectance=0.5), None),
    'slab_tau05_normal': (1500, 1.3, 0.085, 100., 0., 12, 'spectral', dict(tau_tot=0.5),
                          dict(Lambertian_bottom=True, Lambertian_reflectance=0.5), None),
    'slab_tau3_black': (1500, 1.0, 0.085, 250., 60., 13, 'spectral', dict(tau_tot=3.0),
                        dict(Lambertian_bottom=False), None),
    # impurity species branch (condition 5)
    'impurity': (1200, 1.3, 0.085, 100., 30., 14, 'spectral', dict(tau_tot=3.0, imp_cnc=1e-5),
                 dict(Lambertian_bottom=True, Lambertian_reflectance=0.5), None),
    # known-answer setup of monte_carlo3D.py:1849-1866 (with the bottom boundary off)
    'kat_vdh': (4000, 0.5, 0.085, 100., 0., 15, 'const-kat', dict(tau_tot=2.0, imp_cnc=0),
                dict(Lambertian_bottom=False), dict(ssa_ice=0.9, g=0.75)),
    # constants of monte_carlo3D.py:1871-1884 (visible, long walks, negative g as written there, BC impurity)
    'vis_debug': (300, 0.5, 0.085, 250., 15., 16, 'const-vis', dict(tau_tot=10, imp_cnc=1e-7),
                  dict(Lambertian_bottom=True, Lambertian_reflectance=1.0),
                  dict(ext_cff_mss_ice=6.6, ssa_ice=0.999989859099, g=-0.89, ext_cff_mss_imp=12000, ssa_imp=0.30)),
    # weakly absorbing visible ice, forward peaked, semi-infinite: thousands of events per photon
    'vis_long': (120, 0.5, 0.085, 250., 15., 19, 'const-vis', {}, dict(Lambertian_bottom=True,
                 Lambertian_reflectance=0.5), dict(ext_cff_mss_ice=6.6, ssa_ice=0.999989859099, g=0.89)),
    # Lambertian_surface mode: the snow replaced by a Lambertian reflector (monte_carlo3D.py:1228-1250, 1385-1387)
    'lambert_surface': (3000, 1.3, 0.085, 100., 40., 20, 'spectral', dict(tau_tot=5.0),
                        dict(Lambertian_surface=True, Lambertian_bottom=False, Lambertian_reflectance=0.7), None),
    # isotropic scattering: the g == 0 branch of Henyey_Greenstein2 (monte_carlo3D.py:794-795)
    'isotropic': (1500, 0.5, 0.085, 100., 45., 17, 'const-kat', dict(tau_tot=5.0), dict(Lambertian_bottom=False),
                  dict(ssa_ice=0.9, g=0.0)),
    # wide band, out-of-table wavelengths exercise the nearest-row fallback (monte_carlo3D.py:533-537)
    'edge_of_table': (800, 0.33, 0.26, 100., 15., 18, 'spectral', dict(tau_tot=4.0),
                      dict(Lambertian_bottom=True, Lambertian_reflectance=0.3), None),
}


def regenerate_stream(seed, n_photon, wvl0, half_width, n_walk_draws):
    """The reference's draw order from a fresh legacy RandomState(seed):
    normal(size=n) -> 3 uniforms per photon (initial_pdfs) -> the walk draws, photon after photon."""
    rs = np.random.RandomState(seed)
    wvls = np.around(rs.normal(loc=wvl0, scale=half_width / 2.355, size=(n_photon)), decimals=2)
    init = rs.random_sample(3 * n_photon)
    stream = rs.random_sample(int(n_walk_draws))
    return wvls, init, stream


def make_case(name, optics_root):
    n, wvl0, hw, rds, theta, seed, kind, mkw, rkw, preset = CASES[name]
    optics = os.path.join(optics_root, kind)
    if not os.path.isdir(optics):
        ssp_fixtures.write_optics_dir(optics, kind, (50, 100, 250, 500, 1000))
    r = ref_shim.run_reference(n, wvl0, hw, rds, theta, seed=seed, optics_dir=optics, model_kwargs=mkw,
                               run_kwargs=rkw, preset=preset, record=True)
    wvls, init, stream = regenerate_stream(seed, n, wvl0, hw, r['offsets'][-1])
    assert np.array_equal(wvls, r['wvl']) and np.array_equal(init, r['init_draws'])
    assert np.array_equal(stream, r['stream']), 'legacy RandomState stream is not what the reference consumed'
    k = np.rint(r['wvl'] * 100).astype(np.int64)
    assert np.array_equal(k / 100.0, r['wvl'])
    uk, first = np.unique(k, return_index=True)
    rows = np.zeros(len(uk), dtype=[('wvl_um', 'f8'), ('ssa_ice', 'f8'), ('ssa_imp', 'f8'), ('g', 'f8'),
                                    ('ext_cff_mss', 'f8'), ('p_ext_imp', 'f8')])
    rows['wvl_um'] = uk / 100.0
    for col, src in (('ssa_ice', 'ssa_ice'), ('ssa_imp', 'ssa_imp'), ('g', 'g'), ('ext_cff_mss', 'ext_cff_mss'),
                     ('p_ext_imp', 'P_ext_imp')):
        rows[col] = r[src][first]
        assert np.array_equal(rows[col][np.searchsorted(uk, k)], r[src]), col  # SSPs depend on wavelength only
    cfg = dict(n_photon=n, wvl0=wvl0, half_width=hw, rds_snw=rds, theta_0=theta, seed=seed, fixture=kind,
               tau_tot=float(mkw.get('tau_tot', 1e6)), imp_cnc=float(mkw.get('imp_cnc', 0.0)), rho_snw=300.0,
               Lambertian_bottom=bool(rkw.get('Lambertian_bottom', True)),
               Lambertian_surface=bool(rkw.get('Lambertian_surface', False)),
               Lambertian_reflectance=float(rkw.get('Lambertian_reflectance', 1.0)),
               preset=repr(preset))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + '.npz'),
                        config=np.array(repr(cfg)),
                        condition=r['condition'].astype(np.int8), wvl_k=k.astype(np.int16),
                        theta_n=r['theta_n'], phi_n=r['phi_n'], n_scat=r['n_scat'].astype(np.int32),
                        path_length=r['path_length'], offsets=r['offsets'], rows=rows)
    print('%-18s n=%5d draws=%9d  cond=%s  mean n_scat=%.1f' % (
        name, n, r['offsets'][-1], np.bincount(r['condition'], minlength=6)[1:].tolist(), r['n_scat'].mean()))


def make_text_case(optics_root):
    """Golden output file (name + bytes) of a tiny default run.  The unpinned reference prints
    'np.float64(x)' under numpy >= 2 (an artefact); the golden keeps the numpy<2 form 'x'."""
    optics = os.path.join(optics_root, 'spectral')
    r = ref_shim.run_reference(40, 1.3, 0.085, 100., 15., seed=5, optics_dir=optics, model_kwargs={},
                               run_kwargs=dict(Lambertian_bottom=True, Lambertian_reflectance=0.5), record=False,
                               keep_text=True)
    name, text = r['text']
    text = re.sub(r'np\.float64\(([^)]*)\)', r'\1', text)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'text_default.npz'), name=np.array(name), text=np.array(text),
                        condition=r['condition'].astype(np.int8), wvl_k=np.rint(r['wvl'] * 100).astype(np.int16),
                        theta_n=r['theta_n'], phi_n=r['phi_n'], n_scat=r['n_scat'].astype(np.int32),
                        path_length=r['path_length'], snow_depth=r['snow_depth'], wvn=r['wvn'])
    print('text_default       %s (%d bytes)' % (name, len(text)))


# aspherical habit with --HG (monte_carlo3D.py:173-266, 316-336, 396-415, 1529-1545): SSPs from the habit's isca.dat,
# nearest library wavelength replaces the photon's wavelength
ASPHERICAL = dict(name='aspherical_hg', n_photon=1500, wvl0=1.3, half_width=0.26, rds_snw=80., theta_0=30., seed=21,
                  shape='droxtal', roughness='moderately rough', tau_tot=4.0, imp_cnc=1e-6)


def make_aspherical_case(optics_root):
    c = ASPHERICAL
    optics = os.path.join(optics_root, 'spectral')
    if not os.path.isdir(os.path.join(optics, 'mie')):
        ssp_fixtures.write_optics_dir(optics, 'spectral', (50, 100, 250, 500, 1000))
    ssp_fixtures.write_isca(optics, 'droxtal', 'Rough003')
    mkw = dict(tau_tot=c['tau_tot'], imp_cnc=c['imp_cnc'], HG=True)
    rkw = dict(shape=c['shape'], roughness=c['roughness'], Lambertian_bottom=True, Lambertian_reflectance=0.5)
    r = ref_shim.run_reference(c['n_photon'], c['wvl0'], c['half_width'], c['rds_snw'], c['theta_0'], seed=c['seed'],
                               optics_dir=optics, model_kwargs=mkw, run_kwargs=rkw, record=True, keep_text=True)
    drawn, init, stream = regenerate_stream(c['seed'], c['n_photon'], c['wvl0'], c['half_width'], r['offsets'][-1])
    assert np.array_equal(init, r['init_draws']) and np.array_equal(stream, r['stream'])
    k = np.rint(drawn * 100).astype(np.int64)
    uk, first = np.unique(k, return_index=True)
    rows = np.zeros(len(uk), dtype=[('wvl_um', 'f8'), ('ssa_ice', 'f8'), ('ssa_imp', 'f8'), ('g', 'f8'),
                                    ('ext_cff_mss', 'f8'), ('p_ext_imp', 'f8')])
    idx = np.searchsorted(uk, k)
    for col, src in (('wvl_um', 'wvl'), ('ssa_ice', 'ssa_ice'), ('ssa_imp', 'ssa_imp'), ('g', 'g'),
                     ('ext_cff_mss', 'ext_cff_mss'), ('p_ext_imp', 'P_ext_imp')):
        rows[col] = r[src][first]
        assert np.array_equal(rows[col][idx], r[src]), col          # everything depends on the drawn wavelength only
    assert np.array_equal(r['wvn'], 1.0 / rows['wvl_um'][idx])
    cfg = dict(n_photon=c['n_photon'], wvl0=c['wvl0'], half_width=c['half_width'], rds_snw=c['rds_snw'],
               theta_0=c['theta_0'], seed=c['seed'], fixture='spectral', tau_tot=c['tau_tot'], imp_cnc=c['imp_cnc'],
               rho_snw=300.0, rho_ice=917.0, Lambertian_bottom=True, Lambertian_surface=False,
               Lambertian_reflectance=0.5, preset='None', shape=c['shape'], roughness=c['roughness'])
    name, text = r['text']
    np.savez_compressed(os.path.join(GOLDEN_DIR, c['name'] + '.npz'), config=np.array(repr(cfg)),
                        condition=r['condition'].astype(np.int8), wvl_k=k.astype(np.int16), rows_k=uk.astype(np.int16),
                        theta_n=r['theta_n'], phi_n=r['phi_n'], n_scat=r['n_scat'].astype(np.int32),
                        path_length=r['path_length'], offsets=r['offsets'], rows=rows, file_name=np.array(name))
    print('%-18s n=%5d draws=%9d  cond=%s  mean n_scat=%.1f  %s' % (
        c['name'], c['n_photon'], r['offsets'][-1], np.bincount(r['condition'], minlength=6)[1:].tolist(),
        r['n_scat'].mean(), name))


def main():
    if not ref_shim.reference_available():
        raise SystemExit('reference not present; fixtures can only be regenerated in the build container')
    root = tempfile.mkdtemp(prefix='mc3d_optics_')
    for name in (sys.argv[1:] or CASES):
        if name == ASPHERICAL['name']:
            make_aspherical_case(root)
        else:
            make_case(name, root)
    if not sys.argv[1:]:
        make_aspherical_case(root)
        make_text_case(root)


if __name__ == '__main__':
    main()
