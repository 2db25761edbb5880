#!/usr/bin/env python
"""Driver script, same user surface as the reference's monte_carlo3D-run.py: edit the USER INPUT block and run

    python monte_carlo3D-run.py [--tau_tot 10 --optics_dir inputdata ...]

One process drives every visible B200 (no mpirun needed); launching one process per GPU with torchrun / mpirun
also works (monte_carlompi_b200/parallelize.py).  If ``inputdata/`` holds no Mie tables (the Bohren & Huffman
tarball of the reference is a separate download), set MC3D_SYNTHETIC_OPTICS=1 to have deterministic synthetic
tables written there first (monte_carlompi_b200/ssp_fixtures.py).
"""
import os

import numpy as np
from monte_carloMPI import monte_carlo3D

DEBUG = False
LAMBERTIAN_SURFACE = False      # True simulates a bare Lambertian surface instead of snow (monte_carlo3D.py:1228-1229)
LAMBERTIAN_BOTTOM = True        # Lambertian lower boundary with the reflectance below
LAMBERTIAN_REFLECTANCE = 0.5    # reflectance of the underlying surface beneath the snow


def run():
    """ USER INPUT
    """
    # number of photon packets
    n_photon = 10000

    # incidence zenith angle (degrees)
    theta_0 = 15.

    # initial Stokes parameters (unused by the Henyey-Greenstein path, kept for the reference's signature)
    stokes_params = np.array([1, 0, 0, 0])

    shape = 'sphere'
    roughness = 'smooth'

    # centre wavelength and full width at half maximum [um]
    wvl = 1.3
    half_width = 0.085

    multiple_wavelengths = False
    many_grain_sizes = False

    if multiple_wavelengths:
        wvls = [1.3, 1.55]
        half_widths = [0.085, 0.130]
        for i, wvl in enumerate(wvls):
            rds_snw = np.array([50, 100, 250, 500, 1000])
            multiple_grain_sizes(n_photon, wvl, half_widths[i], rds_snw, theta_0=theta_0,
                                 stokes_params=stokes_params, shape=shape, roughness=roughness)
    elif many_grain_sizes:
        rds_snw = np.array([50, 100, 250, 500, 1000])
        multiple_grain_sizes(n_photon, wvl, half_width, rds_snw, theta_0=theta_0, stokes_params=stokes_params,
                             shape=shape, roughness=roughness)
    else:
        # snow effective grain radius [um]
        rds_snw = 100.
        single_grain_size(n_photon, wvl, half_width, rds_snw, theta_0=theta_0, stokes_params=stokes_params,
                          shape=shape, roughness=roughness)
    """ END USER INPUT
    """


def _model():
    model = monte_carlo3D.MonteCarlo()
    if os.environ.get('MC3D_SYNTHETIC_OPTICS'):
        from monte_carlompi_b200 import ssp_fixtures
        if not os.path.isdir(os.path.join(model.optics_dir, 'mie', 'snicar')):
            ssp_fixtures.write_optics_dir(model.optics_dir, 'spectral', (50, 100, 250, 500, 1000), model.fi_imp)
    return model


def single_grain_size(n_photon, wvl, half_width, rds_snw, theta_0=0., stokes_params=np.array([1, 0, 0, 0]),
                      shape='sphere', roughness='smooth'):
    monte_carlo_run = _model()
    monte_carlo_run.run(n_photon, wvl, half_width, rds_snw, theta_0=theta_0, stokes_params=stokes_params,
                        shape=shape, roughness=roughness, debug=DEBUG, Lambertian_surface=LAMBERTIAN_SURFACE,
                        Lambertian_bottom=LAMBERTIAN_BOTTOM, Lambertian_reflectance=LAMBERTIAN_REFLECTANCE)
    monte_carlo_run.close()


def multiple_grain_sizes(n_photon, wvl, half_width, rds_snw, theta_0=0., stokes_params=np.array([1, 0, 0, 0]),
                         shape='sphere', roughness='smooth'):
    monte_carlo_run = _model()
    # one case per grain size, several in flight on the GPU (the reference calls run() once per radius)
    monte_carlo_run.run_sweep([dict(n_photon=n_photon, wvl0=wvl, half_width=half_width, rds_snw=rds, theta_0=theta_0,
                                    Lambertian_surface=LAMBERTIAN_SURFACE, Lambertian_bottom=LAMBERTIAN_BOTTOM,
                                    Lambertian_reflectance=LAMBERTIAN_REFLECTANCE) for rds in rds_snw])
    monte_carlo_run.close()


def main():
    run()


if __name__ == '__main__':
    main()
