"""Drop-in ``MonteCarlo`` for the reference's monte_carloMPI/monte_carlo3D.py -- spheres / Henyey-Greenstein path.

Same driver surface as the reference (monte_carlo3D.py:42-94 constructor + config/flags 1778-1843, ``run``
1492-1657, ``setup_output`` 96-143): ``MonteCarlo(**model_kwargs).run(n_photon, wvl0, half_width, rds_snw,
theta_0=..., Lambertian_bottom=..., ...)`` reads ``./config.ini`` and the Mie NetCDF tables, walks ``n_photon``
photon packets and writes ``<output_dir>/sphere/WVL_FWHM_REFF_NPHOTON_THETA_HG.txt``, then prints the path.

What changed underneath: the Python photon loop (monte_carlo3D.py:1613-1616) and the mpi4py scatter / gather
(parallelize.py) are replaced by one call into libmc3d.so (hand-written sm_100a CUDA; engine.py is the ctypes
binding).  Wavelengths are drawn on the GPU from a Philox4x32-10 stream keyed on (seed, photon id), so the
per-photon SSP arrays of the reference become one small per-wavelength table (ssp.py).

Aspherical habits run with the Henyey-Greenstein phase function (``HG=True`` / ``--HG``, same walk, SSPs from the
habit's ``isca.dat``).  Not carried over (out of scope): the full phase matrix /
Stokes path (its data files are not part of the reference archive), plotting and debug demos.  Those raise
NotImplementedError rather than silently doing something else.
"""
import argparse
import configparser
import os

import numpy as np

from . import engine, output, ssp
from .parallelize import Parallel

TWO_PIE = 2 * np.pi
FOUR_PIE = 4 * np.pi

DEFAULT_N_THETA_BINS = 137   # int(pi * 175 / 4): post_processing.py:331-334, 435-444


class MonteCarlo(object):
    def __init__(self, **model_kwargs):
        """ valid model_kwargs (reference monte_carlo3D.py:44-60):
            tau_tot, imp_cnc, rho_snw, rho_ice, rsensor, hsensor, flg_crt, flg_3D, output_dir, optics_dir,
            fi_imp, HG, phase_functions
            additional, B200 build only: seed (int, default: OS entropy like the unseeded reference),
            devices (list of CUDA ordinals, default: all visible), n_theta_bins (BRF tally bins, default 137),
            n_phi_bins (azimuth bins of a full-hemisphere BRF tally, default 0 = zenith only)
        """
        model_args = self.get_model_args()
        model_args_dict = {'tau_tot': model_args.tau_tot,
                           'imp_cnc': model_args.imp_cnc,
                           'rho_snw': model_args.rho_snw,
                           'rho_ice': model_args.rho_ice,
                           'rsensor': model_args.rsensor,
                           'hsensor': model_args.hsensor,
                           'flg_crt': model_args.flg_crt,
                           'flg_3D': model_args.flg_3D,
                           'output_dir': model_args.output_dir,
                           'optics_dir': model_args.optics_dir,
                           'fi_imp': model_args.fi_imp,
                           'HG': model_args.HG,
                           'phase_functions': model_args.phase_functions,
                           'seed': None,
                           'devices': None,
                           'n_theta_bins': DEFAULT_N_THETA_BINS,
                           'n_phi_bins': 0}
        # kwargs given at instantiation win over config.ini / command line (monte_carlo3D.py:79-80)
        for kwarg, val in list(model_kwargs.items()):
            model_args_dict[kwarg] = val
        for key, val in model_args_dict.items():
            setattr(self, key, val)
        self._parallel = None
        self._rec_buf = None
        self.last_records = None
        self.last_tally = None
        self.last_table = None
        self.last_stats = None

    # ---- configuration (monte_carlo3D.py:1778-1843) -----------------------------------------------------------
    def get_model_args(self):
        """ Specify model kwargs at run time or get values from config.ini (read from the current directory)
        """
        config = configparser.ConfigParser()
        config.read('config.ini')

        section_name = 'model parameters'
        tau_tot = config.getfloat(section_name, 'tau_tot')
        imp_cnc = config.getfloat(section_name, 'imp_cnc')
        rho_snw = config.getfloat(section_name, 'rho_snw')
        rho_ice = config.getfloat(section_name, 'rho_ice')

        section_name = 'plot options'
        flg_crt = config.getint(section_name, 'flg_crt')
        flg_3D = config.getint(section_name, 'flg_3D')

        section_name = 'data'
        output_dir = config.get(section_name, 'output_dir')
        optics_dir = config.get(section_name, 'optics_dir')
        fi_imp = config.get(section_name, 'fi_imp')

        parser = argparse.ArgumentParser(description='[DESCRIPTION]')
        parser.add_argument('--tau_tot', type=float, default=tau_tot, help='snow optical depth')
        parser.add_argument('--imp_cnc', type=float, default=imp_cnc,
                            help='mass concentration of impurity [mIMP/(mIMP+mICE)]')
        parser.add_argument('--rho_snw', type=float, default=rho_snw,
                            help='snow density (kg/m3, only needed if flg_crt=1)')
        parser.add_argument('--rho_ice', type=float, default=rho_ice, help='ice density (kg/m3)')
        parser.add_argument('--rsensor', type=float, default=None, help='sensor radius [m]')
        parser.add_argument('--hsensor', type=float, default=None, help='sensor height above snow [m]')
        parser.add_argument('--flg_crt', type=int, default=flg_crt,
                            help='plot in optical depth space (=0) or Cartesian space (=1)?')
        parser.add_argument('--flg_3D', type=int, default=flg_3D,
                            help='plot in 2-D (=0), 3-D (=1). or no plot (=999)?')
        parser.add_argument('--output_dir', type=str, default=output_dir, help='directory to write output data')
        parser.add_argument('--optics_dir', type=str, default=optics_dir, help='directory of optics files')
        parser.add_argument('--fi_imp', type=str, default=fi_imp)
        parser.add_argument('--HG', action='store_true',
                            help='Use Henyey-Greenstein phase function instead of full scattering phase matrix '
                                 '(this is done automatically for spherical particles)')
        parser.add_argument('--phase_functions', action='store_true', help='Plot phase functions')
        return parser.parse_args()

    # ---- output (monte_carlo3D.py:96-143) ---------------------------------------------------------------------
    def setup_output(self, n_photon, wvl0, half_width):
        """ Create output dir for writing data to; returns output_file path
        """
        shape_dir = 'sphere'
        if getattr(self, 'shape', 'sphere') != 'sphere':            # monte_carlo3D.py:105-117
            shape_dir = (self.shape_dir, self.roughness_dir)
        return output.setup_output(self.output_dir, wvl0, half_width, self.snow_effective_radius, n_photon,
                                   self.theta_0, shape_dir=shape_dir)

    # ---- input preparation (monte_carlo3D.py:498-777, 1553-1588) ----------------------------------------------
    def _test_overrides(self):
        """The reference's ``test=True`` hook: attributes preset on the instance replace table values."""
        out = {}
        for name in ('ssa_ice', 'ext_cff_mss_ice', 'g', 'ssa_imp', 'ext_cff_mss_imp'):
            val = self.__dict__.get('_preset_' + name, None)
            if val is not None:
                out[name] = val
        return out

    def __setattr__(self, name, value):
        # remember user presets (test_case.ssa_ice = 0.9 ...) separately: run() overwrites self.ssa_ice etc. with
        # the per-wavelength arrays, exactly like the reference does at monte_carlo3D.py:1584-1588
        if name in ('ssa_ice', 'ext_cff_mss_ice', 'g', 'ssa_imp', 'ext_cff_mss_imp') and np.isscalar(value):
            self.__dict__['_preset_' + name] = value
        object.__setattr__(self, name, value)

    def build_table(self, wvl0, half_width, rds_snw, test=False, shape='sphere', roughness='smooth'):
        """Per-wavelength SSP rows covering every wavelength the Gaussian draw can produce (ssp.py): Mie spheres
        from the SNICAR NetCDF tables, or -- with ``--HG`` -- an aspherical habit from its ``isca.dat`` library."""
        scale = half_width / 2.355                                  # monte_carlo3D.py:1516
        k_lo, k_hi = ssp.wavelength_grid(wvl0, scale)
        overrides = self._test_overrides() if test else None
        if shape == 'sphere':
            self.snow_effective_radius = rds_snw                    # monte_carlo3D.py:505
            table = ssp.build_table(self.optics_dir, self.fi_imp, rds_snw, k_lo, k_hi, self.imp_cnc,
                                    overrides=overrides, quiet=bool(test))
        else:
            _, self.shape_dir, self.roughness_dir = ssp.aspherical_dirs(shape, roughness, wvl0)
            table, self.snow_effective_radius = ssp.build_table_aspherical(
                self.optics_dir, self.fi_imp, shape, roughness, wvl0, rds_snw, k_lo, k_hi, self.imp_cnc, self.rho_ice,
                overrides=overrides, quiet=bool(test))             # monte_carlo3D.py:1529-1545
        return table, k_lo, scale

    def _setup_case(self, n_photon, wvl0, half_width, rds_snw, theta_0, stokes_params, shape, roughness, test, debug,
                    Lambertian_surface, Lambertian_bottom, Lambertian_reflectance, seed):
        """Everything ``run`` does before the photon loop (monte_carlo3D.py:1498-1612): returns (params, table)."""
        if shape != 'sphere' and not self.HG:
            raise NotImplementedError('aspherical shapes are built for B200 with the Henyey-Greenstein phase function '
                                      'only (MonteCarlo(HG=True) / --HG); the full scattering phase matrix / Stokes '
                                      'path is out of scope (its data files are not part of the reference archive, '
                                      'README.md:40-42)')
        if debug:
            raise NotImplementedError('the two-scatter plotting demo (debug=True) is not part of the B200 build')
        if self.phase_functions:
            raise NotImplementedError('phase-function plotting is not part of the B200 build')
        self.debug = debug
        self.Lambertian_surface = Lambertian_surface
        self.Lambertian_bottom = Lambertian_bottom
        self.R_Lambertian = Lambertian_reflectance
        self.theta_0 = (np.pi * theta_0) / 180.                     # theta_0 deg -> rad, monte_carlo3D.py:1508
        self.shape = shape
        self.roughness = roughness
        self.wvl0 = wvl0
        self.initial_stokes_params = stokes_params
        n_photon = int(n_photon)

        table, k_first, scale = self.build_table(wvl0, half_width, rds_snw, test=test, shape=shape,
                                                 roughness=roughness)
        # same attribute names as the reference, one value per table row instead of per photon
        self.ext_cff_mss = table['ext_cff_mss']
        self.P_ext_imp = table['p_ext_imp']
        self.g = table['g']
        self.ssa_ice = table['ssa_ice']
        self.ssa_imp = table['ssa_imp']
        self.snow_depth = ssp.snow_depth(table, self.tau_tot, self.rho_snw)     # monte_carlo3D.py:1612

        if seed is None:
            seed = self.seed
        if seed is None:
            seed = int.from_bytes(os.urandom(8), 'little')          # the reference never seeds np.random
        self.last_seed = int(seed)

        params = engine.make_params(self.theta_0, self.tau_tot, self.rho_snw, Lambertian_reflectance, wvl0, scale,
                                    k_first, lambert_bottom=bool(Lambertian_bottom),
                                    lambert_surface=bool(Lambertian_surface), n_theta_bins=int(self.n_theta_bins),
                                    n_phi_bins=int(self.n_phi_bins))
        return params, table

    # ---- the run (monte_carlo3D.py:1492-1657) -----------------------------------------------------------------
    def run(self, n_photon, wvl0, half_width, rds_snw, theta_0=0., stokes_params=np.array([1, 0, 0, 0]),
            shape='sphere', roughness='smooth', test=False, debug=False, Lambertian_surface=False,
            Lambertian_bottom=True, Lambertian_reflectance=1., seed=None, write_output=True, first_photon_id=0):
        """ Run the Monte Carlo model given a normal distribution of wavelengths [um].
            ALL VALUES IN MICRONS

            write_output: True = the reference's text file (default); 'both' = text plus a binary sidecar
            <run>.npz (record columns, per-row wvn / snow depth, tallies, SSP table; output.load_run reads either);
            'binary' = the sidecar only (19 B instead of ~100 B per photon); False = nothing (results stay in
            self.last_records / self.last_tally).
            first_photon_id: id of photon 0 in the random stream ``seed`` (photon j is photon first_photon_id + j);
            case c of ``run_sweep`` is ``run(..., first_photon_id=c << 40)``.
        """
        params, table = self._setup_case(n_photon, wvl0, half_width, rds_snw, theta_0, stokes_params, shape, roughness,
                                         test, debug, Lambertian_surface, Lambertian_bottom, Lambertian_reflectance,
                                         seed)
        n_photon = int(n_photon)
        par = self._parallel
        if par is None:
            par = self._parallel = Parallel(n_photon, devices=self.devices)
        begin, count = par._map(n_photon)
        ctx = par.open()
        self.last_seed = par.broadcast_seed(self.last_seed)         # one stream for all ranks (rank 0's, if from entropy)
        tally = np.zeros((len(table), params.tally_width), np.uint64)
        records, stats = self._walk_records(ctx, params, table, self.last_seed, int(first_photon_id) + begin, count, tally)
        if par.size > 1:
            ctx.reduce_tally(tally, root=0)
        self.last_table, self.last_stats = table, stats

        all_answers = par.answer_and_reduce(records, MonteCarlo.flatten_list)
        if all_answers is None:
            return                                                  # not the root rank
        self.last_records, self.last_tally = all_answers, tally
        self._write(write_output, all_answers, tally, table, self.snow_depth, n_photon, wvl0, half_width)

    def _write(self, write_output, records, tally, table, snow_depth, n_photon, wvl0, half_width):
        """The output step of ``run`` (monte_carlo3D.py:1621-1648) for one case; returns the path printed (None when
        ``write_output`` is False)."""
        if not write_output:
            return None
        output_file = self.setup_output(n_photon, wvl0, half_width)
        if write_output != 'binary':
            output.write_run(output_file, records, 1. / table['wvl_um'], snow_depth)
        if write_output in ('both', 'binary'):
            sidecar = output.write_sidecar(output_file, records, 1. / table['wvl_um'], snow_depth, tally=tally, table=table)
            if write_output == 'binary':
                output_file = sidecar
        print('%s' % output_file)   # for easy post processing
        return output_file

    def _walk_records(self, ctx, params, table, seed, begin, count, tally):
        """One synchronous walk returning (record columns, stats).  The records come back packed (16 B per photon) into
        page-locked host memory -- copy-back at PCIe speed; the buffer is kept for reuse -- and are expanded on the
        host; tables beyond the packed format's 512 rows, or a walk beyond its 2^23 scatterings, take the columns."""
        if len(table) <= engine.PACKED_MAX_ROWS:
            if self._rec_buf is None or self._rec_buf.capacity < count:
                if self._rec_buf is not None:
                    self._rec_buf.free()
                self._rec_buf = engine.RecordBuffers(max(count, 1))
            stats = ctx.run_sync(params, table, seed, begin, count, self._rec_buf, tally)
            if not stats['packed_saturated']:
                return self._rec_buf.view(count), stats
        records = {name: np.empty(count, dtype=dt) for name, dt in engine.RECORD_COLUMNS}
        return records, ctx.run_sync(params, table, seed, begin, count, records, tally)

    # ---- sweeps (reference monte_carlo3D-run.py:60-96, 112-122: one run() per wavelength / grain size / angle) -------
    SWEEP_MAX_ROWS = 640            # rows of a batch's concatenated table (40 KB of shared memory per block)
    SWEEP_MAX_CASES = 128           # cases per batch
    SWEEP_MAX_PHOTONS = 1 << 27     # photons per batch (2 GiB of packed records in page-locked memory)

    def run_sweep(self, cases, write_output=True, seed=None):
        """Run many cases -- the loops of the reference's driver script over wavelength, grain size and zenith angle --
        with the SAME launches: the cases of a batch share one concatenated SSP table and their photons are walked
        together (libmc3d's mc3d_run_sweep), so a short case does not pay a launch and its tail, and long and short
        cases fill the GPU together.  Batches are pipelined on the context's slots: the copy-back and the text
        formatting of one overlap the walk of the next.

        ``cases``: iterable of dicts with the arguments of ``run`` (``n_photon, wvl0, half_width, rds_snw`` and
        optionally ``theta_0, Lambertian_bottom, Lambertian_reflectance, Lambertian_surface, test, shape, roughness,
        seed``).  All cases use one random stream ``seed`` (argument, else the instance's, else OS entropy): photon j
        of case c is photon ``(c << 40) + j`` of it, i.e. case c is bit-identical to
        ``run(..., seed=seed, first_photon_id=c << 40)``.  (A case dict with its own ``seed`` starts a new batch; its
        photon ids still carry its case index.)

        Returns one entry per case, on the root rank (``None`` per case on the other ranks of a one-process-per-GPU
        launch): the output path when ``write_output`` is set (True | 'both' | 'binary', as in ``run``), else
        ``(records, tally, table)``."""
        cases = [dict(c) for c in cases]
        if not cases:
            return []
        if len(cases) > engine.SWEEP_MAX_CASES:
            raise ValueError('run_sweep takes at most %d cases per call (the case index is part of the photon id): split '
                             'the sweep' % engine.SWEEP_MAX_CASES)
        if seed is None:
            seed = self.seed
        if seed is None:
            seed = int.from_bytes(os.urandom(8), 'little')
        par = self._parallel
        if par is None:
            par = self._parallel = Parallel(cases[0]['n_photon'], devices=self.devices)
        seed = par.broadcast_seed(int(seed))
        self.last_seed = int(seed)
        ctx = par.open()

        # ---- per-case inputs; cases with the same optics share their rows
        prepared = []
        for k, c in enumerate(cases):
            shape = c.get('shape', 'sphere')
            if shape != 'sphere' and not self.HG:
                raise NotImplementedError('aspherical shapes need HG=True (see run)')
            key = (c['wvl0'], c['half_width'], c['rds_snw'], shape, c.get('roughness', 'smooth'), bool(c.get('test', False)))
            table, k_first, scale = self.build_table(c['wvl0'], c['half_width'], c['rds_snw'], test=c.get('test', False),
                                                     shape=shape, roughness=c.get('roughness', 'smooth'))
            params = engine.make_params((np.pi * c.get('theta_0', 0.)) / 180., self.tau_tot, self.rho_snw,
                                        c.get('Lambertian_reflectance', 1.), c['wvl0'], scale, k_first,
                                        lambert_bottom=bool(c.get('Lambertian_bottom', True)),
                                        lambert_surface=bool(c.get('Lambertian_surface', False)),
                                        n_theta_bins=int(self.n_theta_bins), n_phi_bins=int(self.n_phi_bins))
            prepared.append(dict(k=k, c=c, key=key, table=table, params=params, n=int(c['n_photon']),
                                 seed=int(c.get('seed', seed)), r_eff=self.snow_effective_radius,
                                 dirs=(getattr(self, 'shape_dir', None), getattr(self, 'roughness_dir', None))))

        # ---- batches: consecutive cases, one seed, bounded table / case count / photons
        batches, cur = [], None
        for q in prepared:
            new_rows = 0 if cur is not None and q['key'] in cur['row_of'] else len(q['table'])
            if (cur is None or q['seed'] != cur['seed'] or cur['rows'] + new_rows > self.SWEEP_MAX_ROWS or
                    len(cur['cases']) >= self.SWEEP_MAX_CASES or cur['photons'] + q['n'] > self.SWEEP_MAX_PHOTONS):
                cur = dict(seed=q['seed'], rows=0, photons=0, cases=[], row_of={}, tables=[])
                batches.append(cur)
                new_rows = len(q['table'])
            if q['key'] not in cur['row_of']:
                cur['row_of'][q['key']] = cur['rows']
                cur['tables'].append(q['table'])
                cur['rows'] += new_rows
            q['row_begin'] = cur['row_of'][q['key']]
            cur['cases'].append(q)
            cur['photons'] += q['n']

        results = [None] * len(cases)
        depth = max(1, min(3, engine.N_SLOTS, len(batches)))
        cap = max(max(b['photons'] for b in batches), 1)
        bufs = [engine.RecordBuffers(-(-cap // par.size) + 1) for _ in range(depth)]
        pending = [None] * depth

        def launch(slot, b):
            table = np.concatenate(b['tables'])
            # the engine numbers a sweep's cases from 0: a batch that starts at case k0 of the call pads with empty
            # cases so that photon ids keep the case's index in the whole call
            k0 = b['cases'][0]['k']
            spec = [(b['cases'][0]['params'], 0, len(b['cases'][0]['table']), 0)] * k0
            spec += [(q['params'], q['row_begin'], len(q['table']), q['n']) for q in b['cases']]
            total = b['photons']
            begin, count = par._map(total)
            tally = np.zeros((len(table), b['cases'][0]['params'].tally_width), np.uint64)
            events = np.zeros(len(spec), np.uint64)
            ctx.run_sweep_async(slot, spec, table, b['seed'], bufs[slot], tally, events, range_begin=begin, range_count=count)
            pending[slot] = (b, table, tally, events, begin, count, k0)

        def finish(slot):
            b, table, tally, events, begin, count, k0 = pending[slot]
            pending[slot] = None
            stats = ctx.wait(slot)
            if stats['packed_saturated']:
                raise engine.Mc3dError('a walk of this sweep exceeded 2^23 scatterings: run that case with run()')
            rec = bufs[slot].view(count)
            if par.size > 1:
                ctx.reduce_tally(tally, root=0)
                ctx.reduce_tally(events, root=0)
            rec = par.answer_and_reduce(rec, MonteCarlo.flatten_list)
            self.last_stats = stats
            if rec is None:
                return                                              # not the root rank
            at = 0
            for q in b['cases']:
                r = {name: col[at:at + q['n']] for name, col in rec.items()}
                at += q['n']
                t = tally[q['row_begin']:q['row_begin'] + len(q['table'])].copy()
                shared = sum(1 for o in b['cases'] if o['row_begin'] == q['row_begin'])
                if shared > 1:                                      # cases that share rows share those tally rows too:
                    t = _tally_of_records(r, len(q['table']), q['params'])   # rebuild this case's own from its records
                c = q['c']
                depth_m = ssp.snow_depth(q['table'], self.tau_tot, self.rho_snw)
                self.last_records, self.last_tally, self.last_table = r, t, q['table']
                if write_output:
                    self.snow_effective_radius = q['r_eff']
                    self.shape = c.get('shape', 'sphere')
                    self.shape_dir, self.roughness_dir = q['dirs']
                    self.theta_0 = (np.pi * c.get('theta_0', 0.)) / 180.
                    results[q['k']] = self._write(write_output, r, t, q['table'], depth_m, q['n'], c['wvl0'], c['half_width'])
                else:
                    results[q['k']] = (r, t, q['table'])

        try:
            for j, b in enumerate(batches):
                slot = j % depth
                if pending[slot] is not None:
                    finish(slot)
                launch(slot, b)
            for j in range(len(batches) - depth, len(batches)):
                if j >= 0 and pending[j % depth] is not None:
                    finish(j % depth)
        finally:
            for slot in range(depth):
                if pending[slot] is not None:
                    try:
                        ctx.wait(slot)
                    except engine.Mc3dError:
                        pass
            for buf in bufs:
                buf.free()
        return results

    # ---- n_scat / path-length histograms without records (post_processing.py:162-223) -------------------------------
    def histograms(self, n_photon, wvl0, half_width, rds_snw, n_scat_bins=200, path_length_bins=1000, theta_0=0.,
                   test=False, Lambertian_surface=False, Lambertian_bottom=True, Lambertian_reflectance=1., seed=None):
        """The two histograms the reference's post-processing draws from the output file --
        ``np.histogram(n_scat, bins=200)`` and ``np.histogram(path_length * 100, bins=1000)`` [cm], each over the
        data's own (min, max) -- binned on the GPU(s), so no per-photon record leaves the device: a first pass finds
        the extrema, a second pass over the same photons (same seed) fills the bins.  Counts equal ``np.histogram``
        of the records of ``run`` with that seed.  Returns ``{'n_scat': (counts, edges), 'path_length_cm':
        (counts, edges)}`` on rank 0 (``None`` on the other ranks of a one-process-per-GPU launch)."""
        params, table = self._setup_case(n_photon, wvl0, half_width, rds_snw, theta_0, np.array([1, 0, 0, 0]), 'sphere',
                                         'smooth', test, False, Lambertian_surface, Lambertian_bottom,
                                         Lambertian_reflectance, seed)
        n_photon = int(n_photon)
        par = self._parallel
        if par is None:
            par = self._parallel = Parallel(n_photon, devices=self.devices)
        begin, count = par._map(n_photon)
        ctx = par.open()
        ctx.set_histograms()
        ctx.run_sync(params, table, self.last_seed, begin, count, None, None)
        ext = np.array(ctx.extrema(0) if count else (2**32 - 1, 0, np.inf, 0.), np.float64)
        parts = [np.frombuffer(b, np.float64) for b in par.allgather_bytes(ext.tobytes())]
        ns_lo, ns_hi = min(p[0] for p in parts), max(p[1] for p in parts)
        pl_lo, pl_hi = min(p[2] for p in parts) * 100., max(p[3] for p in parts) * 100.
        if ns_lo == ns_hi:                                           # np.histogram's rule for a degenerate range
            ns_lo, ns_hi = ns_lo - 0.5, ns_hi + 0.5
        if pl_lo == pl_hi:
            pl_lo, pl_hi = pl_lo - 0.5, pl_hi + 0.5
        ctx.set_histograms(n_scat_bins, (ns_lo, ns_hi), path_length_bins, (pl_lo, pl_hi), 100.)
        try:
            self.last_stats = ctx.run_sync(params, table, self.last_seed, begin, count, None, None)
            ns, pl = ctx.histograms(0)
        finally:
            ctx.set_histograms()
        if par.size > 1:
            both = np.concatenate([ns, pl])
            ctx.reduce_tally(both, root=0)
            ns, pl = both[:len(ns)], both[len(ns):]
            if par.rank != 0:
                return None
        return {'n_scat': (ns, np.linspace(ns_lo, ns_hi, n_scat_bins + 1)),
                'path_length_cm': (pl, np.linspace(pl_lo, pl_hi, path_length_bins + 1))}

    def close(self):
        if self._rec_buf is not None:
            self._rec_buf.free()
            self._rec_buf = None
        if self._parallel is not None:
            self._parallel.close()
            self._parallel = None

    def calculate_albedo(self, answers=None):
        """ Black sky albedo Q_up / Q_down weighted by wavenumber (monte_carlo3D.py:1659-1671), from the records of
            the last run (or a dict of record columns)
        """
        answers = self.last_records if answers is None else answers
        wvn = (1. / self.last_table['wvl_um'])[answers['wvl_row'].astype(np.int64)]
        return wvn[answers['condition'] == 1].sum() / wvn.sum()

    @classmethod
    def flatten_list(klass, l):
        """Concatenate per-rank record columns in rank order == photon order (monte_carlo3D.py:1845-1847)."""
        return {k: np.concatenate([part[k] for part in l]) for k in l[0]}


def _tally_of_records(rec, n_rows, params):
    """The tally block of one case from its record columns (same definition as the GPU's: counts by condition and
    np.histogram / np.histogram2d of the reflected photons' angles over float64(angle))."""
    n_theta, n_phi = int(params.n_theta_bins), max(1, int(params.n_phi_bins))
    t = np.zeros((n_rows, engine.N_COND + n_theta * n_phi), np.uint64)
    row = rec['wvl_row'].astype(np.int64)
    np.add.at(t, (row, 0), 1)
    np.add.at(t, (row, rec['condition'].astype(np.int64)), 1)
    refl = rec['condition'] == 1
    if n_theta > 0 and refl.any():
        th = rec['theta_n'][refl].astype(np.float64)
        r = row[refl]
        if n_phi > 1:
            ph = rec['phi_n'][refl].astype(np.float64)
            for k in np.unique(r):
                h, _, _ = np.histogram2d(th[r == k], ph[r == k], bins=(n_theta, n_phi), range=((0, np.pi / 2), (0, 2 * np.pi)))
                t[k, engine.N_COND:] += h.astype(np.uint64).ravel()
        else:
            for k in np.unique(r):
                h, _ = np.histogram(th[r == k], bins=n_theta, range=(0, np.pi / 2))
                t[k, engine.N_COND:] += h.astype(np.uint64)
    return t


def test(n_photon=50000, wvl=0.5, half_width=0.085, rds_snw=100, **run_kwargs):
    """ Test case for comparison with Wang et al (1995) Table 1, and van de Hulst (1980).  Albedo should be
        ~0.09739.  Total transmittance (diffuse+direct) should be ~0.66096 (monte_carlo3D.py:1849-1866; those
        numbers hold with the bottom boundary off: pass Lambertian_bottom=False).
    """
    test_case = MonteCarlo(tau_tot=2.0, imp_cnc=0)
    test_case.ssa_ice = 0.9
    test_case.g = 0.75
    test_case.run(n_photon, wvl, half_width, rds_snw, test=True, **run_kwargs)
    return test_case


def test_and_debug(n_photon=100, wvl=0.5, half_width=0.085, rds_snw=250, **run_kwargs):
    """ manually specified optical properties (monte_carlo3D.py:1868-1894)
    """
    test_case = MonteCarlo(tau_tot=10)
    test_case.ext_cff_mss_ice = 6.6
    test_case.ssa_ice = 0.999989859099
    test_case.g = -0.89
    test_case.ext_cff_mss_imp = 12000
    test_case.ssa_imp = 0.30
    test_case.run(n_photon, wvl, half_width, rds_snw, test=True, **run_kwargs)
    return test_case
