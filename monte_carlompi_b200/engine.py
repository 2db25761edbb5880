"""ctypes binding of libmc3d.so (include/mc3d.h) -- the only way this package computes anything.

There is no CPU fallback: if the shared library is missing, or no B200 is visible, the constructors raise.
No PyTorch, no Triton; numpy arrays in, numpy arrays out.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MC3D_LIB') or os.path.join(_HERE, 'libmc3d.so')   # MC3D_LIB: A/B-test another build

N_COND = 8
N_SLOTS = 16
FLAG_LAMBERT_BOTTOM = 1
FLAG_LAMBERT_SURFACE = 2
ABI_VERSION = 4
PATH_AUTO, PATH_FUSED, PATH_PERSISTENT = 0, 1, 2
PACKED_MAX_ROWS = 512          # packed 16-byte records hold the SSP row in 9 bits
PACKED_NSCAT_MAX = 0x7fffff    # ... and n_scat in 23 bits (saturating; Stats.packed_saturated reports it)

EXPORTS = ('mc3d_abi_version', 'mc3d_last_error', 'mc3d_query', 'mc3d_create', 'mc3d_nccl_unique_id',
           'mc3d_create_rank', 'mc3d_destroy', 'mc3d_host_alloc', 'mc3d_host_free', 'mc3d_run', 'mc3d_run_async',
           'mc3d_wait', 'mc3d_reduce_tally', 'mc3d_replay', 'mc3d_set_launch', 'mc3d_write_records_text', 'mc3d_py_repr',
           'mc3d_set_histograms', 'mc3d_get_histograms', 'mc3d_records_layout', 'mc3d_set_input_caching',
           'mc3d_unpack_records', 'mc3d_set_walk_path', 'mc3d_run_sweep', 'mc3d_run_sweep_async', 'mc3d_set_tail_kernel')


class Mc3dError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [('theta0_rad', C.c_double), ('tau_tot', C.c_double), ('rho_snw', C.c_double),
                ('r_lambert', C.c_double), ('wvl0_um', C.c_double), ('sigma_um', C.c_double),
                ('k_first', C.c_int32), ('flags', C.c_uint32), ('n_theta_bins', C.c_int32),
                ('n_phi_bins', C.c_int32)]

    @property
    def tally_width(self):
        """uint64 entries per wavelength row of the tally block."""
        return N_COND + self.n_theta_bins * max(1, self.n_phi_bins)


class SweepCase(C.Structure):
    """mc3d_sweep_case: one case of a sweep -- its scalars, its rows in the concatenated table, its photon count."""
    _fields_ = [('params', Params), ('row_begin', C.c_int32), ('n_rows', C.c_int32), ('n_photon', C.c_uint64)]


SWEEP_MAX_CASES = 1024
SWEEP_ID_SHIFT = 40            # photon j of case c is photon id (c << 40) + j


class Records(C.Structure):
    _fields_ = [('condition', C.c_void_p), ('wvl_row', C.c_void_p), ('theta_n', C.c_void_p),
                ('phi_n', C.c_void_p), ('n_scat', C.c_void_p), ('path_length', C.c_void_p), ('packed', C.c_void_p)]


class RecordsF64(C.Structure):
    _fields_ = [('condition', C.c_void_p), ('wvn', C.c_void_p), ('theta_n', C.c_void_p), ('phi_n', C.c_void_p),
                ('n_scat', C.c_void_p), ('path_length', C.c_void_p), ('snow_depth', C.c_void_p),
                ('consumed', C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [('n_photon', C.c_uint64), ('n_events', C.c_uint64), ('kernel_ms', C.c_double),
                ('total_ms', C.c_double), ('n_devices', C.c_int32), ('sm_count', C.c_int32),
                ('sm_clock_khz', C.c_int32), ('grid_blocks', C.c_int32), ('block_threads', C.c_int32),
                ('packed_saturated', C.c_int32), ('walk_path', C.c_int32), ('reserved', C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class HistSpec(C.Structure):
    _fields_ = [('n_scat_bins', C.c_int32), ('path_bins', C.c_int32), ('n_scat_lo', C.c_double),
                ('n_scat_hi', C.c_double), ('path_lo', C.c_double), ('path_hi', C.c_double),
                ('path_scale', C.c_double)]


class Extrema(C.Structure):
    _fields_ = [('n_scat_min', C.c_uint32), ('n_scat_max', C.c_uint32), ('path_min', C.c_float),
                ('path_max', C.c_float)]


ROW_DTYPE = np.dtype([('wvl_um', 'f8'), ('ssa_ice', 'f8'), ('ssa_imp', 'f8'), ('g', 'f8'),
                      ('ext_cff_mss', 'f8'), ('p_ext_imp', 'f8')])

RECORD_COLUMNS = (('condition', np.uint8), ('wvl_row', np.int16), ('theta_n', np.float32),
                  ('phi_n', np.float32), ('n_scat', np.uint32), ('path_length', np.float32))

_lib = None


def load_library():
    """dlopen libmc3d.so and declare the prototypes.  Raises Mc3dError when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise Mc3dError('%s not found: build it with `make -C %s` (or __graft_entry__.build()); there is no '
                        'CPU fallback' % (LIB_PATH, os.path.join(_HERE, 'csrc')))
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    lib.mc3d_abi_version.restype = i32
    lib.mc3d_last_error.restype = C.c_char_p
    lib.mc3d_query.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    lib.mc3d_create.argtypes = [vp, vp, i32]
    lib.mc3d_nccl_unique_id.argtypes = [vp]
    lib.mc3d_create_rank.argtypes = [vp, i32, vp, i32, i32]
    lib.mc3d_destroy.argtypes = [vp]
    lib.mc3d_host_alloc.argtypes = [vp, u64]
    lib.mc3d_host_free.argtypes = [vp]
    lib.mc3d_run.argtypes = [vp, vp, vp, i32, u64, u64, u64, vp, vp, vp]
    lib.mc3d_run_async.argtypes = [vp, i32, vp, vp, i32, u64, u64, u64, vp, vp, vp]
    lib.mc3d_wait.argtypes = [vp, i32, vp]
    lib.mc3d_reduce_tally.argtypes = [vp, vp, u64, i32]
    lib.mc3d_replay.argtypes = [vp, vp, u64] + [vp] * 11
    lib.mc3d_set_launch.argtypes = [vp, i32, i32, i32]
    lib.mc3d_write_records_text.restype = C.c_int64
    lib.mc3d_write_records_text.argtypes = [C.c_char_p, i32, u64, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.mc3d_py_repr.argtypes = [C.c_double, C.c_char_p]
    lib.mc3d_set_histograms.argtypes = [vp, vp]
    lib.mc3d_records_layout.argtypes = [u64, vp, vp]
    lib.mc3d_set_input_caching.argtypes = [vp, i32]
    lib.mc3d_get_histograms.argtypes = [vp, i32, vp, vp, vp]
    lib.mc3d_unpack_records.argtypes = [vp, u64, vp, i32]
    lib.mc3d_set_walk_path.argtypes = [vp, i32]
    lib.mc3d_set_tail_kernel.argtypes = [vp, i32]
    lib.mc3d_run_sweep.argtypes = [vp, vp, i32, vp, i32, u64, vp, vp, vp, vp]
    lib.mc3d_run_sweep_async.argtypes = [vp, i32, vp, i32, vp, i32, u64, u64, u64, vp, vp, vp]
    if lib.mc3d_abi_version() != ABI_VERSION:
        raise Mc3dError('libmc3d.so ABI %d != expected %d' % (lib.mc3d_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise Mc3dError('libmc3d error %d: %s' % (rc, load_library().mc3d_last_error().decode('utf-8', 'replace')))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def query(device=0):
    """(n_devices, sm_count, sm_clock_khz, global_mem_bytes, cc_major, cc_minor) of a CUDA device."""
    lib = load_library()
    n, sm, clk, maj, mnr = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    mem = C.c_uint64(0)
    _check(lib.mc3d_query(device, C.byref(n), C.byref(sm), C.byref(clk), C.byref(mem), C.byref(maj), C.byref(mnr)))
    return dict(n_devices=n.value, sm_count=sm.value, sm_clock_khz=clk.value, global_mem_bytes=mem.value,
                cc_major=maj.value, cc_minor=mnr.value)


def device_count():
    lib = load_library()
    n = C.c_int(0)
    lib.mc3d_query(0, C.byref(n), None, None, None, None, None)
    return n.value


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    _check(load_library().mc3d_nccl_unique_id(buf))
    return bytes(buf)


def make_params(theta0_rad, tau_tot, rho_snw, r_lambert, wvl0_um, sigma_um, k_first, lambert_bottom=True,
                lambert_surface=False, n_theta_bins=0, n_phi_bins=0):
    flags = (FLAG_LAMBERT_BOTTOM if lambert_bottom else 0) | (FLAG_LAMBERT_SURFACE if lambert_surface else 0)
    return Params(float(theta0_rad), float(tau_tot), float(rho_snw), float(r_lambert), float(wvl0_um),
                  float(sigma_um), int(k_first), flags, int(n_theta_bins), int(n_phi_bins))


def py_repr(x):
    """CPython's repr(float) computed by the native formatter (for tests)."""
    buf = C.create_string_buffer(40)
    n = load_library().mc3d_py_repr(float(x), buf)
    return buf.raw[:n].decode('ascii')


def write_records_text(path, records, wvn_by_row, snow_depth_by_row, append=True, n_threads=0):
    """Append the reference's '%d %r %r %r %d %r %r' lines for a dict of record columns (native, multithreaded)."""
    cols = {name: np.ascontiguousarray(records[name], dtype=dt) for name, dt in RECORD_COLUMNS}
    wvn = np.ascontiguousarray(wvn_by_row, dtype=np.float64)
    depth = np.ascontiguousarray(snow_depth_by_row, dtype=np.float64)
    n = len(cols['condition'])
    rc = load_library().mc3d_write_records_text(os.fsencode(path), 1 if append else 0, n, _ptr(cols['condition']),
                                                _ptr(cols['wvl_row']), _ptr(cols['theta_n']), _ptr(cols['phi_n']),
                                                _ptr(cols['n_scat']), _ptr(cols['path_length']), _ptr(wvn), _ptr(depth),
                                                len(wvn), int(n_threads))
    if rc < 0:
        raise Mc3dError('mc3d_write_records_text failed with %d for %s' % (rc, path))
    return int(rc)


class PinnedArray(object):
    """numpy view of page-locked host memory from mc3d_host_alloc (so record copy-back runs at PCIe speed)."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        self._ptr = C.c_void_p(0)
        _check(self._lib.mc3d_host_alloc(C.byref(self._ptr), max(1, n * dtype.itemsize)))
        buf = (C.c_uint8 * max(1, n * dtype.itemsize)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            self._lib.mc3d_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def records_layout(n):
    """(byte offsets of the six record columns, total bytes) of the column block for an n-photon call
    (mc3d_records_layout): record arrays laid out like this come back with one device-to-host copy."""
    off = (C.c_uint64 * 6)()
    total = C.c_uint64(0)
    _check(load_library().mc3d_records_layout(int(n), off, C.byref(total)))
    return [int(x) for x in off], int(total.value)


def unpack_records(packed, n=None, n_threads=0):
    """Expand packed 16-byte records (uint32 array, 4 words per photon; mc3d_records.packed) into a dict of columns."""
    packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1)
    n = packed.size // 4 if n is None else int(n)
    cols = {name: np.empty(n, dtype=dt) for name, dt in RECORD_COLUMNS}
    out = Records(*[cols[name].ctypes.data for name, _ in RECORD_COLUMNS], None)
    _check(load_library().mc3d_unpack_records(_ptr(packed), n, C.byref(out), int(n_threads)))
    return cols


class RecordBuffers(object):
    """Pinned destination for the records of up to ``capacity`` photons in the packed 16-byte form (one device-to-host
    copy of 16 B per photon, include/mc3d.h); ``view(n)`` expands the first n into the six columns."""

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self._block = PinnedArray(max(4 * self.capacity, 4), np.uint32)

    def struct_for(self, n):
        """mc3d_records asking for the packed records of an n-photon call."""
        if int(n) > self.capacity:
            raise ValueError('%d photons do not fit in RecordBuffers(%d)' % (n, self.capacity))
        return Records(None, None, None, None, None, None, self._block.array.ctypes.data)

    def packed(self, n):
        """uint32 view (n, 4) of the packed records."""
        return self._block.array[:4 * int(n)].reshape(int(n), 4)

    def view(self, n):
        """The first n records as a dict of (freshly unpacked) numpy columns."""
        return unpack_records(self._block.array, int(n))

    def free(self):
        self._block.free()


class Context(object):
    """A libmc3d context: either one process driving ``devices`` (list of CUDA ordinals), or rank ``rank`` of
    ``world_size`` single-GPU processes sharing ``nccl_id`` (torchrun / mpirun style)."""

    def __init__(self, devices=None, rank=None, world_size=None, nccl_id=None, device=None):
        self._lib = load_library()
        self._ctx = C.c_void_p(0)
        if rank is None:
            if devices is None:
                devices = [0]
            ids = (C.c_int * len(devices))(*devices)
            _check(self._lib.mc3d_create(C.byref(self._ctx), ids, len(devices)))
            self.rank, self.world_size, self.n_devices = 0, 1, len(devices)
        else:
            idbuf = None
            if world_size > 1:
                idbuf = (C.c_uint8 * 128).from_buffer_copy(nccl_id)
            _check(self._lib.mc3d_create_rank(C.byref(self._ctx), int(device if device is not None else 0), idbuf,
                                              int(rank), int(world_size)))
            self.rank, self.world_size, self.n_devices = int(rank), int(world_size), 1
        self._keep = [None] * N_SLOTS

    def close(self):
        if self._ctx is not None and self._ctx.value:
            self._lib.mc3d_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_input_caching(self, enabled=True):
        """Skip (default) or force the upload of inputs identical to the slot's previous call."""
        _check(self._lib.mc3d_set_input_caching(self._ctx, 1 if enabled else 0))

    def set_walk_path(self, path=PATH_AUTO):
        """PATH_AUTO (default), PATH_FUSED (one kernel, short walks) or PATH_PERSISTENT (three kernels, long walks);
        also accepts 'auto' / 'fused' / 'persistent'.  A performance choice: the results are bit-identical."""
        path = {'auto': PATH_AUTO, 'fused': PATH_FUSED, 'persistent': PATH_PERSISTENT}.get(path, path)
        _check(self._lib.mc3d_set_walk_path(self._ctx, int(path)))

    def set_tail_kernel(self, mode=-1):
        """How a persistent-path call ends: 1 = tail kernel, 0 = the walk kernel drains by itself, -1 = automatic
        (mc3d_set_tail_kernel).  Results do not depend on it."""
        _check(self._lib.mc3d_set_tail_kernel(self._ctx, int(mode)))

    def set_launch(self, blocks_per_sm=0, block_threads=0, refill_threshold=0):
        _check(self._lib.mc3d_set_launch(self._ctx, blocks_per_sm, block_threads, refill_threshold))

    # ---- optional n_scat / path-length histograms (post_processing.py:162-223) --------------------------------
    def set_histograms(self, n_scat_bins=0, n_scat_range=(0., 1.), path_bins=0, path_range=(0., 1.), path_scale=100.):
        """Bin n_scat and path_length * path_scale of every later run on the GPU (np.histogram semantics over the
        given ranges).  ``set_histograms()`` with no bins switches it off."""
        if not n_scat_bins and not path_bins:
            self._hist = None
            _check(self._lib.mc3d_set_histograms(self._ctx, None))
            return
        spec = HistSpec(int(n_scat_bins), int(path_bins), float(n_scat_range[0]), float(n_scat_range[1]),
                        float(path_range[0]), float(path_range[1]), float(path_scale))
        _check(self._lib.mc3d_set_histograms(self._ctx, C.byref(spec)))
        self._hist = spec

    def extrema(self, slot=0):
        """(n_scat_min, n_scat_max, path_min [m], path_max [m]) of the last completed call on ``slot``."""
        x = Extrema()
        _check(self._lib.mc3d_get_histograms(self._ctx, slot, None, None, C.byref(x)))
        return int(x.n_scat_min), int(x.n_scat_max), float(x.path_min), float(x.path_max)

    def histograms(self, slot=0):
        """(n_scat counts, path counts) of the last completed call on ``slot`` (uint64 arrays; empty when off)."""
        spec = getattr(self, '_hist', None)
        if spec is None:
            raise Mc3dError('set_histograms() was not called')
        ns = np.zeros(spec.n_scat_bins, np.uint64)
        pl = np.zeros(spec.path_bins, np.uint64)
        _check(self._lib.mc3d_get_histograms(self._ctx, slot, _ptr(ns) if ns.size else None,
                                             _ptr(pl) if pl.size else None, None))
        return ns, pl

    # ---- production mode ------------------------------------------------------------------------------------
    @staticmethod
    def _records_struct(records, n_photon):
        if isinstance(records, RecordBuffers):
            return records.struct_for(n_photon)
        if isinstance(records, np.ndarray):        # packed records into a caller-owned uint32 array
            assert records.dtype == np.uint32 and records.flags.c_contiguous and records.size >= 4 * int(n_photon)
            return Records(None, None, None, None, None, None, records.ctypes.data)
        if records is not None:
            return Records(*[records[name].ctypes.data if records.get(name) is not None else None
                             for name, _ in RECORD_COLUMNS], None)
        return None

    def run_async(self, slot, params, table, seed, photon_begin, n_photon, records=None, tally=None):
        """Enqueue one walk.  ``records``: RecordBuffers, packed uint32 array, dict of numpy columns, or None.
        ``tally``: uint64 array of shape (n_rows, params.tally_width) or None.  Buffers must stay alive until
        ``wait(slot)``."""
        if not 0 <= int(slot) < N_SLOTS:
            raise Mc3dError('slot must be in [0, %d)' % N_SLOTS)
        table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
        rec_struct = self._records_struct(records, n_photon)
        if tally is not None:
            assert tally.dtype == np.uint64 and tally.flags.c_contiguous
            assert tally.size == len(table) * params.tally_width
        self._keep[slot] = (params, table, rec_struct, records, tally)
        _check(self._lib.mc3d_run_async(self._ctx, int(slot), C.byref(params), _ptr(table), len(table), int(seed),
                                        int(photon_begin), int(n_photon),
                                        C.byref(rec_struct) if rec_struct is not None else None, _ptr(tally), None))

    def run_sync(self, params, table, seed, photon_begin, n_photon, records=None, tally=None):
        """One walk, start to finish (mc3d_run: slot 0); same arguments as ``run_async``, returns the stats dict.  A
        synchronous call that finds the context idle ends in the tail kernel (include/mc3d.h: mc3d_set_tail_kernel)."""
        table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
        rec_struct = self._records_struct(records, n_photon)
        if tally is not None:
            assert tally.dtype == np.uint64 and tally.flags.c_contiguous
            assert tally.size == len(table) * params.tally_width
        st = Stats()
        _check(self._lib.mc3d_run(self._ctx, C.byref(params), _ptr(table), len(table), int(seed), int(photon_begin),
                                  int(n_photon), C.byref(rec_struct) if rec_struct is not None else None, _ptr(tally),
                                  C.byref(st)))
        return st.as_dict()

    def run_sweep_async(self, slot, cases, table, seed, records=None, tally=None, case_events=None, range_begin=0,
                        range_count=None):
        """Enqueue a sweep: ``cases`` is a list of (Params, row_begin, n_rows, n_photon) over the concatenated ``table``;
        photon j of case c is photon id (c << 40) + j of the stream ``seed`` (mc3d_run_sweep in include/mc3d.h).
        ``records``: RecordBuffers / packed uint32 array / dict of columns for the photons of
        [range_begin, range_begin + range_count) in case order, or None; ``tally``: uint64 (len(table), tally_width);
        ``case_events``: uint64 (len(cases),) or None.  Finish with ``wait(slot)``."""
        table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
        arr = (SweepCase * len(cases))()
        total = 0
        for k, (prm, row_begin, n_rows, n) in enumerate(cases):
            arr[k].params = prm
            arr[k].row_begin, arr[k].n_rows, arr[k].n_photon = int(row_begin), int(n_rows), int(n)
            total += int(n)
        if range_count is None:
            range_count = total - int(range_begin)
        rec_struct = None
        if isinstance(records, RecordBuffers):
            rec_struct = records.struct_for(range_count)
        elif isinstance(records, np.ndarray):
            assert records.dtype == np.uint32 and records.flags.c_contiguous and records.size >= 4 * int(range_count)
            rec_struct = Records(None, None, None, None, None, None, records.ctypes.data)
        elif records is not None:
            rec_struct = Records(*[records[name].ctypes.data if records.get(name) is not None else None
                                   for name, _ in RECORD_COLUMNS], None)
        if tally is not None:
            assert tally.dtype == np.uint64 and tally.flags.c_contiguous
            assert tally.size == len(table) * cases[0][0].tally_width
        if case_events is not None:
            assert case_events.dtype == np.uint64 and case_events.flags.c_contiguous and case_events.size == len(cases)
        self._keep[slot] = (arr, table, rec_struct, records, tally, case_events)
        _check(self._lib.mc3d_run_sweep_async(self._ctx, int(slot), arr, len(cases), _ptr(table), len(table), int(seed),
                                              int(range_begin), int(range_count),
                                              C.byref(rec_struct) if rec_struct is not None else None, _ptr(tally),
                                              _ptr(case_events)))

    def run_sweep(self, cases, table, seed, records=True, tally=True):
        """Synchronous sweep.  Returns (list of per-case record dicts or None, tally or None, events per case, stats)."""
        table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
        total = sum(int(c[3]) for c in cases)
        packed = bool(records) and all(int(c[2]) <= PACKED_MAX_ROWS for c in cases)
        rec = None
        if packed:
            rec = np.empty(4 * max(total, 1), np.uint32)
        elif records:
            rec = {name: np.empty(total, dtype=dt) for name, dt in RECORD_COLUMNS}
        t = np.zeros((len(table), cases[0][0].tally_width), np.uint64) if tally else None
        ev = np.zeros(len(cases), np.uint64)
        stats = self._run_sweep_sync(cases, table, seed, rec, t, ev)
        per_case = None
        if records:
            if packed and stats['packed_saturated']:
                return self._sweep_columns(cases, table, seed, tally)
            cols = unpack_records(rec, total) if packed else rec
            edges = np.cumsum([0] + [int(c[3]) for c in cases])
            per_case = [{k: v[edges[i]:edges[i + 1]] for k, v in cols.items()} for i in range(len(cases))]
        return per_case, t, ev, stats

    def _run_sweep_sync(self, cases, table, seed, records, tally, case_events):
        """The whole sweep, start to finish (mc3d_run_sweep); returns the stats dict."""
        table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
        arr = (SweepCase * len(cases))()
        total = 0
        for k, (prm, row_begin, n_rows, n) in enumerate(cases):
            arr[k].params = prm
            arr[k].row_begin, arr[k].n_rows, arr[k].n_photon = int(row_begin), int(n_rows), int(n)
            total += int(n)
        rec_struct = self._records_struct(records, total)
        st = Stats()
        _check(self._lib.mc3d_run_sweep(self._ctx, arr, len(cases), _ptr(table), len(table), int(seed),
                                        C.byref(rec_struct) if rec_struct is not None else None, _ptr(tally),
                                        _ptr(case_events), C.byref(st)))
        return st.as_dict()

    def _sweep_columns(self, cases, table, seed, tally):
        """A sweep whose packed records saturated (a walk beyond 2^23 scatterings): fetch the six columns instead."""
        total = sum(int(c[3]) for c in cases)
        rec = {name: np.empty(total, dtype=dt) for name, dt in RECORD_COLUMNS}
        t = np.zeros((len(table), cases[0][0].tally_width), np.uint64) if tally else None
        ev = np.zeros(len(cases), np.uint64)
        stats = self._run_sweep_sync(cases, table, seed, rec, t, ev)
        edges = np.cumsum([0] + [int(c[3]) for c in cases])
        return [{k: v[edges[i]:edges[i + 1]] for k, v in rec.items()} for i in range(len(cases))], t, ev, stats

    def wait(self, slot):
        st = Stats()
        _check(self._lib.mc3d_wait(self._ctx, slot, C.byref(st)))
        self._keep[slot] = None
        return st.as_dict()

    def run(self, params, table, seed, photon_begin, n_photon, records=True, tally=True):
        """Synchronous walk of photon ids [photon_begin, photon_begin + n_photon).
        Returns (records dict or None, tally array or None, stats dict)."""
        table = np.ascontiguousarray(table, dtype=ROW_DTYPE)
        rec = None
        packed = bool(records) and len(table) <= PACKED_MAX_ROWS and records != 'columns'
        if packed:
            rec = np.empty(4 * max(int(n_photon), 1), np.uint32)
        elif records:
            rec = {name: np.empty(n_photon, dtype=dt) for name, dt in RECORD_COLUMNS}
        t = np.zeros((len(table), params.tally_width), np.uint64) if tally else None
        stats = self.run_sync(params, table, seed, photon_begin, n_photon, rec, t)
        if packed:
            if stats['packed_saturated']:      # a walk longer than 2^23 events: fetch the columns instead
                return self.run(params, table, seed, photon_begin, n_photon, records='columns', tally=tally)
            rec = unpack_records(rec, n_photon)
        return rec, t, stats

    def reduce_tally(self, tally, root=0):
        """Sum a uint64 tally over the ranks of a multi-rank context (one ncclReduce); in place on ``root``."""
        assert tally.dtype == np.uint64 and tally.flags.c_contiguous
        _check(self._lib.mc3d_reduce_tally(self._ctx, _ptr(tally), tally.size, int(root)))
        return tally

    # ---- replay mode ----------------------------------------------------------------------------------------
    def replay(self, params, wvl, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp, init_draws, offsets, stream):
        """fp64 walk over the reference's recorded random stream (see mc3d_replay in include/mc3d.h)."""
        f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        wvl, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp = map(f8, (wvl, ssa_ice, ssa_imp, g, ext_cff_mss,
                                                                     p_ext_imp))
        init_draws, stream = f8(init_draws), f8(stream)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(wvl)
        assert len(init_draws) == 3 * n and len(offsets) == n + 1 and len(stream) >= offsets[-1]
        out = {'condition': np.zeros(n, np.int32), 'wvn': np.zeros(n), 'theta_n': np.zeros(n),
               'phi_n': np.zeros(n), 'n_scat': np.zeros(n, np.int64), 'path_length': np.zeros(n),
               'snow_depth': np.zeros(n), 'consumed': np.zeros(n, np.int64)}
        rec = RecordsF64(*[out[k].ctypes.data for k in ('condition', 'wvn', 'theta_n', 'phi_n', 'n_scat',
                                                        'path_length', 'snow_depth', 'consumed')])
        mism = C.c_uint64(0)
        _check(self._lib.mc3d_replay(self._ctx, C.byref(params), n, _ptr(wvl), _ptr(ssa_ice), _ptr(ssa_imp), _ptr(g),
                                     _ptr(ext_cff_mss), _ptr(p_ext_imp), _ptr(init_draws), _ptr(offsets),
                                     _ptr(stream), C.byref(rec), C.byref(mism)))
        out['n_mismatch'] = int(mism.value)
        return out
