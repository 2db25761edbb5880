#!/usr/bin/env python
"""Benchmark of the photon random walk (BASELINE.json metric: photon packets/s and scatter events/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--photons P] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port PORT \
           bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: P photon packets (default 10^6 = BASELINE.json configs[1]:
spheres, HG, lambda0 1.3 um, FWHM 0.085 um, r_eff 100 um, theta0 15 deg, tau_tot 10^6, Lambertian bottom R 0.5) walked
to completion by libmc3d.so on every rank (weak scaling: each GPU gets its own P photon ids per step).  Synthetic
Mie table (monte_carlompi_b200/ssp_fixtures.py, family 'spectral'); the reference's tarball is not redistributable.

Timed regions (each bracketed by barrier + torch.cuda.synchronize(), max over ranks):
  value   K steps, outputs stay on the GPU except the 61 KB tally block per step; up to four steps in flight on separate
          streams so the long-walk tail of one step overlaps the start of the next; one NCCL reduce of the summed
          tallies at the end when N > 1 (the path's only collective).
  e2e     the same K steps through the public C-ABI call with HOST buffers: SSP table uploaded and all per-photon
          record columns copied back to pinned host memory inside the timed region.
  isolated  a few single steps, one at a time: CUDA-event duration of walk + finalize kernels -> roofline per launch.
`--impl reference` times the CPU restatement of the reference's path (oracle/, all host threads) on the same workload.
PyTorch is used here for the process group, barriers and device synchronisation only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(wvl0=1.3, half_width=0.085, rds_snw=100, theta_0=15.0, tau_tot=1e6, rho_snw=300.0,
                lambert_bottom=True, r_lambert=0.5, n_theta_bins=137, fixture='spectral', seed=20190603)
W_EVENT = 111.0   # algorithmic lane-instructions per scattering event
NCU_TRAFFIC_BYTES_PER_PHOTON = 47.8   # measured once with ncu (profiles/), see roofline.traffic_is
LOOP_CEILING_EVENTS_PER_S = 2.11e11   # tools/microbench/hotloop.cu: the event loop alone, all lanes busy, no refill (profiles/)


def build_table():
    """The per-wavelength SSP table exactly as MonteCarlo.run builds it (ssp.py), from the synthetic NetCDF files."""
    import tempfile
    from monte_carlompi_b200 import ssp, ssp_fixtures
    optics = os.path.join(tempfile.mkdtemp(prefix='mc3d_bench_optics_'), 'spectral')
    ssp_fixtures.write_optics_dir(optics, WORKLOAD['fixture'], (WORKLOAD['rds_snw'],))
    scale = WORKLOAD['half_width'] / 2.355
    k_lo, k_hi = ssp.wavelength_grid(WORKLOAD['wvl0'], scale)
    rows = ssp.build_table(optics, 'mie_sot_ChC90_dns_1317.nc', WORKLOAD['rds_snw'], k_lo, k_hi, 0.0)
    return rows, k_lo, scale


def workload_config(photons, extra=None):
    cfg = {'workload': 'configs[1]: spheres, HG, n_photon=%d per step per GPU, wvl0 1.3 um, FWHM 0.085 um, r_eff 100 um, '
                       'theta0 15 deg, tau_tot 1e6, Lambertian bottom R=0.5' % photons,
           'photons_per_step_per_gpu': photons, 'ssp_table': 'synthetic spectral fixture (ssp_fixtures.py)',
           'l2': 'not applicable: compute-bound walk, input is a 2.6 KB table staged in shared memory; every step '
                 'streams 32 B/photon raw + 19 B/photon records (51 MB at 1e6) through rotating buffers'}
    cfg.update(extra or {})
    return cfg


class ClockSampler(threading.Thread):
    """NVML clocks + throttle reasons of one GPU, sampled every 20 ms while the timed regions run."""
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.sm_max, self._stop_evt, self.active = index, [], set(), None, threading.Event(), False

    def _sample(self):
        import pynvml
        self.samples.append(pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self.ready = True
            while not self._stop_evt.is_set():
                if self.active:
                    self._sample()
                time.sleep(0.02)
        except Exception as e:   # NVML missing: report that, never fake a clock
            self.error = repr(e)

    def sample_now(self):
        """One sample from the calling thread while work is in flight (short timed regions may end between two
        periodic samples)."""
        try:
            if getattr(self, 'ready', False) and self.active:
                self._sample()
        except Exception as e:
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(2.0)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': sorted(self.reasons),
                    'note': getattr(self, 'error', 'no sample fell inside the timed regions')}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.sm_max, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's CPU affinity to the NUMA node its GPU hangs off, before any pinned host memory is allocated,
    so that record copy-back does not cross the socket interconnect.  Returns the node or None (no-op on failure)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus[-12:].lower()).read())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_port_rate(rows, k_lo, scale, photons, n_threads, begin=0):
    """Time the oracle's production-mode restatement (fp64, same Philox draws) on `photons` photon packets."""
    from oracle import oracle
    P = oracle.make_params(np.pi * WORKLOAD['theta_0'] / 180., WORKLOAD['tau_tot'], WORKLOAD['rho_snw'], WORKLOAD['r_lambert'],
                           WORKLOAD['wvl0'], scale, k_lo, lambert_bottom=WORKLOAD['lambert_bottom'],
                           n_theta_bins=WORKLOAD['n_theta_bins'])
    t0 = time.perf_counter()
    o = oracle.philox(P, rows, WORKLOAD['seed'], begin, photons, n_threads=n_threads, records=False)
    dt = time.perf_counter() - t0
    return photons / dt, o['n_events'] / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python (nothing to
    compile into oracle/_ref) and does not exist on the GPU box, so its restatement oracle/mc3d_oracle.c ("port") is
    timed with every host thread; the unmodified Python source runs ~350x slower per core."""
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    rows, k_lo, scale = build_table()
    cores = os.cpu_count()
    rate, _, _ = cpu_port_rate(rows, k_lo, scale, 100000, cores)                     # size the per-step sample
    budget_s = 90.0 / max(1, args.steps + args.warmup)
    sample = int(max(20000, min(args.photons, rate * budget_s)))
    for w in range(args.warmup):
        cpu_port_rate(rows, k_lo, scale, sample, cores, begin=w * sample)
    t0 = time.perf_counter()
    events = 0
    for s in range(args.steps):
        _, ev_rate, dt = cpu_port_rate(rows, k_lo, scale, sample, cores, begin=(args.warmup + s) * sample)
        events += ev_rate * dt
    T = time.perf_counter() - t0
    value = args.steps * sample / T
    line = {'impl': 'reference', 'metric': 'photon_packets_per_s', 'value': value, 'unit': 'photons/s',
            'events_per_s': events / T, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * T / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.photons, {'sample': '%d photon packets per step (bounded sample of the %d-photon '
                                                               'workload), %d threads' % (sample, args.photons, cores)}),
            'cpu_baseline': {'value': value, 'unit': 'photons/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x %d photon packets, oracle/mc3d_oracle.c (fp64 restatement of '
                                       'monte_carlo3D.py:1111-1490), pthreads' % (args.steps, sample)},
            'e2e': {'value': value, 'unit': 'photons/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--photons', type=int, default=1000000, help='photon packets per step per GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--inflight', type=int, default=16, help='steps in flight (1..16 library slots)')
    ap.add_argument('--launch', default='', help='blocks_per_sm,block_threads,refill_threshold (tuning)')
    args = ap.parse_args()
    if args.impl == 'ours':
        # at least 3 warm-up steps, and at least one per slot in flight: a slot allocates its device buffers on first use
        args.warmup = max(args.warmup, 3, min(16, args.inflight))
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    if world != args.gpus:
        raise SystemExit('bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N > 1)' % (args.gpus, world))

    import torch
    import torch.distributed as dist
    from monte_carlompi_b200 import engine
    if not torch.cuda.is_available() or engine.device_count() == 0:
        raise SystemExit('bench.py: no CUDA device; the walk has no CPU fallback')
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    nccl_id = None
    if world > 1:
        dist.init_process_group('cpu:gloo,cuda:nccl', rank=rank, world_size=world)
        box = [engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    ctx = engine.Context(rank=rank, world_size=world, nccl_id=nccl_id, device=local_rank)
    if args.launch:
        ctx.set_launch(*[int(x) for x in args.launch.split(',')])
    depth = max(1, min(engine.N_SLOTS, args.inflight))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    rows, k_lo, scale = build_table()
    P = engine.make_params(np.pi * WORKLOAD['theta_0'] / 180., WORKLOAD['tau_tot'], WORKLOAD['rho_snw'], WORKLOAD['r_lambert'],
                           WORKLOAD['wvl0'], scale, k_lo, lambert_bottom=WORKLOAD['lambert_bottom'],
                           n_theta_bins=WORKLOAD['n_theta_bins'])
    n = args.photons
    n_rows = len(rows)
    tallies = [np.zeros((n_rows, P.tally_width), np.uint64) for _ in range(depth)]
    bufs = [engine.RecordBuffers(n) for _ in range(depth)]
    seed = WORKLOAD['seed']

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def photon_begin(step):          # every (step, rank) walks its own id range: the whole job is distinct photons
        return (step * world + rank) * n

    def pipeline(n_steps, first_step, with_records):
        """n_steps steps, `depth` in flight (library slots, each on its own stream).  Returns (events, summed tally)."""
        total = np.zeros_like(tallies[0])
        events = 0
        for i in range(n_steps):
            slot = i % depth
            if i >= depth:
                events += ctx.wait(slot)['n_events']
                total += tallies[slot]
            ctx.run_async(slot, P, rows, seed, photon_begin(first_step + i), n, bufs[slot] if with_records else None,
                          tallies[slot])
        sampler.sample_now()                         # the last `depth` steps are still running on the GPU here
        for i in range(max(0, n_steps - depth), n_steps):
            events += ctx.wait(i % depth)['n_events']
            total += tallies[i % depth]
        if world > 1:
            ctx.reduce_tally(total, root=0)          # the path's single collective: ncclReduce(sum, uint64)
        return events, total

    def timed(n_steps, first_step, with_records):
        """Barrier + synchronize, CUDA events around exactly n_steps steps, barrier + synchronize; max over ranks.
        Both events are recorded while every library stream is idle (before the first submission / after the last
        mc3d_wait returned), so their device timestamps bracket all of the steps' GPU work and copies."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = True
        t0 = time.perf_counter()
        e0.record()
        events, total = pipeline(n_steps, first_step, with_records)
        checksum = 0
        if with_records:                                  # read the step's result on the host
            checksum = int(bufs[0].view(n)['n_scat'][:16].sum()) + int(total[:, 1].sum())
        e1.record()
        e1.synchronize()
        t_dev, t_host = e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0
        barrier()
        sampler.active = False
        if os.environ.get('MC3D_BENCH_DEBUG'):
            sys.stderr.write('rank %d: %.4f ms/step (CUDA events), %.4f (host clock)\n'
                             % (rank, 1e3 * t_dev / n_steps, 1e3 * t_host / n_steps))
        return max_over_ranks(t_dev), max_over_ranks(t_host), events, total, checksum

    # ---- device-resident throughput ("value")
    pipeline(args.warmup, 0, False)
    T_a, T_a_host, events_a, total_a, _ = timed(args.steps, args.warmup, False)

    # ---- end to end through the C ABI with host buffers ("e2e"): every step uploads its inputs (SSP table, bin edges)
    # from pinned host memory again and copies its records and tallies back
    ctx.set_input_caching(False)
    pipeline(args.warmup, args.warmup + args.steps, True)
    T_e, T_e_host, events_e, total_e, checksum = timed(args.steps, 2 * args.warmup + args.steps, True)

    ctx.set_input_caching(True)
    # ---- isolated launches: CUDA-event time of walk + finalize per step (no overlap), for the roofline
    iso_ms, iso_events = [], []
    for i in range(min(10, max(3, args.steps))):
        _, _, st = ctx.run(P, rows, seed, photon_begin(3 * (args.warmup + args.steps) + i), n, records=False, tally=True)
        iso_ms.append(st['kernel_ms'])
        iso_events.append(st['n_events'])
    sampler.active = False
    stats = st
    events_total = sum_over_ranks(float(events_a))
    if rank == 0:
        sampler.stop()
    clocks = sampler.summary() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1:
        from oracle import oracle
        oracle.build()
        cores = os.cpu_count()
        r0, _, _ = cpu_port_rate(rows, k_lo, scale, 100000, cores)
        sample = int(max(100000, min(20 * n, r0 * 12.0)))
        pr, er, dt = cpu_port_rate(rows, k_lo, scale, sample, cores, begin=10 ** 12)
        cpu = {'value': pr, 'unit': 'photons/s', 'events_per_s': er, 'cores': cores, 'kind': 'port',
               'sample': '%d photon packets of the same workload in %.1f s, oracle/mc3d_oracle.c (fp64 restatement of the '
                         "reference's walk), pthreads on all host cores; the unmodified Python reference measured "
                         '~2.35e4 events/s/core in the build container (BASELINE.md)' % (sample, dt)}

    if rank == 0:
        f_mhz = clocks['sm_mhz'] or (stats['sm_clock_khz'] / 1e3)
        peak = stats['sm_count'] * 128 * f_mhz * 1e6 / W_EVENT
        ach = events_a / T_a                                   # per GPU, pipelined steady state of this workload
        iso = np.mean(iso_events) / (np.mean(iso_ms) * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        alg_bytes = 32.0 * 2 + 19.0                            # raw written + raw read + records, per photon
        line = {
            'metric': 'photon_packets_per_s', 'value': world * args.steps * n / T_a, 'unit': 'photons/s',
            'events_per_s': events_total / T_a, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * T_a / args.steps, 'host_clock_ms_per_step': 1e3 * T_a_host / args.steps,
            'timing': 'CUDA events recorded with all library streams idle, around exactly `steps` steps, barrier + synchronize on both '
                      'sides, max over ranks; host_clock_ms_per_step is the perf_counter cross-check of the same region',
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(n, {'events_per_photon': events_a / float(args.steps * n), 'grid_blocks': stats['grid_blocks'],
                                          'block_threads': stats['block_threads'], 'steps_in_flight': depth, 'rank0_numa_node': numa_node}),
            'clocks': clocks,
            'e2e': {'value': world * args.steps * n / T_e, 'unit': 'photons/s', 'events_per_s': sum_over_ranks(float(events_e)) / T_e
                    if world == 1 else None, 'ms_per_step': 1e3 * T_e / args.steps,
                    'h2d_bytes_per_step': int(64 * n_rows + 8 * (P.n_theta_bins + 1 + max(1, P.n_phi_bins) + 1)),   # DevRow table + bin edges
                    'd2h_bytes_per_step': int(19 * n + tallies[0].nbytes + 8), 'host_checksum': checksum},
            'gpu_launches': 3 * args.steps * world,   # init + walk + finalize kernels per step and rank
            'roofline': {'bound': 'issue', 'unit': 'events/s', 'achieved': ach, 'peak': peak, 'frac': ach / peak,
                         'peak_is': 'N_SM x 128 lanes x f_SM / 111 lane-instructions per event; N_SM=%d queried, f_SM=%.0f MHz '
                                    '%s' % (stats['sm_count'], f_mhz, 'median NVML sample under load' if clocks['sm_mhz'] else 'cudaDevAttrClockRate'),
                         'achieved_is': 'events per step / (timed region / steps), %d steps in flight' % depth,
                         'isolated_launch_ms': float(np.mean(iso_ms)), 'isolated_achieved': iso, 'isolated_frac': iso / peak,
                         'measured_loop_ceiling': LOOP_CEILING_EVENTS_PER_S, 'frac_of_measured_loop_ceiling': ach / LOOP_CEILING_EVENTS_PER_S,
                         'measured_loop_ceiling_is': 'events/s of the same event loop run alone on one B200 (no termination, no refill, 32/32 '
                                                     'lanes, 8 warps per scheduler): half-rate ALU / IMAD.WIDE instructions cost two issue '
                                                     'cycles, so 111 lane-instructions cost ~176 cycles (profiles/r01_microbench_hotloop.log)',
                         'traffic': NCU_TRAFFIC_BYTES_PER_PHOTON * n,
                         'traffic_is': 'dram__bytes_read.sum + dram__bytes_write.sum of the walk kernel from the ncu --set full '
                                       'capture at 1e6 photons per launch (profiles/r01_walk_bench_1e6_ncu_summary.csv: 31.8 MB read + 16.0 MB written, incl. the 16 MB fresh list), '
                                       'scaled per photon; algorithmic bytes of the launch = 16 B/photon fresh-list read + 32 B/photon raw record',
                         'hbm': {'achieved': alg_bytes * n / (np.mean(iso_ms) * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                 'frac': alg_bytes * n / (np.mean(iso_ms) * 1e-3) / 1e9 / hbm_peak,
                                 'peak_is': 'hbm_gbs of MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s'}},
            'cpu_baseline': cpu,
        }
        if world > 1:
            line['e2e'].pop('events_per_s')
        print(json.dumps(line))
    for b in bufs:
        b.free()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
