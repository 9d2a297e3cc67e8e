#!/usr/bin/env python
"""bench.py -- headline benchmark: MLUPS of D3Q19 BGK (AA pattern, fp32) lid-driven cavity.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Workload (BASELINE.json metric / north_star): examples/ldc_3d at 512^3 per GPU
(`--access_pattern=AA --grid=D3Q19 --model=bgk`).  N>1 (torchrun, one rank per GPU): the
headline `value` is WEAK scaling -- the domain is 512 x 512 x (512 N) split along z into N
subdomains, ghost layers exchanged GPU-to-GPU every step -- and the line also carries a
`strong` record: BASELINE config 4, the 512^3 cavity itself split N ways.  A "step" is one
CollideAndPropagate pass over the whole lattice (+ halo push).  MLUPS = fluid nodes x steps /
time x 1e-6 (subdomain_runner.py:1699).

Printed JSON keys (one line, rank 0): see the task contract; `value` is device-resident
throughput (CUDA events, max over ranks), `e2e` the same metric for a whole job through the
public runner with HOST buffers (pinned host -> device upload of map/rho/v, initialisation,
K steps, device -> host download of rho/v, all inside the timed region), `roofline` the
bandwidth roofline of the collide kernel, `cpu_baseline` the oracle on the host cores,
`parity` the correctness evidence of THIS run: GPU vs the fp32/fp64 oracle on a 64^3 cavity
and, for N>1, every layer the halo exchange can have influenced compared bit for bit with a
single-subdomain run of the same global lattice (see verify_decomposition).  The process exits
1 -- after printing the line -- when a result is non-finite or a parity check fails.

`--impl reference` times the reference arm: the CPU restatement of the reference kernels
(oracle/, all host threads, pinned, first-touch placement) -- the reference itself has no CPU
path and its GPU path needs Mako + pycuda.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_LUP = {'D3Q19': 152, 'D2Q9': 72, 'D3Q27': 216}   # Q*4 B read + Q*4 B written
VISC = 0.03175                                              # Re = 400 at 256^3 (regtest/ldc_3d.py:37-40)
PARITY_STEPS = 16


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML in-process (a query takes
    well under a millisecond, so even a 60 ms region yields several samples under load); nvidia-smi as fallback."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index=0):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []            # (sm_mhz, [active reasons])
        self.max_mhz = None
        self.source = 'nvml'
        self._stop_evt = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            # torch / CUDA index the visible devices; NVML indexes the physical ones
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:            # noqa: BLE001
            self._h = None
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        nv = self._nv
        mhz = int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:            # noqa: BLE001
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        bits = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
        self.samples.append((mhz, [n for n in self.NAMES if mask & bits[n]]))

    def _sample_smi(self):
        out = subprocess.check_output(
            ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
            timeout=5).decode().strip()
        s = [x.strip() for x in out.split(',')]
        self.max_mhz = int(float(s[1]))
        self.samples.append((int(float(s[0])), [n for i, n in enumerate(self.NAMES) if s[2 + i].lower().startswith('active')]))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:        # noqa: BLE001
                if self._h is not None:          # NVML query failed: fall back to the command-line tool
                    self._h, self.source = None, 'nvidia-smi'
            self._stop_evt.wait(0.004 if self._h is not None else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock query unavailable']}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted(set(r for s in self.samples for r in s[1]))
        return {'sm_mhz': sm[len(sm) // 2], 'sm_min_mhz': sm[0], 'sm_max_mhz': self.max_mhz, 'reasons': reasons,
                'samples': len(self.samples), 'source': self.source}


# ---------------------------------------------------------------------------------------------
# CPU arms (oracle/ is executed only here: cpu_baseline + --impl reference + the parity blocks)
# ---------------------------------------------------------------------------------------------
def build_oracle_case(size, dtype=np.float32):
    """ldc_3d geometry via the host mirror (no GPU needed) for the CPU arms."""
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.examples.ldc_3d import LDCSim
    from sailfish_b200 import params as P
    ctrl = LBSimulationController(LDCSim)
    args = ['--lat_nx=%d' % size, '--lat_ny=%d' % size, '--lat_nz=%d' % size, '--visc=%r' % VISC,
            '--access_pattern=AA', '--quiet']
    cfg = ctrl._parser.parse(args)
    ctrl.config = cfg
    spec = ctrl._make_specs()[0]
    sim = LDCSim(cfg)
    sd = LDCSim.subdomain((cfg.lat_nz, cfg.lat_ny, cfg.lat_nx), spec, sim.grid, cfg)
    sd.reset()
    base = []
    for _ in range(4):
        f = np.full(sd.arr_shape, np.inf, dtype=np.float32)
        f[spec._nonghost_slice] = 0
        base.append(f)
    sim.rho = base[0][spec._nonghost_slice]
    sim.v = [b[spec._nonghost_slice] for b in base[1:]]
    sim.vx, sim.vy, sim.vz = sim.v
    sd.init_fields(sim)
    lat = list(reversed(sd.lat_shape))
    arr = list(reversed(sd.arr_shape))
    p = P.make_params('D3Q19', 'bgk', 'AA', lat, arr, [False] * 3, sd.bits, sd.type_ids_by_name(),
                      sd.node_params, cfg.visc)
    return p, sd.encoded_map(), base[0], base[1:], sd.num_fluid_nodes, sd.fluid_map(), spec._nonghost_slice


def cpu_sample_size(ncpu):
    """Edge of the bounded CPU sample: 256^3 (1.5 GB lattice) keeps a many-core host busy, 160^3 is enough
    for a small one; both are far beyond the last-level cache."""
    return 256 if ncpu >= 32 else 160


class CpuArm(object):
    """The oracle (fp32, OpenMP over (z, y) rows, threads pinned, first touch by the same rows) on a
    bounded ldc_3d sample; one `sample()` = as many AA step pairs as fit in the budget."""

    def __init__(self):
        from oracle import oracle
        oracle.build()
        self.oracle = oracle
        cpus = oracle.HOST_CPUS
        ncpu = len(cpus)
        # Thread count and placement are chosen by measurement: the "CPUs" of a container may be hyper-threads
        # or be capped by a CPU quota, and then fewer, well-placed threads stream faster.  Set explicitly in any
        # case: torchrun exports OMP_NUM_THREADS=1 and another library may already have initialised OpenMP with it.
        case = build_oracle_case(96)
        cands = []
        for t in sorted(set(max(1, ncpu // d) for d in (1, 2, 4, 8)), reverse=True):
            for pin in (True, False):
                cands.append((t, pin))
        best, self.calibration = None, []
        for t, pin in cands:
            sub = cpus[::max(1, ncpu // t)][:t] if pin else cpus
            oracle.omp_config(t, pin=pin, cpus=sub)
            sim = oracle.Simulation(case[0], case[1], case[2], case[3], dtype=np.float32)    # first touch under this placement
            sim.step(); sim.step()
            t0, n = time.time(), 0
            while time.time() - t0 < 0.6:
                sim.step(); sim.step()
                n += 2
            mlups = case[4] * n / (time.time() - t0) * 1e-6
            self.calibration.append({'threads': t, 'pinned': pin, 'mlups': round(mlups, 1)})
            if best is None or mlups > best[0]:
                best = (mlups, t, pin, sub)
            del sim
        _, t, pin, sub = best
        self.pinned, self.cpus = pin, sub
        self.threads = oracle.omp_config(t, pin=pin, cpus=sub)
        self.size = cpu_sample_size(ncpu)
        p, node_map, rho, v, fluid = build_oracle_case(self.size)[:5]
        self.fluid = fluid
        self.sim = oracle.Simulation(p, node_map, rho, v, dtype=np.float32)
        self.sim.step()
        self.sim.step()                       # warm-up (both AA parities)

    def sample(self, budget_s):
        t0 = time.time()
        n = 0
        while True:
            self.sim.step()
            self.sim.step()
            n += 2
            if time.time() - t0 > budget_s or n >= 4000:
                break
        dt = time.time() - t0
        return self.fluid * n / dt * 1e-6, n, dt

    def describe(self, n, dt):
        return ('ldc_3d D3Q19 BGK AA %d^3, %d steps in %.1f s per sample, oracle fp32, %d OpenMP threads (%s; best of the '
                'calibrated thread counts / placements) on a host with %d usable CPUs'
                % (self.size, n, dt, self.threads, 'pinned 1:1' if self.pinned else 'free to move', len(self.oracle.HOST_CPUS)))


def workload_name(size, gpus):
    return ('examples/ldc_3d D3Q19 BGK, AA access pattern, %d^3 per GPU (global %dx%dx%d, z-split)'
            % (size, size, size, size * gpus))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm()
    total = max(1, args.steps + args.warmup)
    budget = args.cpu_budget if args.cpu_budget > 0 else max(1.0, min(20.0, 100.0 / total))
    vals, durs, nsteps = [], [], 0
    for k in range(args.warmup + args.steps):
        mlups, n, dt = arm.sample(budget)
        if k >= args.warmup:
            vals.append(mlups)
            durs.append(dt)
            nsteps = n
    value = float(np.median(vals))
    sample = arm.describe(nsteps, float(np.median(durs)))
    line = {
        'impl': 'reference', 'metric': 'MLUPS D3Q19 BGK', 'value': value, 'unit': 'MLUPS',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.median(durs)), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.size, args.gpus), 'visc': VISC,
                   'sample': 'each step = a bounded %d^3 run of the same cavity on the host cores (value = median over '
                             'the %d timed samples; ms_per_step = median sample duration)' % (arm.size, len(vals)),
                   'note': 'reference has no CPU path (backend_dummy is a no-op) and its GPU path needs Mako+pycuda; '
                           'this is the CPU restatement of its kernels (oracle/), all host threads'},
        'cpu_baseline': {'value': value, 'min': float(np.min(vals)), 'max': float(np.max(vals)), 'unit': 'MLUPS',
                         'cores': arm.threads, 'kind': 'port', 'sample': sample, 'calibration': arm.calibration},
        'e2e': {'value': value, 'unit': 'MLUPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def cpu_baseline_subprocess():
    """The main arm's cpu_baseline runs the reference arm in a fresh process: this one has torch (and its
    OpenMP runtime, configured by torchrun's OMP_NUM_THREADS=1) loaded."""
    env = dict(os.environ)
    for k in ('OMP_NUM_THREADS', 'RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'CUDA_VISIBLE_DEVICES'):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '3',
                              '--warmup', '1', '--cpu_budget', '4'], env=env, capture_output=True, text=True,
                             timeout=600)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)['cpu_baseline']
        return {'error': (out.stderr or 'no output')[-300:]}
    except Exception as e:                       # noqa: BLE001
        return {'error': repr(e)[:300]}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def make_ctrl(size, nz_global, nsub, local=False, geo=None):
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.geo import EqualSubdomainsGeometry3D
    from sailfish_b200.examples.ldc_3d import LDCSim
    ctrl = LBSimulationController(LDCSim, geo or EqualSubdomainsGeometry3D)
    ctrl.setup(['--lat_nx=%d' % size, '--lat_ny=%d' % size, '--lat_nz=%d' % nz_global, '--subdomains=%d' % nsub,
                '--visc=%r' % VISC, '--access_pattern=AA', '--grid=D3Q19', '--model=bgk', '--quiet'], local=local)
    return ctrl


def window_geometry(zlo, zhi):
    """A single subdomain covering the global layers [zlo, zhi) of the same global geometry."""
    from sailfish_b200.geo import LBGeometry3D
    from sailfish_b200.subdomain import SubdomainSpec

    class WindowGeometry(LBGeometry3D):
        @classmethod
        def add_options(cls, group):
            LBGeometry3D.add_options(group)
            group.add_argument('--subdomains', type=int, default=1)

        def subdomains(self, n=None):
            return [SubdomainSpec((0, 0, zlo), (self.gx, self.gy, zhi - zlo), id_=0)]
    return WindowGeometry


def host_fields(r):
    """[rho, vx, vy, vz] over the real nodes of the subdomain + the wet-node mask."""
    sl = r._spec._nonghost_slice
    f = np.array([np.array(r.field_base(r._sim.rho))[sl]] + [np.array(r.field_base(c))[sl] for c in r._sim.v])
    return f, r._subdomain.fluid_map()


def verify_decomposition(ctrl, size, nz_global, dist, steps=PARITY_STEPS):
    """Decomposition invariance (regtest/subdomains/3d_ldc.py, util.py:27-30 -- there to 6 decimals, here
    bit for bit).  The linked run is re-initialised and stepped `steps` times; each rank then recomputes, as
    ONE subdomain without neighbours on its own GPU, a window of the global lattice around each of its
    interfaces and compares its own layers with it.  Garbage enters a window through its cut faces at one
    layer per step, so windows extend steps + 4 layers beyond the compared layers; halo data likewise
    travels one layer per step, so comparing steps + 8 layers on this side of every interface covers
    everything the exchange can have influenced."""
    import torch
    r = ctrl.runners[0]
    ctrl.reinitialize(reset_fields=True)      # the host fields hold the results of the previous job
    ctrl.run_steps(steps, save_last=True)
    ctrl.sync()
    r._fields_to_host()
    mine, wet = host_fields(r)
    z0 = r._spec.location[2]
    nz = r._spec.size[2]
    z1 = z0 + nz
    m, c = steps + 4, steps + 8
    ranges = []                                    # global layer ranges [a, b) of this rank to check
    if 4 in r._spec._face_conn:
        ranges.append((z0, min(z1, z0 + c)))
    if 5 in r._spec._face_conn:
        ranges.append((max(z0, z1 - c), z1))
    if len(ranges) == 2 and ranges[0][1] >= ranges[1][0]:
        ranges = [(z0, z1)]
    bad, maxdiff, layers, finite = 0, 0.0, 0, True
    for a, b in ranges:
        zlo, zhi = max(0, a - m), min(nz_global, b + m)
        w = make_ctrl(size, nz_global, 1, local=True, geo=window_geometry(zlo, zhi))
        wr = w.runners[0]
        wr.run_steps(steps, save_last=True)
        w.sync()
        wr._fields_to_host()
        ref, _ = host_fields(wr)
        mask = wet[a - z0:b - z0]
        x = mine[:, a - z0:b - z0][:, mask]
        y = ref[:, a - zlo:b - zlo][:, mask]
        finite = finite and bool(np.isfinite(x).all()) and bool(np.isfinite(y).all())
        neq = x != y
        bad += int(neq.sum())
        if neq.any():
            maxdiff = max(maxdiff, float(np.nanmax(np.abs(x[neq] - y[neq]))))
        layers += b - a
        del w, wr
    finite = finite and bool(np.isfinite(mine[:, wet]).all())
    t = torch.tensor([bad, layers, 0 if finite else 1], dtype=torch.int64, device='cuda')
    dist.all_reduce(t)
    d = torch.tensor([maxdiff], dtype=torch.float64, device='cuda')
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    bad, layers, nonfinite = (int(v) for v in t.tolist())
    return {'steps': steps, 'reference': 'single-subdomain run of the same global lattice (windows around every interface)',
            'layers_compared': layers, 'values_differing': bad, 'max_abs_diff': float(d.item()),
            'bit_identical': bad == 0, 'result_finite': nonfinite == 0, 'ok': bad == 0 and nonfinite == 0}


def verify_oracle(size=64, steps=20):
    """GPU (public runner) vs the fp32 and fp64 oracles on ldc_3d size^3: the fp32 oracle's own deviation
    from exact arithmetic is printed beside the GPU's (relative max norm over wet nodes)."""
    from oracle import oracle
    oracle.build()
    p, node_map, rho, v, _, fluid_map, sl = build_oracle_case(size)
    ctrl = make_ctrl(size, size, 1, local=True)
    r = ctrl.runners[0]
    r.run_steps(steps, save_last=True)
    ctrl.sync()
    r._fields_to_host()
    g, wet = host_fields(r)
    out = {}
    res = {}
    for name, dt in (('f32', np.float32), ('f64', np.float64)):
        sim = oracle.Simulation(p, node_map, rho, v, dtype=dt)
        sim.run(steps)
        res[name] = np.array([sim.rho.reshape(node_map.shape)[sl]] + [c.reshape(node_map.shape)[sl] for c in sim.v],
                             dtype=np.float64)

    def rel(a, b, k):
        a, b = (a[0][wet], b[0][wet]) if k == 'rho' else (a[1:][:, wet], b[1:][:, wet])
        return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    g64 = g.astype(np.float64)
    for k in ('rho', 'v'):
        out['gpu_vs_f64_' + k] = rel(g64, res['f64'], k)
        out['f32oracle_vs_f64_' + k] = rel(res['f32'], res['f64'], k)
        out['gpu_vs_f32oracle_' + k] = rel(g64, res['f32'], k)
    out['config'] = 'ldc_3d D3Q19 BGK AA %d^3, %d steps, relative max norm over wet nodes' % (size, steps)
    # the reference computes in fp32: the GPU must be as close to exact arithmetic as the fp32 restatement is
    out['ok'] = bool(out['gpu_vs_f64_rho'] <= max(1e-6, 2 * out['f32oracle_vs_f64_rho']) and
                     out['gpu_vs_f64_v'] <= max(1e-6, 2 * out['f32oracle_vs_f64_v']))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--size', type=int, default=512, help='cube edge per GPU')
    ap.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu_budget', type=float, default=0.0, help='seconds per CPU sample (reference arm)')
    ap.add_argument('--no_cpu_baseline', action='store_true')
    ap.add_argument('--no_e2e', action='store_true')
    ap.add_argument('--no_strong', action='store_true')
    ap.add_argument('--no_parity', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

    ws = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if ws != args.gpus:
        if ws == 1 and args.gpus > 1:
            sys.stderr.write('bench.py: --gpus %d needs torchrun --nproc-per-node %d\n' % (args.gpus, args.gpus))
            sys.exit(2)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if ws > 1:
        dist.init_process_group('cpu:gloo,cuda:nccl', device_id=torch.device('cuda', local_rank))

    from sailfish_b200 import backend_b200 as bk

    S, N = args.size, max(1, ws)
    W = max(4, args.warmup + (args.warmup & 1))      # even: the timed region starts on the parity the graph was recorded on
    K = max(1, args.steps)
    lib = bk.lib()
    failed = []

    def allmax(x):
        if ws > 1:
            t = torch.tensor([x], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def allsum(x):
        if ws > 1:
            t = torch.tensor([x], dtype=torch.int64, device='cuda')
            dist.all_reduce(t)
            return int(t.item())
        return x

    def timed(ctrl, warm, steps, sampler_box=None):
        """`warm` untimed + `steps` timed steps through the public runner (graph replay of step pairs);
        device time on the launching stream, max over ranks."""
        r = ctrl.runners[0]

        def barrier():
            ctrl.sync()                      # own streams idle first, so that every rank leaves the rendezvous together
            if ws > 1 and not ctrl._local:
                dist.barrier()
            torch.cuda.synchronize()
        if sampler_box is not None and rank == 0:
            # started BEFORE the rendezvous: initialising NVML takes milliseconds, and a rank that enters the timed
            # region late makes its neighbours wait for its ghost layers inside theirs
            sampler_box.append(ClockSampler(local_rank))
            sampler_box[0].start()
        ctrl.run_steps(warm)
        barrier()
        l0 = lib.launch_count()
        ev0 = r.backend.make_event(r._calc_stream, timing=True)
        ctrl.run_steps(steps)
        ev1 = r.backend.make_event(r._calc_stream, timing=True)
        barrier()
        ms = ev1.time_since(ev0)
        return (allmax(ms) if not ctrl._local else ms), lib.launch_count() - l0

    # ---- weak scaling (headline): 512^3 per GPU ------------------------------------------------------
    t_setup = time.time()
    ctrl = make_ctrl(S, S * N, N)
    t_setup = time.time() - t_setup
    r = ctrl.runners[0]
    dev = r._dev
    fluid_local = r.num_fluid_nodes
    fluid_total = allsum(fluid_local)
    box = []
    ms, launches = timed(ctrl, W, K, box)
    value = fluid_total * K / (ms * 1e-3) * 1e-6

    # ---- roofline of the collide kernel on this rank (whole-lattice launches, single stream) ----
    ctrl.quiesce()
    for _ in range(4):
        dev.collide()
        dev.iteration += 1
    ctrl.sync()
    e0 = r.backend.make_event(r._calc_stream, timing=True)
    reps = 20
    for _ in range(reps):
        dev.collide()
        dev.iteration += 1
    e1 = r.backend.make_event(r._calc_stream, timing=True)
    ctrl.sync()
    clocks = box[0].stop() if box else None
    kms = e1.time_since(e0) / reps
    peak, peak_src = measured_peak()
    alg_bytes = BYTES_PER_LUP['D3Q19'] * fluid_local
    achieved = alg_bytes / (kms * 1e-3) * 1e-9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'kernel': 'collide_bulk_aa_tab<D3Q19,BGK> (+collide_boundary)', 'ms_per_launch': kms,
                'algorithmic_bytes_per_launch': alg_bytes, 'peak_source': peak_src}
    tr = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roofline['traffic'] = tj.get('dram_bytes_per_launch')
            roofline['traffic_source'] = tj.get('source', 'profiles/traffic.json (ncu --set full capture of this kernel)')
        except Exception:
            pass

    # ---- end to end: whole job with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        be = r.backend
        h_map = dev.h_map
        nb_up = h_map.nbytes + dev.h_rho.nbytes + sum(a.nbytes for a in dev.h_v)
        nb_down = dev.h_rho.nbytes + sum(a.nbytes for a in dev.h_v)
        p_map = be.alloc_async_host_buf(h_map.shape, np.uint32)     # pinned staging copy of the map
        p_map[:] = h_map
        ctrl.quiesce()                      # everybody idle: the neighbours' pushes of the last step have landed
        t0 = time.time()
        ee0 = be.make_event(r._calc_stream, timing=True)
        lib.call('slf_memcpy_h2d_async', int(dev.d_map), p_map.ctypes.data, p_map.nbytes, r._calc_stream.handle)
        r.reinitialize(upload=True)         # rho, v: pinned host -> device, SetInitialConditions
        if ws > 1:
            ctrl.quiesce()                  # nobody pushes into a lattice that is still being initialised
        ctrl.run_steps(K, save_last=True)
        be.from_buf_async(dev.d_rho, r._calc_stream)
        for b in dev.d_v:
            be.from_buf_async(b, r._calc_stream)
        ee1 = be.make_event(r._calc_stream, timing=True)
        ctrl.quiesce()
        e_ms = allmax(ee1.time_since(ee0))
        wall = time.time() - t0
        f, wet = host_fields(r)
        finite = allsum(0 if bool(np.isfinite(f[:, wet]).all()) else 1) == 0
        if not finite:
            failed.append('e2e result non-finite')
        e2e = {'value': fluid_total * K / (e_ms * 1e-3) * 1e-6, 'unit': 'MLUPS',
               'h2d_bytes_per_step': int(nb_up // K), 'd2h_bytes_per_step': int(nb_down // K),
               'bytes_note': 'whole-job copies (%d B up, %d B down per GPU) divided by the %d steps' % (nb_up, nb_down, K),
               'job': 'pinned host map+rho+v -> device, SetInitialConditions, %d steps, rho+v -> host' % K,
               'ms_job': e_ms, 'wall_ms_job': wall * 1e3, 'result_finite': finite}

    # ---- correctness of THIS run ---------------------------------------------------------------------
    parity = {}
    if not args.no_parity:
        if ws > 1:
            parity['decomposition'] = verify_decomposition(ctrl, S, S * N, dist)
            if not parity['decomposition']['ok']:
                failed.append('weak-scaling run differs from the single-subdomain run')
        if rank == 0:
            parity['oracle'] = verify_oracle()
            if not parity['oracle']['ok']:
                failed.append('GPU vs oracle out of tolerance')

    # ---- strong scaling: BASELINE config 4, the 512^3 cavity split N ways --------------------------------
    strong = None
    if ws > 1 and not args.no_strong:
        sctrl = make_ctrl(S, S, N)
        s_fluid = allsum(sctrl.runners[0].num_fluid_nodes)
        s_ms, _ = timed(sctrl, W, K)
        strong = {'value': s_fluid * K / (s_ms * 1e-3) * 1e-6, 'unit': 'MLUPS', 'ms_per_step': s_ms / K,
                  'workload': 'examples/ldc_3d D3Q19 BGK AA, global %d^3 split into %d z-slabs of %d layers'
                              % (S, N, S // N), 'fluid_nodes': s_fluid, 'scaling': 'strong'}
        if not args.no_parity:
            strong['parity'] = verify_decomposition(sctrl, S, S, dist)
            if not strong['parity']['ok']:
                failed.append('strong-scaling run differs from the single-subdomain run')
        # the same lattice on ONE GPU, measured in this job on rank 0 (the other ranks wait)
        if rank == 0:
            one = make_ctrl(S, S, 1, local=True)
            o_ms, _ = timed(one, W, K)
            o_val = one.runners[0].num_fluid_nodes * K / (o_ms * 1e-3) * 1e-6
            strong['one_gpu_value'] = o_val
            strong['speedup_vs_one_gpu'] = strong['value'] / o_val
        dist.barrier()

    if rank != 0:
        if ws > 1:
            dist.barrier()
        sys.exit(1 if failed else 0)
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess()
    line = {
        'metric': 'MLUPS D3Q19 BGK', 'value': value, 'unit': 'MLUPS', 'n_gpus': N, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(S, N), 'visc': VISC,
                   'l2': 'lattice (%.1f GB/GPU) is larger than L2, no flush needed' % (19 * 4 * dev.n / 1e9),
                   'fluid_nodes': fluid_total, 'setup_s': round(t_setup, 1),
                   'stepping': 'CUDA-graph replay of step pairs (N=1: slf_run_steps; N>1: recorded from SubdomainRunner.step)'},
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
        'parity': parity, 'strong': strong,
    }
    if failed:
        line['failed'] = failed
    print(json.dumps(line))
    sys.stdout.flush()
    if ws > 1:
        dist.barrier()
    sys.exit(1 if failed else 0)


if __name__ == '__main__':
    main()
