#!/usr/bin/env python
"""bench.py -- headline benchmark: MLUPS of D3Q19 BGK (AA pattern, fp32) lid-driven cavity.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Workload (BASELINE.json metric / north_star): examples/ldc_3d at 512^3 per GPU
(`--access_pattern=AA --grid=D3Q19 --model=bgk`).  N>1 (torchrun, one rank per GPU): the
domain is 512 x 512 x (512 N) split along z into N subdomains (weak scaling), ghost layers
exchanged GPU-to-GPU every step.  A "step" is one CollideAndPropagate pass over the whole
lattice (+ halo push).  MLUPS = fluid nodes x steps / time x 1e-6 (subdomain_runner.py:1699).

Printed JSON keys (one line, rank 0): see the task contract; `value` is device-resident
throughput (CUDA events, max over ranks), `e2e` the same metric for a whole job through the
public runner with HOST buffers (pinned host -> device upload of map/rho/v, initialisation,
K steps, device -> host download of rho/v, all inside the timed region), `roofline` the
bandwidth roofline of the collide kernel, `cpu_baseline` the oracle on the host cores.
`--impl reference` times the reference arm: the CPU restatement of the reference kernels
(oracle/, all host threads) -- the reference itself has no CPU path and its GPU path needs
Mako + pycuda.
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


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.check_output(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                    timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': int(float(self.samples[0][1])), 'reasons': reasons, 'samples': len(self.samples)}


def build_oracle_case(size, dtype=np.float32):
    """ldc_3d geometry via the host mirror (no GPU needed) for the CPU arms."""
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.examples.ldc_3d import LDCSim
    from sailfish_b200 import params as P
    ctrl = LBSimulationController(LDCSim)
    args = ['--lat_nx=%d' % size, '--lat_ny=%d' % size, '--lat_nz=%d' % size, '--visc=0.03175',
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
    return p, sd.encoded_map(), base[0], base[1:], sd.num_fluid_nodes


def cpu_sample_size():
    """Edge of the bounded CPU sample: 256^3 (1.5 GB lattice) keeps a many-core host busy, 160^3 is enough
    for a small one; both are far beyond the last-level cache."""
    return 256 if (os.cpu_count() or 1) >= 32 else 160


def cpu_run(size, budget_s, threads=None):
    """Times the oracle (fp32, OpenMP over z) on ldc_3d size^3; returns MLUPS and a description."""
    from oracle import oracle
    oracle.build()
    cores = threads or os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    p, node_map, rho, v, fluid = build_oracle_case(size)
    sim = oracle.Simulation(p, node_map, rho, v, dtype=np.float32)
    sim.step()
    sim.step()                       # warm-up (both AA parities)
    t0 = time.time()
    n = 0
    while True:
        sim.step()
        sim.step()
        n += 2
        if time.time() - t0 > budget_s or n >= 2000:
            break
    dt = time.time() - t0
    return fluid * n / dt * 1e-6, cores, 'ldc_3d D3Q19 BGK AA %d^3, %d steps, %.1f s, oracle fp32 OpenMP' % (size, n, dt)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    size = cpu_sample_size()
    per_step_budget = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    for k in range(args.warmup + args.steps):
        mlups, cores, sample = cpu_run(size, per_step_budget)
        if k >= args.warmup:
            vals.append(mlups)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'MLUPS D3Q19 BGK', 'value': value, 'unit': 'MLUPS',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * per_step_budget, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'examples/ldc_3d D3Q19 BGK, AA access pattern, %d^3 per GPU (global %dx%dx%d, z-split)'
                               % (args.size, args.size, args.size, args.size * args.gpus),
                   'visc': 0.03175,
                   'sample': 'each step = a bounded %d^3 run of the same cavity on the host cores' % size,
                   'note': 'reference has no CPU path (backend_dummy is a no-op) and its GPU path needs Mako+pycuda; '
                           'this is the CPU restatement of its kernels (oracle/), all host threads'},
        'cpu_baseline': {'value': value, 'unit': 'MLUPS', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'MLUPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--size', type=int, default=512, help='cube edge per GPU')
    ap.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no_cpu_baseline', action='store_true')
    ap.add_argument('--no_e2e', action='store_true')
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
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.geo import EqualSubdomainsGeometry3D
    from sailfish_b200.examples.ldc_3d import LDCSim

    S, N = args.size, max(1, ws)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    ctrl = LBSimulationController(LDCSim, EqualSubdomainsGeometry3D)
    t_setup = time.time()
    ctrl.setup(['--lat_nx=%d' % S, '--lat_ny=%d' % S, '--lat_nz=%d' % (S * N), '--subdomains=%d' % N,
                '--visc=0.03175', '--access_pattern=AA', '--grid=D3Q19', '--model=bgk', '--quiet'])
    r = ctrl.runners[0]
    t_setup = time.time() - t_setup
    lib = bk.lib()
    fluid_local = r.num_fluid_nodes
    fluid_total = fluid_local
    if ws > 1:
        t = torch.tensor([fluid_local], dtype=torch.int64, device='cuda')
        dist.all_reduce(t)
        fluid_total = int(t.item())

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctrl.sync()

    # ---- device-resident throughput -----------------------------------------------------
    for _ in range(W):
        r.step(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.launch_count()
    ev0 = r.backend.make_event(r._calc_stream, timing=True)
    for _ in range(K):
        r.step(False)
    ev1 = r.backend.make_event(r._calc_stream, timing=True)
    barrier()
    ms = ev1.time_since(ev0)
    launches = lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if ws > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = fluid_total * K / (ms * 1e-3) * 1e-6

    # ---- roofline of the collide kernel on this rank (whole-lattice launches, single stream) ----
    dev = r._dev
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
            roofline['traffic'] = json.load(open(tr)).get('dram_bytes_per_launch')
        except Exception:
            pass

    # ---- end to end: whole job with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        barrier()
        be = r.backend
        h_map = dev.h_map
        nb_up = h_map.nbytes + dev.h_rho.nbytes + sum(a.nbytes for a in dev.h_v)
        nb_down = dev.h_rho.nbytes + sum(a.nbytes for a in dev.h_v)
        # pinned staging copy of the map (fields are pinned already)
        p_map = be.alloc_async_host_buf(h_map.shape, np.uint32)
        p_map[:] = h_map
        t0 = time.time()
        ee0 = be.make_event(r._calc_stream, timing=True)
        lib.call('slf_memcpy_h2d_async', int(dev.d_map), p_map.ctypes.data, p_map.nbytes, r._calc_stream.handle)
        be.to_buf_async(dev.d_rho, r._calc_stream)
        for b in dev.d_v:
            be.to_buf_async(b, r._calc_stream)
        dev.set_initial_conditions()
        r._sim.iteration = 0
        for k in range(K):
            r.step(k == K - 1)
        be.from_buf_async(dev.d_rho, r._calc_stream)
        for b in dev.d_v:
            be.from_buf_async(b, r._calc_stream)
        ee1 = be.make_event(r._calc_stream, timing=True)
        barrier()
        e_ms = ee1.time_since(ee0)
        wall = time.time() - t0
        if ws > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        finite = bool(np.isfinite(np.asarray(dev.h_rho)[r._spec._nonghost_slice][r._subdomain.fluid_map()]).all())
        e2e = {'value': fluid_total * K / (e_ms * 1e-3) * 1e-6, 'unit': 'MLUPS',
               'h2d_bytes_per_step': int(nb_up // K), 'd2h_bytes_per_step': int(nb_down // K),
               'job': 'pinned host map+rho+v -> device, SetInitialConditions, %d steps, rho+v -> host' % K,
               'ms_job': e_ms, 'wall_ms_job': wall * 1e3, 'result_finite': finite}

    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline:
        mlups, cores, sample = cpu_run(cpu_sample_size(), 12.0)
        cpu = {'value': mlups, 'unit': 'MLUPS', 'cores': cores, 'kind': 'port', 'sample': sample}
    line = {
        'metric': 'MLUPS D3Q19 BGK', 'value': value, 'unit': 'MLUPS', 'n_gpus': N, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'examples/ldc_3d D3Q19 BGK, AA access pattern, %d^3 per GPU (global %dx%dx%d, z-split)'
                               % (S, S, S, S * N),
                   'visc': 0.03175, 'l2': 'lattice (%.1f GB/GPU) is larger than L2, no flush needed'
                                          % (19 * 4 * dev.n / 1e9),
                   'fluid_nodes': fluid_total, 'setup_s': round(t_setup, 1)},
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
