#!/usr/bin/env python
"""bench.py -- headline benchmark of the bhmm HMM hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference impl_c on the host cores

Metric (BASELINE.json): frames x iterations per second of Baum-Welch EM.  Workload (config.workload): "c3" =
10-state dalton-style Gaussian HMM, 1024 trajectories x 1e5 frames PER GPU (weak scaling), synthetic data drawn
from the model, EM started from the fixed perturbed initial model of SURVEY.md 8d.  A "step" is one full EM
iteration: fused E-step over every resident trajectory (emission + forward + backward + statistics), one
all-reduce of the packed statistics across ranks, host M-step.

One JSON line on stdout (rank 0).  `value` is measured with the observations resident in HBM; `e2e` repeats the
measurement through the public API with the observations in pinned HOST memory, copied to the device inside the
timed region every step, and the statistics read back.  `roofline` is for the dominant kernel, timed live with CUDA
events on the launching stream.  `cpu_baseline` times the reference's own C implementation (oracle/_ref, built from
/root/reference) on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nstates, trajectories per GPU, frames per trajectory, description)
    'c3': (10, 1024, 100000, 'C3: 10-state Gaussian HMM (dalton recipe), 1024 trajectories x 1e5 frames per GPU, Baum-Welch EM'),
    'c1': (3, 10, 10000, 'C1: 3-state Gaussian HMM, 10 trajectories x 1e4 frames, Baum-Welch EM'),
    'c2': (3, 100, 10000, 'C2-shape: 3-state Gaussian HMM, 100 trajectories x 1e4 frames, Baum-Welch EM'),
    'small': (10, 64, 20000, 'reduced C3 shape for quick checks: 10 states, 64 x 2e4 frames'),
    'n3': (3, 4096, 100000, 'memory-bound regime: 3-state Gaussian HMM (C1/C2 model), 4096 trajectories x 1e5 frames per GPU, Baum-Welch EM'),
    'n32': (32, 256, 100000, 'C5 model family, batched: 32-state Gaussian HMM, 256 trajectories x 1e5 frames per GPU, Baum-Welch EM (FP64-pipe bound; set BHMM_B200_PANEL=1|2 for the tensor-pipe kernels)'),
}


def algorithmic_bytes_per_frame(N):
    """Baum-Welch Gaussian moves 16 + 16 N bytes per frame and iteration: the forward kernel reads
    the observation (8) and writes alpha (8N); the backward+statistics kernel reads the observation (8) and alpha (8N)."""
    return {'forward': 8 + 8 * N, 'backward_stats': 8 + 8 * N, 'iteration': 16 + 16 * N}


# ------------------------------------------------------------------------------------------------- data
def synth_gaussian_gpu(N, K, T, seed, device):
    """Draw K trajectories of T frames from the dalton model on the GPU (torch ops; setup, not the hot path)."""
    import torch
    from bhmm_b200.util import testsystems as ts
    rng = np.random.default_rng(seed)
    pi, A, means, sigmas = ts.dalton_parameters(N, rng=rng)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    cumA = torch.tensor(np.cumsum(A, axis=1), device=device)
    cumA[:, -1] = 1.0
    cpi = torch.tensor(np.cumsum(pi), device=device)
    cpi[-1] = 1.0
    S = torch.empty((T, K), dtype=torch.int64, device=device)
    u = torch.rand(K, generator=g, device=device, dtype=torch.float64)
    s = (u[:, None] > cpi[None, :]).sum(dim=1).clamp_(max=N - 1)
    S[0] = s
    block = 2000
    for t0 in range(1, T, block):
        t1 = min(T, t0 + block)
        U = torch.rand((t1 - t0, K), generator=g, device=device, dtype=torch.float64)
        for k in range(t1 - t0):
            s = (U[k][:, None] > cumA[s]).sum(dim=1).clamp_(max=N - 1)
            S[t0 + k] = s
    S = S.t().contiguous()                       # (K, T)
    mu = torch.tensor(means, device=device)
    sg = torch.tensor(sigmas, device=device)
    O = mu[S] + sg[S] * torch.randn((K, T), generator=g, device=device, dtype=torch.float64)
    return pi, A, means, sigmas, O


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi sampled every 100 ms while the timed region runs (B200_PROFILING.md clocks line)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.fh = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.fh.close()
            sm, mx, reasons = [], [], set()
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            for line in open(self.path):
                p = [x.strip() for x in line.split(',')]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0]))
                    mx.append(float(p[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[3:7]):
                    if v.lower() == 'active':
                        reasons.add(nm)
            os.unlink(self.path)
            if sm:
                hi = [v for v in sm if v >= 0.5 * max(sm)]      # samples under load
                out.update(sm_mhz=float(np.median(hi)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                           samples=len(sm))
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------- reference arm
def _testsystems():
    """bhmm_b200/util/testsystems.py loaded BY PATH (numpy only): the reference arm and the cpu_baseline leg draw the
    same synthetic data as the CUDA arm without importing the bhmm_b200 package (which would map libbhmm_b200.so)."""
    import importlib.util
    name = '_bench_testsystems'
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'bhmm_b200', 'util', 'testsystems.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


_THREAD_VARS = ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS')


def _single_threaded_blas():
    """Every CPU worker is ONE core: numpy's BLAS (gamma.T.dot(obs), np.dot in the M-step) must not spawn a thread
    team per worker (round 1 ran cores x cores threads and under-reported the reference about 50-fold)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:                                   # pragma: no cover
        import contextlib
        return contextlib.nullcontext()


def _ref_worker(args):
    """One host core, one leg of the reference's call sequence through the reference's own C functions:
    'em'      maximum_likelihood.py:249-265 per trajectory (p_obs, forward, backward, state_probabilities,
              transition_counts) + the two-pass Gaussian M-step gaussian.py:214-272 on this core's share;
    'gibbs'   bayesian_sampling.py:325-329 per trajectory (p_obs, forward, sample_path) + the path statistics of
              generic_hmm.py:297-334,398-431;
    'viterbi' maximum_likelihood.py:347-349 per trajectory (p_obs, viterbi)."""
    obs_list, A, pi, means, sigmas, kind, leg = args
    from oracle import oracle as orc
    with _single_threaded_blas():
        o = orc.Oracle(kind)
        t0 = time.perf_counter()
        if leg == 'em':
            st = o.estep_gaussian(obs_list, A, pi, means, sigmas, ignore_outliers=True)
            orc.mstep_gaussian(obs_list, st['gammas'])
        elif leg == 'gibbs':
            paths = []
            for k, obs in enumerate(obs_list):
                pobs = o.gaussian_p_obs(obs, means, sigmas, True)
                lp, alpha = o.forward(A, pobs, pi)
                paths.append(o.sample_path(alpha, A, seed=1 + k))
            o.path_stats(paths, obs_list, A.shape[0])
        elif leg == 'viterbi':
            for obs in obs_list:
                o.viterbi(A, o.gaussian_p_obs(obs, means, sigmas, True), pi)
        else:
            raise ValueError(leg)
        return time.perf_counter() - t0, sum(len(x) for x in obs_list)


def cpu_reference_iteration(obs, A, pi, means, sigmas, cores, kind, leg='em'):
    """One step of `leg` over the sample `obs` (list of arrays) spread over `cores` spawned single-threaded workers.
    Returns (seconds = slowest worker, frames)."""
    import multiprocessing as mp
    shares = [obs[i::cores] for i in range(cores)]
    shares = [s for s in shares if s]
    jobs = [(s, A, pi, means, sigmas, kind, leg) for s in shares]
    # fresh interpreters (spawn) whose numpy/BLAS start with one thread: a forked child of a process that already runs a
    # BLAS thread team keeps paying for it even under threadpoolctl (measured here: 5.3 M vs 7.8 M frames*iters/s on 8 cores)
    saved = {k: os.environ.get(k) for k in _THREAD_VARS}
    os.environ.update({k: '1' for k in _THREAD_VARS})
    try:
        ctx = mp.get_context('spawn')
        with ctx.Pool(len(jobs)) as pool:
            res = pool.map(_ref_worker, jobs)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return max(r[0] for r in res), sum(r[1] for r in res)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def reference_kind():
    from oracle import oracle as orc
    orc.build(ref=os.path.isdir('/root/reference'))
    return 'reference' if orc.have_reference_lib() else 'port'


def cpu_sample(N, K, T, seed, cores, per_core):
    ts = _testsystems()
    ntraj = max(1, min(K, cores * per_core))
    pi, A, means, sigmas, O, S = ts.gaussian_observations(N, ntraj, T, seed=seed)
    pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
    return [O[k] for k in range(ntraj)], A0, pi0, m0, s0


def cpu_baseline_block(N, K, T, cores, per_core, kind, legs=('em', 'gibbs', 'viterbi')):
    """The reported CPU baseline: (1) ONE core, as shipped -- the reference has no parallelism of its
    own; (2) all host cores, one spawned single-threaded worker per core over disjoint trajectory subsets.  The headline
    `value` is the all-cores Baum-Welch figure; the Gibbs and Viterbi legs and the one-core figures ride beside it."""
    obs, A0, pi0, m0, s0 = cpu_sample(N, K, T, 3, cores, per_core)
    workers = min(cores, len(obs))
    out = {'unit': 'frames*iters/s', 'cores': workers, 'kind': kind, 'cpu_model': cpu_model(),
           'sample': '%d trajectories x %d frames per leg (of %d x %d per GPU): %d spawned single-threaded workers, '
                     'BLAS capped at 1 thread per worker; the one-core figures time the first %d trajectories'
                     % (len(obs), T, K, T, workers, min(len(obs), max(1, per_core // 2)))}
    one = obs[:max(1, per_core // 2)]
    for leg in legs:
        t, f = cpu_reference_iteration(obs, A0, pi0, m0, s0, cores, kind, leg)
        t1, f1 = cpu_reference_iteration(one, A0, pi0, m0, s0, 1, kind, leg)
        out[leg] = {'all_cores': f / t, 'single_core_as_shipped': f1 / t1, 'per_core_efficiency': (f / t) / workers / (f1 / t1)}
    out['value'] = out['em']['all_cores']
    out['single_core_as_shipped'] = out['em']['single_core_as_shipped']
    return out


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the same EM iteration on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    N, K, T, desc = WORKLOADS[args.workload]
    cores = host_cores()
    kind = reference_kind()
    obs, A0, pi0, m0, s0 = cpu_sample(N, K, T, 3, cores, args.cpu_traj_per_core)
    for _ in range(args.warmup):
        cpu_reference_iteration(obs, A0, pi0, m0, s0, cores, kind)
    t_tot, frames = 0.0, 0
    for _ in range(args.steps):
        t, f = cpu_reference_iteration(obs, A0, pi0, m0, s0, cores, kind)
        t_tot += t
        frames += f
    value = frames / t_tot
    workers = min(cores, len(obs))
    t1, f1 = cpu_reference_iteration(obs[:1], A0, pi0, m0, s0, 1, kind)
    sample = ('%d trajectories x %d frames per step (of %d x %d per GPU), %d spawned single-threaded workers (BLAS capped '
              'at 1 thread each)' % (len(obs), T, K, T, workers))
    line = {
        'impl': 'reference', 'metric': 'Baum-Welch EM throughput (frames x iterations per second)', 'value': value,
        'unit': 'frames*iters/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': desc, 'nstates': N, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'frames*iters/s', 'cores': workers, 'kind': kind,
                         'sample': sample, 'cpu_model': cpu_model(), 'single_core_as_shipped': f1 / t1,
                         'per_core_efficiency': value / workers / (f1 / t1)},
        'e2e': {'value': value, 'unit': 'frames*iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
def em_step(batch, model, dist, N):
    """One Baum-Welch iteration through the public pieces: engine E-step, all-reduce, host M-step."""
    from bhmm_b200.engine import unpack_stats
    from bhmm_b200.util import tmatrix
    A, pi, means, sigmas = model
    stats = batch.estep_gaussian(A, pi, means, sigmas)
    stats = dist.allreduce_sum(stats)
    st = unpack_stats(stats.cpu().numpy(), N)
    A = tmatrix.estimate_P(st['C'], reversible=False, mincount_connectivity=1e-16)
    pi = st['gamma0'] / st['gamma0'].sum()
    shift = st['wd'] / st['wsum']
    means = means + shift
    sigmas = np.sqrt(np.maximum(st['wdd'] / st['wsum'] - shift * shift, 1e-300))
    return (A, pi, means, sigmas), st['loglik']


def run_ours(args):
    import torch
    import torch.distributed as td
    from bhmm_b200 import _lib, dist
    from bhmm_b200.engine import TrajectoryBatch
    from bhmm_b200.util import testsystems as ts

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        td.init_process_group('nccl', device_id=dev)

    N, K, T, desc = WORKLOADS[args.workload]
    if args.trajectories:
        K = args.trajectories
    pi, A, means, sigmas, O = synth_gaussian_gpu(N, K, T, 3 + rank, dev)
    host_obs = torch.empty((K * T,), dtype=torch.float64, pin_memory=True)
    host_obs.copy_(O.reshape(-1))
    rows = K * T

    batch = TrajectoryBatch.from_concatenated(O.reshape(-1), [T] * K, N, chunk=args.chunk, warm=args.warm)
    del O
    batch.set_profiling(True)
    lane_family = batch.uses_lane_kernels
    pi0, A0, m0, s0 = ts.perturbed_initial_model(A, means, N)
    model = (A0, pi0, m0, s0)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement (the clock sampler starts before the warm-up so that nvidia-smi is up and
    # sampling every 100 ms by the time the timed region runs; only samples under load are summarised)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(args.warmup):
        model, ll = em_step(batch, model, dist, N)
    barrier()
    launches0 = _lib.lib.bhmm_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = {'forward': 0.0, 'backward_stats': 0.0}
    barrier()
    e0.record()
    for _ in range(args.steps):
        model, ll = em_step(batch, model, dist, N)
        k = batch.kernel_ms()
        kms['forward'] += k['forward']
        kms['backward_stats'] += k['backward_stats']
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib.bhmm_b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms = float(t.item())
    value = world * rows * args.steps / (ms * 1e-3)
    info = batch.info()
    # what the slowest rank looked like: launches of the timed region (fix-up sweeps and redone passes show up here) and
    # kernel time, maximum over ranks
    rk = torch.tensor([float(launches), kms['forward'] + kms['backward_stats']], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(rk, op=td.ReduceOp.MAX)
    max_rank_launches, max_rank_kernel_ms = int(rk[0].item()), float(rk[1].item()) / args.steps

    # ---- end-to-end: observations start in pinned host memory every step.  Every step's inputs are copied host ->
    # device inside the timed region and its statistics are read back; the copy of step k+1 runs on a second stream
    # while step k computes (double-buffered device input), as a serving loop would do it.
    e2e_steps = max(1, min(args.steps, 5))
    obs_a = batch.obs
    obs_b = torch.empty_like(obs_a)
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    bufs = [obs_a, obs_b]
    barrier()
    e0.record()
    with torch.cuda.stream(copy_stream):
        bufs[0].copy_(host_obs, non_blocking=True)
        ready[0].record(copy_stream)
    for k in range(e2e_steps):
        if k + 1 < e2e_steps:
            with torch.cuda.stream(copy_stream):
                bufs[(k + 1) & 1].copy_(host_obs, non_blocking=True)
                ready[(k + 1) & 1].record(copy_stream)
        torch.cuda.current_stream(dev).wait_event(ready[k & 1])
        batch.obs = bufs[k & 1]
        model, ll = em_step(batch, model, dist, N)
    e1.record()
    barrier()
    batch.obs = obs_a
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_value = world * rows * e2e_steps / (float(t.item()) * 1e-3)
    # the same without overlap (copy, then compute), for reference
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        batch.set_observations(host_obs, non_blocking=True)
        model, ll = em_step(batch, model, dist, N)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_serial = world * rows * e2e_steps / (float(t.item()) * 1e-3)
    stats_bytes = 8 * (1 + N + N * N + 3 * N)

    # ---- Gibbs sweep (second half of the metric), device resident
    gsteps = max(1, min(args.steps, 5))
    A_, pi_, m_, s_ = model
    batch.gibbs_gaussian(A_, pi_, m_, s_, seed=1, sweep=0)
    barrier()
    e0.record()
    for sidx in range(gsteps):
        path, counts, sums, gll = batch.gibbs_gaussian(A_, pi_, m_, s_, seed=1, sweep=1 + sidx)
        c = dist.allreduce_sum(counts.clone()).cpu()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    gibbs_value = world * rows * gsteps / (float(t.item()) * 1e-3)

    if rank == 0:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        ab = algorithmic_bytes_per_frame(N)
        family = 'lane<N=%d,EM_GAUSS>' % N if lane_family else ('panel<EM_GAUSS>' if (os.environ.get('BHMM_B200_PANEL', '1') in ('1', '2') and 17 <= N <= 104) else 'team<EM_GAUSS>')
        dom = 'backward_stats' if kms['backward_stats'] >= kms['forward'] else 'forward'
        dom_ms = kms[dom] / args.steps
        # the dominant kernel also walks the warm-up frames; only the chain's own frames count as algorithmic bytes
        achieved = ab[dom] * rows / (dom_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom, {}).get('dram_bytes_per_frame')
                traffic = traffic * rows if traffic is not None else None
            except Exception:
                traffic = None
        # CPU baseline: reference C implementation on the host cores, bounded sample
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_baseline_block(N, K, T, host_cores(), args.cpu_traj_per_core, reference_kind())
        line = {
            'metric': 'Baum-Welch EM throughput (frames x iterations per second)',
            'value': value, 'unit': 'frames*iters/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': desc, 'nstates': N, 'trajectories_per_gpu': K, 'frames_per_trajectory': T,
                       'chunk': info['chunk'], 'warm': info['warm'], 'chains_per_gpu': info['chains'],
                       'l2': 'inputs larger than L2: %.1f GB observations + %.1f GB forward variables streamed per step'
                             % (rows * 8 / 1e9, rows * N * 8 / 1e9),
                       'certification': {'fixups_fwd': info['fixups_fwd'], 'fixups_bwd': info['fixups_bwd'],
                                         'worst_fwd': info['worst_fwd'], 'worst_bwd': info['worst_bwd'],
                                         'max_rank_launches': max_rank_launches,
                                         'max_rank_kernel_ms_per_step': max_rank_kernel_ms}},
            'roofline': {'bound': 'hbm', 'kernel': ('k_backward_stats_%s' if dom == 'backward_stats' else 'k_forward_%s') % family,
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_frame': ab[dom], 'kernel_ms': dom_ms,
                         'all_kernels_ms': {k2: v / args.steps for k2, v in kms.items()},
                         'iteration_frac': (ab['iteration'] * rows / (ms / args.steps * 1e-3) / 1e9) / peak,
                         # companion FP64 roofline: algorithmic flops 6N^2+40N per frame and iteration
                         #  against the vector-DFMA peak measured with tools/micro/fp64_peak.cu
                         'fp64': {'flops_per_frame': 6 * N * N + 40 * N,
                                  'achieved_tflops': (6 * N * N + 40 * N) * value / max(world, 1) / 1e12,
                                  'peak_tflops': 34.2, 'frac': (6 * N * N + 40 * N) * value / max(world, 1) / 34.2e12}},
            'cpu_baseline': cpu_baseline,
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'frames*iters/s', 'h2d_bytes_per_step': rows * 8,
                    'd2h_bytes_per_step': stats_bytes, 'steps': e2e_steps,
                    'note': 'input copy of step k+1 overlaps the compute of step k (double-buffered device input)',
                    'value_without_overlap': e2e_serial},
            'gpu_launches': int(launches),
            'gibbs': {'value': gibbs_value, 'unit': 'frames*sweeps/s', 'steps': gsteps,
                      'note': 'forward + time-parallel backward sampling + path statistics, Philox uniforms'},
            'loglik': ll,
        }
        print(json.dumps(line))
    batch.close()
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--trajectories', type=int, default=0, help='override trajectories per GPU')
    ap.add_argument('--chunk', type=int, default=0)
    ap.add_argument('--warm', type=int, default=0)
    ap.add_argument('--cpu-traj-per-core', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
