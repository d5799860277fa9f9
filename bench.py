#!/usr/bin/env python
'''
bench.py -- agent-days/s of the per-timestep simulation hot path on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload = "C2"): one 1M-agent hybrid sim (h/s/w/c layers, ~17.8M edges), 180 days
(181 time points), cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20) +
cv.contact_tracing(trace_probs=0.3, start_day=30), 0.5 % initially infected, synthetic population.

One "step" = one complete run of that sim (181 simulated days of the hot path).
  value : agent-days/s with the initial People + Layer arrays already resident in HBM; timed on the
          device with CUDA events around the day loop of Sim.run() (181 days = one cvb_run_days call).
  e2e   : the same through the public API with HOST buffers inside the timed region: Sim.restore()
          (H2D from pinned memory of the saved People state -- in its compact form: per array either one
          value + the exceptions or the array as it is -- and of every edge list) + Sim.run() (181 days +
          finalize(), which reads the result tables back to the host).  h2d_bytes_per_step counts what is sent.
  roofline : the dominant kernel of the step (by CUDA-event time measured live around every launch of the day
          loop, in a separate pass): algorithmic bytes per launch / mean launch duration vs the measured HBM peak.  "kernels"
          lists the same for every kernel of the day; "edge_pass_dense" times the dense edge-streaming pass
          (12*E + 8*N bytes per launch, the form dynamic layers use) on its own with L2 flushed between launches.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, Numba parallel='full' on every host core) continuing the SAME
          sim from the GPU's day-40 People state for ~20 s (rank 0, N=1 only); the oracle port only if oracle/_ref is missing.
  ensembles : BASELINE configs 3 and 5 as blocks -- members of 100k / 50k agents advanced in lockstep on one GPU.

N > 1 (torchrun): weak scaling -- every rank runs its own member of an ensemble (same configuration,
seed + rank; reference run.py:1363-1365), no data-path collective; value = total agent-days / max time.
In addition (block "partition"): ONE simulation agent-partitioned over the N GPUs (BASELINE config 4 recipe, 12.5M agents per
GPU), with one exchange of 1 byte per agent per day on the data path (over peer memory: cvb_peer_push + signal barrier; ncclAllGather
where peer memory cannot be mapped), and a bit-identity check of a partitioned run against the single-GPU run of the same simulation.

--impl reference: ONE complete run of the workload by the unmodified reference on the host cores (oracle/_ref; the
oracle port, time-bounded, only if the install is missing), timed as --steps consecutive sim.run(until=...) chunks.
'''
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
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'agent-days/sec (1M-agent hybrid)'
UNIT = 'agent-days/s'


def workload_pars(args, seed):
    return dict(pop_size=args.pop_size, pop_type='hybrid', n_days=args.n_days, pop_infected=max(1, int(0.005 * args.pop_size)),
                rand_seed=seed, verbose=0)


def workload_interventions(mod):
    return [mod.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), mod.contact_tracing(trace_probs=0.3, start_day=30)]


def config_block(args, extra=None):
    cfg = dict(workload='C2', pop_size=args.pop_size, pop_type='hybrid', n_days=args.n_days, npts=args.n_days + 1,
               interventions='test_prob(symp_prob=0.1,asymp_prob=0.01,start_day=20)+contact_tracing(trace_probs=0.3,start_day=30)',
               pop_infected=max(1, int(0.005 * args.pop_size)), rng='philox (native)',
               l2_policy='every step starts after the 200 MB People arena has been rewritten from its device-resident copy (a write larger than the 126 MB L2, so '
                         'the L2 holds none of the step\'s inputs when the timed region starts); inside a step the 181 days run back to back as in production')
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    ''' nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe) '''
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in open(self.path):
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(smax)), reasons=sorted(reasons), samples=len(sm))


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import covasim_b200 as cv

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (covasim_b200 has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    numa_cpus = cv.bind_to_device_numa(local_rank) if world > 1 else None      # pinned snapshot buffers next to this rank's GPU
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    n_gpus = world

    pars = workload_pars(args, seed=1 + rank)
    sim = cv.Sim(pars, interventions=workload_interventions(cv), pop_exact=False)
    t0 = time.time()
    sim.initialize()
    t_init = time.time() - t0
    snap = sim.snapshot(pinned=True)                    # the step's inputs, in pinned host memory
    dev_snap = {k: sim.people[k].clone() for k in sim.people.keys()}      # device-resident copy for the `value` leg
    n_edges = {lk: len(l) for lk, l in sim.people.contacts.items()}
    E, N, npts = sum(n_edges.values()), sim.n, sim.npts
    agent_days = N * npts

    def reset_on_device():
        ''' Rewind to day 0 without touching the host: inputs stay resident in HBM '''
        for k, v in dev_snap.items():
            sim.people[k].copy_(v)
        sim.restore_light(snap)

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

    # ---- value: device-resident inputs, CUDA events around the day loop --------------------------------
    # The day loop is what Sim.run() executes (Sim._advance): every stretch of days that needs no host decision is ONE
    # cvb_run_days call -- in this workload all 181 days.
    def run_days():
        sim.set_seed()
        sim._advance(sim.npts)

    for _ in range(args.warmup):
        reset_on_device()
        run_days()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = cv._capi.lib.cvb_launch_count()
    dev_ms = 0.0
    barrier()
    for _ in range(args.steps):
        reset_on_device()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_days()
        b.record()
        barrier()
        dev_ms += max_over_ranks(a.elapsed_time(b))
    launches = cv._capi.lib.cvb_launch_count() - launches0          # every kernel launched inside the timed region (all steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = dev_ms / args.steps
    value = n_gpus * agent_days / (ms_per_step / 1e3)
    fused_days = sim.fused_days

    # per-kernel device time: the same day loop once more with CUDA events around every launch (outside the timed region: the
    # events serialise the launches; the shares are what the ncu launch list must agree with)
    peak, peak_src = measured_peak_gbs()
    nv, L = sim['n_variants'], len(n_edges)
    kernels, roofline = {}, None
    if fused_days == sim.npts:
        reset_on_device()
        sim.fused_timing(True)
        run_days()
        timing = sim.fused_timing()
        sim.fused_timing(False)
        work = sim.edge_work()                               # per day: adjacency entries visited, transmitters
        total_timed = sum(ms for ms, _ in timing.values())
        # Per-kernel byte models: every array the unfused kernels of the same work must read or write, once
        # (comparable with round 1); the fused kernels move fewer bytes (dram_traffic_bytes is what ncu measured)
        algo = {
            'day_begin': (13 + 12 + 2 * nv + 8 * nv) * N + (9 + 16 + 12 * nv) * N + 11 * N + 4 * N,   # update_nab + counts, update_states_pre + check_immunity, test_prob, case selection
            'day_mid': (8 + 28 + 16) * N + N // 8,
            'edge_pass': float(24 * work[:, 0].sum() + 32 * work[:, 1].sum()) / max(npts, 1) if sim._adj is not None else 12 * E + 8 * N,
        }
        for name, (ms, cnt) in timing.items():
            if not cnt:
                continue
            us = 1e3 * ms / cnt
            entry = dict(us_per_launch=us, launches_per_step=cnt, ms_per_step=ms, share_of_step=ms / total_timed if total_timed else None)
            if algo.get(name):
                entry['algorithmic_bytes_per_launch'] = algo[name]
                entry['achieved_gbs'] = algo[name] / (us * 1e-6) / 1e9 if us > 0 else 0.0
                entry['frac'] = entry['achieved_gbs'] / peak
            kernels[name] = entry
        traffic = {}
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath))
            except Exception:
                traffic = {}
        for name, entry in kernels.items():                  # measured DRAM bytes per launch (one ncu --set full capture, committed)
            if traffic.get(name):
                entry['dram_traffic_bytes'] = traffic[name]
                entry['dram_traffic_frac_of_peak'] = traffic[name] / (entry['us_per_launch'] * 1e-6) / 1e9 / peak
        # The roofline object: the AGENT STATE PASS of SURVEY section 8(d) -- everything per-agent of the day (A2, A3, A6-A9, A16 and the
        # built-in testing), i.e. day_begin_kernel + day_mid_kernel -- with the survey's own algorithmic figure, 202 bytes per agent per
        # day (one read and at most one write of the People arrays the day touches, n_variants = 1).  The two kernels move far fewer
        # bytes than that (`traffic`): the packed state word lets them skip the arrays of agents nothing can happen to.
        pa = [kernels[k] for k in ('day_begin', 'day_mid') if k in kernels]
        us_pass = sum(k['us_per_launch'] for k in pa)
        algo_pass = (202 + 28 * (nv - 1)) * N
        tr_pass = sum(k.get('dram_traffic_bytes', 0) for k in pa) or None
        roofline = dict(bound='hbm', kernel='agent state pass = day_begin_kernel + day_mid_kernel', achieved=algo_pass / (us_pass * 1e-6) / 1e9, peak=peak,
                        unit='GB/s', frac=algo_pass / (us_pass * 1e-6) / 1e9 / peak, traffic=tr_pass, algorithmic_bytes_per_launch=algo_pass,
                        us_per_launch=us_pass, peak_source=peak_src, share_of_step=sum(k['share_of_step'] for k in pa),
                        note='algorithmic bytes = SURVEY section 8(d): 202 B x N per day for the fused agent state pass (230 B at three variants); '
                             'us_per_launch = the two kernels of the pass, CUDA events around every launch of cvb_run_days in a separate pass; per-kernel '
                             'figures (own byte models, measured DRAM traffic) under "kernels"; the dense edge-streaming pass (12*E + 8*N) under "edge_pass_dense"')

    # ---- the dense edge-streaming pass on its own (what dynamic layers use, and the survey's 12*E + 8*N figure) ----
    edge_dense = measure_dense_edge_pass(args, cv, sim, snap, peak, E, N) if (rank == 0 and not args.no_dense) else None

    # ---- e2e: host buffers in, results out, through the public API -----------------------------------------
    for _ in range(min(args.warmup, 2)):
        sim.restore(snap)
        sim.run()
    e2e_s = 0.0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        sim.restore(snap)                 # H2D: People arrays + edge lists from pinned host memory
        sim.run()                          # 181 days + finalize (D2H of the result tables)
        torch.cuda.synchronize()
        e2e_s += max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_gpus * agent_days / (e2e_s / args.steps)
    # where the end-to-end time goes (one extra, separately synchronised pass; not part of the timed region above)
    phases = {}
    barrier()
    t0 = time.perf_counter(); sim.restore(snap); torch.cuda.synchronize(); phases['restore_h2d_synchronised_ms'] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); sim.set_seed(); sim._advance(sim.npts); torch.cuda.synchronize(); phases['days_ms'] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); sim.finalize(); torch.cuda.synchronize(); phases['finalize_d2h_ms'] = 1e3 * (time.perf_counter() - t0)
    summary = dict(cum_infections=sim.summary['cum_infections'], cum_deaths=sim.summary['cum_deaths'],
                   cum_diagnoses=sim.summary['cum_diagnoses'], cum_quarantined=sim.summary['cum_quarantined'])

    # ---- cpu_baseline: the oracle continues the same sim from the GPU's day-40 state (rank 0, N = 1) ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_from_gpu_state(args, cv, sim, snap)

    ens = None
    if world == 1 and rank == 0 and not args.no_ensemble:
        del dev_snap
        torch.cuda.empty_cache()
        ens = ensemble_block(args, cv)
    part = None
    if world > 1 and not args.no_partition:
        part = partition_block(args, cv, world, rank, max_over_ranks)

    if rank == 0:
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                   higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                   config=config_block(args, dict(edges=E, edges_by_layer=n_edges, parallelism=f'ensemble x{n_gpus} (one member per GPU, no collective)' if n_gpus > 1 else 'single GPU',
                                                  init_s=t_init, numa_cpus=(f'{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)} cores next to the GPU)' if numa_cpus else None))),
                   clocks=clocks,
                   e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=sim.h2d_bytes(snap), d2h_bytes_per_step=sim.d2h_bytes(), ms_per_step=1e3 * e2e_s / args.steps, phases=phases),
                   gpu_launches=int(launches), roofline=roofline, kernels=kernels, edge_pass_dense=edge_dense, us_per_day=1e3 * ms_per_step / npts,
                   fused_days=int(fused_days),
                   epidemic=summary)
        if cpu is not None:
            out['cpu_baseline'] = cpu
        if ens is not None:
            out['ensembles'] = ens
        if part is not None:
            out['partition'] = part
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ensemble_block(args, cv):
    """
    Ensembles of SMALL simulations on one GPU (BASELINE configs 3 and 5; reference run.py:1406-1519 multi_run, one process per member).
    Every member is far too small to fill a B200, so MultiSim advances all of them in lockstep through cvb_run_days_multi: one host
    thread, one stream per member, five launches per member-day.  Reported: agent-days/s over the members' day loops (wall clock
    around the lockstep run with a device synchronisation on both sides; initialisation and finalize excluded, as in `value`).
      C3: --ens-members members of the 100k-agent hybrid sim (seed + i), test_prob + contact_tracing, 60 days
      C5: --ens-members members of a 50k-agent random-network sim with a dynamic layer (regenerated on the device every day), each
          with its own (beta, rel_death_prob) from the sweep grid of the reference's calibration example; Fit mismatch of every
          member against the example data file (cumulative diagnoses / deaths), computed from the result tables.
    """
    import torch
    out = {}
    M = int(args.ens_members)

    def run(sims):
        msim = cv.MultiSim(sims)
        t0 = time.perf_counter()
        for sim in sims:
            sim.initialize()
            sim.set_seed()
        torch.cuda.synchronize()
        t_init = time.perf_counter() - t0
        snaps = [sim.snapshot(pinned=False) for sim in sims] if args.ens_reps > 1 else None
        best, launches = None, 0
        for rep in range(args.ens_reps):
            if rep:
                for sim, snap in zip(sims, snaps):
                    sim.restore(snap)
                    sim.set_seed()
            torch.cuda.synchronize()
            launches0 = cv._capi.lib.cvb_launch_count()
            t0 = time.perf_counter()
            msim._advance_members(sims)
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            launches = cv._capi.lib.cvb_launch_count() - launches0
            best = el if best is None else min(best, el)
        for sim in sims:
            sim.finalize()
        return msim, best, t_init, launches

    # ---- C3 ----
    sims = [cv.Sim(dict(pop_size=100_000, pop_type='hybrid', n_days=60, pop_infected=500, rand_seed=1 + i, verbose=0), pop_gen='device',
                   interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=10), cv.contact_tracing(trace_probs=0.3, start_day=15)])
            for i in range(M)]
    msim, el, t_init, launches = run(sims)
    npts = sims[0].npts
    out['C3'] = dict(members=M, pop_size=100_000, n_days=60, seconds=el, us_per_member_day=1e6 * el / (M * npts), agent_days_per_s=M * 100_000 * npts / el,
                     init_s=t_init, gpu_launches=int(launches), fused_days=int(min(s.fused_days for s in sims)),
                     cum_infections_median=float(np.median([s.summary['cum_infections'] for s in sims])))
    del sims, msim
    torch.cuda.empty_cache()
    # ---- C5 ----
    betas = np.linspace(0.005, 0.020, max(int(np.sqrt(M)), 1))
    rdps = np.linspace(0.5, 3.0, max(-(-M // len(betas)), 1))
    grid = [(b, r) for b in betas for r in rdps][:M]
    sims = [cv.Sim(dict(pop_size=50_000, pop_type='random', n_days=45, pop_infected=100, rand_seed=1, verbose=0, beta=float(b), rel_death_prob=float(r),
                        dynam_layer=dict(a=1)), pop_gen='device',
                   interventions=[cv.test_prob(symp_prob=0.2, asymp_prob=0.005, start_day=0)]) for b, r in grid]
    msim, el, t_init, launches = run(sims)
    npts = sims[0].npts
    results = [{k: s.results[k].values for k in s.result_keys()} for s in sims]
    data = example_data()
    source = data.pop('source')
    mism = cv.fit_members(results, data, npts)
    best = int(np.argmin(mism))
    out['C5'] = dict(members=len(sims), pop_size=50_000, n_days=45, dynamic_layer=True, seconds=el, us_per_member_day=1e6 * el / (len(sims) * npts),
                     agent_days_per_s=len(sims) * 50_000 * npts / el, init_s=t_init, gpu_launches=int(launches), fused_days=int(min(s.fused_days for s in sims)),
                     fit=dict(data=source, best_member=best, best_beta=float(grid[best][0]), best_rel_death_prob=float(grid[best][1]),
                              best_mismatch=float(mism[best]), worst_mismatch=float(np.max(mism))))
    del sims, msim
    torch.cuda.empty_cache()
    return out


def example_data():
    """ The data file BASELINE config 5 fits against (reference examples/example_data.csv, copied into oracle/_ref by build()); a synthetic
    series of the same shape when it is not there """
    path = os.path.join(ROOT, 'oracle', '_ref', 'example_data.csv')
    if os.path.exists(path):
        import csv
        rows = list(csv.DictReader(open(path)))
        data = dict(date=[r['date'] for r in rows], source='reference examples/example_data.csv')
        for k in ('new_diagnoses', 'new_tests', 'new_deaths'):
            col = np.array([float(r[k]) if r[k] not in ('', None) else np.nan for r in rows])
            data['cum_' + k[4:]] = np.nancumsum(col)
        return data
    days = np.arange(32)
    diag = np.round(2 * np.exp(0.12 * days))
    return dict(day=days, cum_diagnoses=np.cumsum(diag), cum_deaths=np.cumsum(np.round(0.02 * diag)), source='synthetic exponential series (example file not present)')


def partition_block(args, cv, world, rank, max_over_ranks):
    """
    N > 1: ONE simulation agent-partitioned over the N GPUs (BASELINE config 4 recipe: hybrid, alpha + delta, waning, test_prob +
    contact_tracing + vaccinate_prob + booster), weak-scaled at --part-agents agents per GPU, population generated on the device.
    One NCCL all-gather of 1 byte per agent per day on the data path (plus a 1-bit-per-agent case bitmap on tracing days).
    Also checks a 200k-agent partitioned run against the single-GPU run of the same simulation on rank 0, bit for bit.
    """
    import torch
    import torch.distributed as dist
    import scenarios
    out = {}
    # ---- bit identity: N ranks == 1 GPU -------------------------------------------------------------------
    spec = dict(pars=dict(pop_size=200_000, pop_infected=1000, pop_type='hybrid', n_days=60, verbose=0, rand_seed=1),
                interventions=[('test_prob', dict(symp_prob=0.1, asymp_prob=0.01, start_day=10)), ('contact_tracing', dict(trace_probs=0.3, start_day=15))])
    psim = cv.Sim(**scenarios.build(cv, spec), partition=True, pop_exact=False)
    psim.run()
    log = psim.infection_log
    gathered = {k: psim._comm.gather_objects(psim.people.to_numpy(k)) for k in ('exposed', 'date_exposed', 'date_recovered', 'date_dead', 'quarantined', 'date_diagnosed', 'n_infections')}
    ok = None
    if rank == 0:
        ref = cv.Sim(**scenarios.build(cv, spec), pop_exact=False).run()
        ok = all(np.array_equal(ref.people.to_numpy(k), np.concatenate(v, axis=-1), equal_nan=True) for k, v in gathered.items())
        ok = ok and all(np.allclose(psim.results[k].values, ref.results[k].values, rtol=1e-6 if k == 'r_eff' else 1e-12, atol=0, equal_nan=True) for k in ref.result_keys())
        ok = ok and all(np.array_equal(log[k], ref.infection_log[k]) for k in ('source', 'target', 'date', 'layer', 'variant'))
        out['bit_identity'] = dict(ok=bool(ok), pop_size=200_000, n_days=60, ranks=world, cum_infections=float(ref.summary['cum_infections']),
                                   what='People arrays, every result series and the infection log of the partitioned run == the single-GPU run')
    del psim
    torch.cuda.empty_cache()
    # ---- the weak-scaled C4 recipe ---------------------------------------------------------------------------
    n = int(args.part_agents) * world
    pars = dict(pop_size=n, pop_type='hybrid', n_days=args.part_days, pop_infected=max(1, n // 200), rand_seed=1, verbose=0, use_waning=True)
    variants = [cv.variant('alpha', days=5, n_imports=max(10, n // 20000)), cv.variant('delta', days=15, n_imports=max(10, n // 20000))]
    ivs = [cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=10), cv.contact_tracing(trace_probs=0.3, start_day=15),
           cv.vaccinate_prob('pfizer', days=list(range(10, 30)), prob=0.01), cv.vaccinate_prob('pfizer', days=[40], prob=0.05, booster=True, label='booster')]
    t0 = time.time()
    sim = cv.Sim(pars, variants=variants, interventions=ivs, pop_exact=False, partition=True, pop_gen='device')
    sim.initialize()
    torch.cuda.synchronize()
    t_init = time.time() - t0
    snap = sim.snapshot(pinned=False)
    best = None
    for rep in range(3):
        sim.restore(snap)
        sim.set_seed()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sim._advance(sim.npts)               # fused days in blocks (cvb_fused_phase + the two exchanges), vaccination / variant days per step
        b.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(a.elapsed_time(b))
        if rep > 0:
            best = ms if best is None else min(best, ms)
        fused_days = sim.fused_days
        sim.finalize()
    # one more run with CUDA events around every launch and collective: fused days through the library's timers, per-step days through the host's
    sim.restore(snap)
    sim.set_seed()
    sim.fused_timing(True)
    sim.collective_timers = {}
    sim._advance(sim.npts)
    torch.cuda.synchronize()
    kernels = {'fused/' + k: round(1e3 * ms_ / sim.npts, 1) for k, (ms_, cnt) in sim.fused_timing().items() if cnt}
    kernels.update({k: round(float(np.sum([x.elapsed_time(y) for x, y in v])) * 1e3 / sim.npts, 1) for k, v in sim.collective_timers.items()})
    sim.fused_timing(False)
    sim.collective_timers = None
    sim.finalize()
    if rank == 0:
        out.update(workload='C4 recipe (hybrid, alpha + delta, waning, test_prob + contact_tracing + vaccinate_prob + booster), weak-scaled, device population',
                   pop_size=n, agents_per_gpu=int(args.part_agents), n_days=args.part_days, n_gpus=world, ms_per_run=best, us_per_day=1e3 * best / sim.npts,
                   agent_days_per_s=n * sim.npts / (best / 1e3), init_s=t_init, fused_days=int(fused_days), kernel_us_per_day=kernels,
                   allgather_us_per_day=round(kernels.get('allgather_codes', 0.0) + kernels.get('allgather_cases', 0.0), 1),
                   exchange_bytes_per_day_per_rank=int(sim._chunk * world + sim._chunk * world // 8), collective=('peer-memory exchange (cvb_peer_push: every rank stores its chunk into all ranks\' buffers over NVLink, torch symmetric memory + signal barrier)' if sim._peer is not None else 'ncclAllGather (torch.distributed all_gather_into_tensor)') + ', 1 byte per agent per day + 1 bit per agent on tracing days',
                   hbm_gb_per_gpu=torch.cuda.max_memory_allocated() / 1e9,
                   cum_infections=float(sim.summary['cum_infections']), cum_deaths=float(sim.summary['cum_deaths']), cum_doses=float(sim.summary['cum_doses']))
    del sim
    torch.cuda.empty_cache()
    return out


def measure_dense_edge_pass(args, cv, sim, snap, peak, E, N, day=60, reps=20):
    ''' Time cvb_edge_pass alone, with every layer streamed densely, on the mid-epidemic state of the same sim '''
    import torch
    day = min(day, max(args.n_days - 2, 0))
    sim.restore(snap)
    sim.set_seed()
    while sim.t < day:
        sim.step()
    call = cv._capi.call
    h, st, t = sim._handle, sim._stream_ptr, sim.t
    call('cvb_bind_adjacency', h, None, None, 0, 0)          # force the dense path for every layer
    call('cvb_update_states_pre', h, t, st)
    call('cvb_post_and_prepare', h, t, st)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device=sim.device)         # 256 MB > L2 (126 MB)
    times = []
    for r in range(reps + 3):
        # evict the edge lists from L2 between repetitions by READING a larger buffer: a write-flush (fill_) would leave
        # ~100 MB of dirty lines whose write-back competes with the timed kernel's reads for HBM bandwidth
        flush.max()
        # ... and put the per-agent records back the way the simulated day leaves them: prepare_transmission has just
        # written the per-layer {rel_trans, rel_sus} pairs the dense pass gathers and the transmit bitmap (32 MB + 125 KB at C2), so the edge pass finds
        # them in L2 while the edge lists (213 MB) come from HBM
        call('cvb_prepare_transmission', h, t, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        call('cvb_edge_pass', h, t, st)
        b.record()
        torch.cuda.synchronize()
        if r >= 3:
            times.append(a.elapsed_time(b))
    us = 1e3 * float(np.mean(times))
    algo = 12 * E + 8 * N
    sim._adj_dirty = True                                    # re-bind the adjacency before the sim is used again
    sim._build_adjacency()
    return dict(kernel='edge_pass_kernel (dense streaming, all layers)', us_per_launch=us, algorithmic_bytes_per_launch=algo,
                achieved_gbs=algo / (us * 1e-6) / 1e9, frac=algo / (us * 1e-6) / 1e9 / peak, day=int(t), reps=reps,
                l2='edge lists evicted between repetitions by reading 256 MB; per-agent records re-written by cvb_prepare_transmission as in the simulated day')


def run_ref_leg(leg, args, extra, timeout):
    """ One leg of oracle/ref_arm.py (the UNMODIFIED reference on host cores) in a fresh process; returns its JSON or None """
    cmd = [sys.executable, '-m', 'oracle.ref_arm', leg, '--pop-size', str(args.pop_size), '--n-days', str(args.n_days)] + [str(x) for x in extra]
    env = dict(os.environ)
    for k in ('OMP_NUM_THREADS', 'NUMBA_NUM_THREADS', 'MKL_NUM_THREADS'):      # torchrun pins these to 1; the reference arm uses every host core
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None
    for line in reversed(res.stdout.strip().splitlines()):
        if line.startswith('{'):
            return json.loads(line)
    sys.stderr.write(f'bench.py: reference leg "{leg}" failed (rc={res.returncode}):\n{res.stderr[-2000:]}\n')
    return None


def reference_installed():
    return os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'covasim')) or os.path.isdir('/root/reference/covasim')


def cpu_baseline_from_gpu_state(args, cv, sim, snap, t0=40, budget_s=20.0):
    """
    cpu_baseline: the UNMODIFIED reference (oracle/_ref, Numba parallel='full' on every host core) continuing the SAME sim from
    the GPU's day-t0 People state for a bounded number of days.  Falls back to the oracle port (kind "port") only when the
    reference install is missing.
    """
    import torch
    t0 = min(t0, max(args.n_days - 4, 0))
    sim.restore(snap)
    sim.set_seed()
    sim.run(until=t0)
    torch.cuda.synchronize()
    if reference_installed():
        path = os.path.join(tempfile.gettempdir(), f'cvb_state_{os.getpid()}.npz')
        arrays = {f'people/{k}': sim.people.to_numpy(k) for k in sim.people.keys()}
        lkeys = sim.people.layer_keys()
        for lk, l in sim.people.contacts.items():
            cols = l.to_numpy()
            for c in ('p1', 'p2', 'beta'):
                arrays[f'{c}/{lk}'] = cols[c]
        np.savez(path, t=np.int64(t0), layer_keys=np.array(lkeys), **arrays)
        try:
            out = run_ref_leg('continue', args, ['--state', path, '--budget', budget_s, '--numba-parallel', 'full'], timeout=600)
        finally:
            os.remove(path)
        if out is not None:
            return dict(value=out['agent_days_per_s'], unit=UNIT, cores=out['cores'], kind='reference',
                        sample=(f"unmodified Covasim {out['version']} (oracle/_ref; numba {out['numba']}, numba_parallel='full', {out['cores']} threads of "
                                f"{out['host_cores']} host cores) on days {out['first_day']}..{out['first_day'] + out['days'] - 1} of the same sim, continued from the "
                                f"GPU's People state; {out['s_per_day']:.3f} s/day"),
                        s_per_day=out['s_per_day'])
    from oracle import cvoracle as cvo
    pop = dict(age=sim.people.to_numpy('age').astype(np.float64), sex=sim.people.to_numpy('sex'),
               contacts={lk: l.to_numpy() for lk, l in sim.people.contacts.items()})
    orc = cvo.OracleSim(workload_pars(args, seed=1), interventions=workload_interventions(cvo), rng='philox', popdict=pop)
    orc.keep_log = False
    orc.initialize()
    for k in cvo.cvd.all_states:
        orc.P[k] = sim.people.to_numpy(k).copy()
    orc.t = t0
    orc.rng.set_seed(1)
    orc.step()                                           # warm-up day (page faults, allocator)
    days, t_start = 0, time.perf_counter()
    while time.perf_counter() - t_start < budget_s and orc.t < orc.npts - 1 and days < 40:
        orc.step()
        days += 1
    el = time.perf_counter() - t_start
    return dict(value=args.pop_size * days / el, unit=UNIT, cores=1, kind='port',
                sample=f'reference install missing (oracle/_ref): oracle port (NumPy restatement) on days {t0 + 1}..{t0 + days} of the same sim, continued from the GPU state; {el / days:.3f} s/day',
                s_per_day=el / days)


# ---------------------------------------------------------------------------------------------------
# the reference arm: the UNMODIFIED reference (oracle/_ref) on the box's host cores, same configuration
# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """
    One full run of the BASELINE workload by the reference's own code (numba_parallel='full', every host core), timed as
    --steps consecutive sim.run(until=...) chunks that together span day 0 .. the last day; the --warmup steps are a small sim
    of the same shape that compiles the Numba kernels.  A second, shorter leg reports numba_parallel='none' (the reference's
    deterministic default) on the first third of the run.  Falls back to the oracle port only if oracle/_ref is missing.
    """
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if not reference_installed():
        return run_reference_port(args)
    full = run_ref_leg('full', args, ['--chunks', max(args.steps, 1), '--numba-parallel', 'full'], timeout=1500)
    if full is None or not full['days_done']:
        return run_reference_port(args)
    third = max((args.n_days + 1) // 3, 4)
    none = run_ref_leg('full', args, ['--chunks', 1, '--numba-parallel', 'none', '--n-days', third - 1], timeout=900)
    value = full['agent_days_per_s']
    k = max(len(full['chunk_seconds']), 1)
    sample = (f"unmodified Covasim {full['version']} (oracle/_ref; numba {full['numba']}, numba_parallel='full', {full['cores']} threads of {full['host_cores']} host "
              f"cores): ONE complete run, days 0..{full['days_done'] - 1}, timed as {k} consecutive sim.run(until=) chunks ({full['seconds']:.1f} s in all, "
              f"{full['seconds'] / full['days_done']:.3f} s/day; init {full['init_s']:.1f} s and Numba compilation {full['warmup_s']:.1f} s excluded)")
    cpu = dict(value=value, unit=UNIT, cores=full['cores'], kind='reference', sample=sample, complete_run=full['complete'],
               chunk_seconds=full['chunk_seconds'], chunk_days=full['chunk_days'], cum_infections=full['cum_infections'])
    if none is not None and none['days_done']:
        cpu['numba_parallel_none'] = dict(value=none['agent_days_per_s'], cores=1, days=f"0..{none['days_done'] - 1}", seconds=none['seconds'],
                                          note="the reference's default (deterministic) setting on the first third of the run")
    out = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1e3 * full['seconds'] / k, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=config_block(args, dict(rng='mt19937 (reference streams)', init_s=full['init_s'], parallelism=f"host CPU, {full['cores']} Numba threads")),
               cpu_baseline=cpu, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


def run_reference_port(args):
    """ Fallback when oracle/_ref is missing: the oracle port (NumPy restatement, one core), time-bounded samples from day 0 """
    from oracle import cvoracle as cvo
    total = args.steps + args.warmup
    budget = 150.0 / max(total, 1)                        # seconds per step so the whole run ends within a few minutes
    pars = workload_pars(args, seed=1)
    t0 = time.time()
    base = cvo.OracleSim(pars, interventions=workload_interventions(cvo), rng='mt')
    base.keep_log = False
    base.initialize()
    t_init = time.time() - t0
    P0 = {k: v.copy() for k, v in base.P.items()}

    def one_step():
        for k, v in P0.items():
            base.P[k][...] = v
        base.t = 0
        base.pending_quar = {}
        base.pars['n_days'] = args.n_days
        base.rng.set_seed(pars['rand_seed'])
        ts = time.perf_counter()
        days = 0
        while days < args.n_days + 1 and (days < 4 or time.perf_counter() - ts < budget):
            base.step()
            days += 1
        return time.perf_counter() - ts, days

    for _ in range(args.warmup):
        one_step()
    runs = [one_step() for _ in range(args.steps)]
    el = float(np.sum([r[0] for r in runs])) / args.steps
    days = float(np.sum([r[1] for r in runs])) / args.steps
    value = args.pop_size * days / el
    sample = (f'reference install missing (oracle/_ref): oracle port (NumPy restatement, 1 host core), first {days:.0f} of {args.n_days + 1} days per step '
              f'(time-bounded at {budget:.0f} s per step; {el / days:.3f} s/day)')
    out = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * el,
               higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=config_block(args, dict(rng='mt19937 (reference streams)', init_s=t_init)),
               cpu_baseline=dict(value=value, unit=UNIT, cores=1, kind='port', sample=sample),
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--pop-size', type=int, default=1_000_000)
    ap.add_argument('--n-days', type=int, default=180)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-dense', action='store_true', help='skip the separate dense edge-pass measurement')
    ap.add_argument('--no-ensemble', action='store_true', help='skip the small-simulation ensemble block (BASELINE configs 3 and 5)')
    ap.add_argument('--ens-members', type=int, default=64, help='members per GPU in the ensemble block')
    ap.add_argument('--ens-reps', type=int, default=2)
    ap.add_argument('--no-partition', action='store_true', help='N > 1: skip the agent-partitioned single-simulation block')
    ap.add_argument('--part-agents', type=int, default=12_500_000, help='N > 1: agents per GPU of the partitioned simulation (BASELINE config 4: 100M over 8)')
    ap.add_argument('--part-days', type=int, default=60)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
