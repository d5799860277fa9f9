'''
Launch-shape sweep of the fused day kernels on the BASELINE C2 workload: for each (what, value) setting of cvb_tune the whole
181-day run is timed with CUDA events (device-resident inputs, as bench.py's `value`).
    python profiles/tune_fused.py [--pop-size N] > gpurun_out/tune.json
'''
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import covasim_b200 as cv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--pop-size', type=int, default=1_000_000)
ap.add_argument('--n-days', type=int, default=180)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--quick', action='store_true', help='only the default launch shapes')
ap.add_argument('--edge', action='store_true', help='sweep the sparse edge pass shapes')
args = ap.parse_args()
pars = dict(pop_size=args.pop_size, pop_type='hybrid', n_days=args.n_days, pop_infected=max(1, int(0.005 * args.pop_size)), rand_seed=1, verbose=0)
sim = cv.Sim(pars, interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), cv.contact_tracing(trace_probs=0.3, start_day=30)], pop_exact=False, pop_gen='device' if args.pop_size > 2_000_000 else 'host')
sim.initialize()
snap = sim.snapshot(pinned=True)
dev = {k: sim.people[k].clone() for k in sim.people.keys()}


def run_once(timing=False):
    for k, v in dev.items():
        sim.people[k].copy_(v)
    sim.restore_light(snap)
    sim.set_seed()
    if timing:
        sim.fused_timing(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    sim._advance(sim.npts)
    b.record()
    torch.cuda.synchronize()
    out = dict(us_per_day=1e3 * a.elapsed_time(b) / sim.npts)
    if timing:
        out['kernels'] = {k: round(1e3 * ms / max(n, 1), 2) for k, (ms, n) in sim.fused_timing().items() if n}
        sim.fused_timing(False)
    return out


def tune(**kw):
    ids = dict(begin_threads=0, begin_chunk=1, mid_threads=2, mid_chunk=3, edge_shape=4, edge_grid=5, infect_grid=6)
    for k, v in kw.items():
        cv._capi.call('cvb_tune', sim._handle, ids[k], int(v))


results = []
run_once()
configs = [dict()]
for th, ch in itertools.product((128, 256), (256, 512, 1024, 2048)):
    if ch >= th:
        configs.append(dict(begin_threads=th, begin_chunk=ch))
for th, ch in itertools.product((128, 256), (256, 512, 1024)):
    if ch >= th:
        configs.append(dict(mid_threads=th, mid_chunk=ch))
if args.quick:
    configs = [dict()]
if args.edge:
    configs = [dict()] + [dict(edge_shape=k) for k in range(1, 8)] + [dict(edge_grid=g) for g in (4, 6, 12, 16)] + [dict(infect_grid=g) for g in (1, 2)]
for cfg in configs:
    tune(begin_threads=0, begin_chunk=0, mid_threads=0, mid_chunk=0, edge_shape=0, edge_grid=0, infect_grid=0)
    tune(**cfg)
    best = min(run_once()['us_per_day'] for _ in range(args.reps))
    detail = run_once(timing=True)
    results.append(dict(config=cfg, us_per_day=round(best, 2), kernels=detail['kernels']))
    print(json.dumps(results[-1]), flush=True)
print(json.dumps(dict(cum_infections=float(sim.finalize().summary['cum_infections']))))
