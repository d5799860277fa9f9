'''
Profiling driver for the fused day pipeline: builds the BASELINE C2 sim (or --pop-size agents), runs it to --day with cvb_run_days,
then brackets --days more days with cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly those launches.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/fused python profiles/prof_fused.py
'''
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import covasim_b200 as cv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--pop-size', type=int, default=1_000_000)
ap.add_argument('--day', type=int, default=80)
ap.add_argument('--days', type=int, default=2)
ap.add_argument('--n-days', type=int, default=180)
ap.add_argument('--no-fused', action='store_true')
args = ap.parse_args()

pars = dict(pop_size=args.pop_size, pop_type='hybrid', n_days=args.n_days, pop_infected=max(1, int(0.005 * args.pop_size)), rand_seed=1, verbose=0)
sim = cv.Sim(pars, interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), cv.contact_tracing(trace_probs=0.3, start_day=30)],
             pop_exact=False, fused=not args.no_fused)
sim.initialize()
sim.run(until=args.day)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sim.run(until=args.day + args.days, reset_seed=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('profiled days', args.day, '..', args.day + args.days - 1, 'n_exposed', int(sim.people.exposed.sum()), 'fused_days', sim.fused_days)
