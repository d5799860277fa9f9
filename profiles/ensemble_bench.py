#!/usr/bin/env python
'''
Ensemble throughput (BASELINE config 3 shape, scaled): cv.MultiSim of `--members` hybrid sims of `--pop-size` agents on the
visible GPU(s) of this process (members are stepped round-robin, one stream).  Prints one JSON line.

    python profiles/ensemble_bench.py --members 32 --pop-size 100000 --n-days 180
'''
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--members', type=int, default=32)
    ap.add_argument('--pop-size', type=int, default=100_000)
    ap.add_argument('--n-days', type=int, default=180)
    ap.add_argument('--pop-gen', default='device')
    args = ap.parse_args()
    import covasim_b200 as cv
    base = cv.Sim(pop_size=args.pop_size, pop_type='hybrid', n_days=args.n_days, pop_infected=max(1, args.pop_size // 200), rand_seed=1, verbose=0,
                  interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), cv.contact_tracing(trace_probs=0.3, start_day=30)],
                  pop_exact=False, pop_gen=args.pop_gen)
    msim = cv.MultiSim(base, n_runs=args.members)
    msim.init_sims()
    t0 = time.perf_counter()
    for sim in msim.sims:
        sim.initialize()
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    msim.run(keep_people=False)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    msim.reduce()
    ad = args.members * args.pop_size * (args.n_days + 1)
    print(json.dumps(dict(workload='ensemble (C3 shape)', members=args.members, pop_size=args.pop_size, n_days=args.n_days, n_gpus=torch.cuda.device_count(),
                          run_s=el, init_s=t_init, agent_days_per_s=ad / el, us_per_member_day=1e6 * el / (args.members * (args.n_days + 1)),
                          median_cum_infections=float(msim.results['cum_infections'].values[-1]))))


if __name__ == '__main__':
    main()
