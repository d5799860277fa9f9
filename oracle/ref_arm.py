'''
TEST / BENCH INFRASTRUCTURE -- runs the UNMODIFIED reference (Covasim 3.1.7) on host cores and times it.

The reference is pure Python + Numba, so "building" it is an install: ``install()`` (called by
``__graft_entry__.build()`` in the build container, where /root/reference exists) copies the reference's own
package directory, byte for byte, into the git-ignored ``oracle/_ref/covasim`` next to the sciris / pylab /
matplotlib stand-ins of ``oracle/shim``.  ``oracle/_ref/`` travels to the GPU box with the snapshot (it is
git-ignored, not gpurun-ignored), so ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the
reference's own code there -- never the oracle port, unless the install is missing (then the caller falls back
to the port and says so).  No reference source is ever committed.

Legs (each prints ONE JSON line; run as ``python -m oracle.ref_arm <leg> ...`` in a fresh process because the
Numba options are read when ``covasim`` is imported, reference settings.py:203-207, utils.py:26-34):

  full      the BASELINE workload from day 0 to the end in ``--chunks`` consecutive ``sim.run(until=...)`` calls
            (or as many days as fit ``--budget`` seconds), after a small warm-up sim that compiles the Numba kernels
  continue  the same sim continued from a People state saved by the GPU run (``--state file.npz``) for a bounded
            number of days: bench.py's cpu_baseline
'''
import argparse
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = '/root/reference'
REF_DST = os.path.join(HERE, '_ref')


def install(force=False):
    ''' Copy the reference package (unmodified) + the stand-in third-party modules into oracle/_ref; returns the path or None '''
    src = os.path.join(REF_SRC, 'covasim')
    if not os.path.isdir(src):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, 'covasim')) else None
    dst = os.path.join(REF_DST, 'covasim')
    if force or not os.path.isdir(dst):
        shutil.rmtree(REF_DST, ignore_errors=True)
        os.makedirs(REF_DST)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        for name in os.listdir(os.path.join(HERE, 'shim')):
            s, d = os.path.join(HERE, 'shim', name), os.path.join(REF_DST, name)
            if os.path.isdir(s):
                shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
            else:
                shutil.copy2(s, d)
        ex = os.path.join(REF_SRC, 'examples', 'example_data.csv')     # the data file BASELINE config 5 fits against
        if os.path.exists(ex):
            shutil.copy2(ex, os.path.join(REF_DST, 'example_data.csv'))
    return REF_DST


def available():
    return os.path.isdir(os.path.join(REF_DST, 'covasim')) or os.path.isdir(os.path.join(REF_SRC, 'covasim'))


def verify_unmodified():
    ''' In the build container: every .py of the install is byte-identical to /root/reference (returns the file count) '''
    import filecmp
    n = 0
    for dirpath, _, files in os.walk(os.path.join(REF_SRC, 'covasim')):
        for f in files:
            if f.endswith('.py'):
                a = os.path.join(dirpath, f)
                b = os.path.join(REF_DST, os.path.relpath(a, REF_SRC))
                assert filecmp.cmp(a, b, shallow=False), f'{b} differs from the reference'
                n += 1
    return n


def import_reference(numba_parallel='none', threads=None):
    ''' Import the installed reference with the given Numba options (must run before anything imported numba / covasim) '''
    assert 'covasim' not in sys.modules and 'numba' not in sys.modules, 'import_reference must run in a fresh process'
    os.environ['COVASIM_NUMBA_PARALLEL'] = str(numba_parallel)
    os.environ['COVASIM_VERBOSE'] = '0'
    if threads:
        os.environ['NUMBA_NUM_THREADS'] = str(int(threads))
    os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join('/tmp', f'numba_cache_ref_{numba_parallel}'))
    if os.path.isdir(os.path.join(REF_DST, 'covasim')):
        sys.path.insert(0, REF_DST)
    elif os.path.isdir(os.path.join(REF_SRC, 'covasim')):
        sys.path.insert(0, os.path.join(HERE, 'shim'))
        sys.path.insert(0, REF_SRC)
    else:
        raise RuntimeError('the reference is not installed (oracle/_ref is written by __graft_entry__.build() in the build container)')
    import covasim as cv
    import numba
    return cv, numba


def workload(cv, pop_size, n_days, seed=1):
    ''' The BASELINE C2 configuration (bench.py:workload_pars / workload_interventions) '''
    pars = dict(pop_size=pop_size, pop_type='hybrid', n_days=n_days, pop_infected=max(1, int(0.005 * pop_size)), rand_seed=seed, verbose=0)
    ivs = [cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), cv.contact_tracing(trace_probs=0.3, start_day=30)]
    return pars, ivs


def warm_up(cv, n_days):
    ''' Compile every Numba kernel of the path on a small sim of the same shape '''
    pars, ivs = workload(cv, 20000, min(n_days, 35))
    pars['pop_infected'] = 400
    ivs = [cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=2), cv.contact_tracing(trace_probs=0.3, start_day=4)]
    cv.Sim(pars, interventions=ivs).run()


def leg_full(args):
    threads = args.threads or os.cpu_count()
    cv, numba = import_reference(args.numba_parallel, threads)
    t0 = time.perf_counter()
    warm_up(cv, args.n_days)
    t_warm = time.perf_counter() - t0
    pars, ivs = workload(cv, args.pop_size, args.n_days)
    t0 = time.perf_counter()
    sim = cv.Sim(pars, interventions=ivs)
    sim.initialize()
    t_init = time.perf_counter() - t0
    npts = args.n_days + 1
    bounds = [round(npts * (i + 1) / args.chunks) for i in range(args.chunks)]
    chunk_s, chunk_days, done = [], [], 0
    t_start = time.perf_counter()
    for b in bounds:
        if b <= done:
            continue
        if args.budget and time.perf_counter() - t_start > args.budget:
            break
        ts = time.perf_counter()
        if b >= npts:
            sim.run(reset_seed=False) if done else sim.run()
        else:
            sim.run(until=b, reset_seed=bool(done == 0))
        chunk_s.append(time.perf_counter() - ts)
        chunk_days.append(b - done)
        done = b
    total = float(sum(chunk_s))
    out = dict(leg='full', kind='reference', version=cv.__version__, numba=numba.__version__, numba_parallel=args.numba_parallel,
               cores=int(numba.get_num_threads()) if args.numba_parallel != 'none' else 1, host_cores=os.cpu_count(),
               pop_size=args.pop_size, n_days=args.n_days, days_done=done, complete=bool(done == npts), seconds=total,
               agent_days_per_s=args.pop_size * done / total if total else 0.0, chunk_seconds=chunk_s, chunk_days=chunk_days,
               init_s=t_init, warmup_s=t_warm,
               cum_infections=float(sim.results['cum_infections'].values[done - 1]) if done else 0.0)
    print(json.dumps(out))


def leg_continue(args):
    import numpy as np
    threads = args.threads or os.cpu_count()
    cv, numba = import_reference(args.numba_parallel, threads)
    warm_up(cv, args.n_days)
    state = np.load(args.state)
    t0 = int(state['t'])
    pars, ivs = workload(cv, args.pop_size, args.n_days)
    lkeys = [str(k) for k in state['layer_keys']]
    contacts = cv.Contacts(layer_keys=lkeys)
    for lk in lkeys:
        contacts[lk] = cv.Layer(p1=state[f'p1/{lk}'], p2=state[f'p2/{lk}'], beta=state[f'beta/{lk}'], label=lk)
    sim = cv.Sim(pars, interventions=ivs)
    sim.popdict = dict(uid=np.arange(args.pop_size, dtype=np.int32), age=state['people/age'], sex=state['people/sex'], contacts=contacts, layer_keys=lkeys)
    sim.initialize()
    ppl = sim.people
    for key in ppl.meta.all_states:
        if f'people/{key}' in state.files:
            arr = state[f'people/{key}']
            assert ppl[key].shape == arr.shape, (key, ppl[key].shape, arr.shape)
            ppl[key] = arr.astype(ppl[key].dtype)
    sim.t = t0
    sim.step()                                           # one untimed day (page faults, allocator)
    days, t_start = 0, time.perf_counter()
    while time.perf_counter() - t_start < args.budget and sim.t < sim.npts - 1 and days < args.max_days:
        sim.step()
        days += 1
    el = time.perf_counter() - t_start
    out = dict(leg='continue', kind='reference', version=cv.__version__, numba=numba.__version__, numba_parallel=args.numba_parallel,
               cores=int(numba.get_num_threads()) if args.numba_parallel != 'none' else 1, host_cores=os.cpu_count(),
               pop_size=args.pop_size, first_day=t0 + 1, days=days, seconds=el, s_per_day=el / max(days, 1),
               agent_days_per_s=args.pop_size * days / el if el else 0.0,
               n_exposed=int(ppl.exposed.sum()), n_infectious=int(ppl.infectious.sum()))
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('leg', choices=['full', 'continue', 'install'])
    ap.add_argument('--pop-size', type=int, default=1_000_000)
    ap.add_argument('--n-days', type=int, default=180)
    ap.add_argument('--numba-parallel', default='full', choices=['none', 'safe', 'full'])
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--chunks', type=int, default=20)
    ap.add_argument('--budget', type=float, default=0.0, help='stop starting new chunks / days after this many seconds (0: none)')
    ap.add_argument('--max-days', type=int, default=40)
    ap.add_argument('--state', default=None)
    args = ap.parse_args()
    if args.leg == 'install':
        print(install(force=True))
    elif args.leg == 'full':
        leg_full(args)
    else:
        leg_continue(args)


if __name__ == '__main__':
    main()
