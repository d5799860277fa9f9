'''
Agent-partitioned run under torchrun (one process per GPU, NCCL) against the single-GPU run of the same simulation:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_partition.py

Every rank runs its share of the partitioned simulation; rank 0 also runs the whole simulation on its GPU and checks
that the People arrays (gathered over ranks), every result series and the infection log are identical.  Not collected
by pytest (needs N GPUs); tests/test_gpu_partition.py makes the same comparison on one GPU with in-process ranks.
'''
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import covasim_b200 as cv
    import scenarios
    from test_gpu_partition import PART_SCENARIOS
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    names = sys.argv[1:] or ['hybrid3k', 'variants4k', 'odd5003', 'big200k']
    specs = dict(PART_SCENARIOS)
    specs['big200k'] = dict(pars=dict(pop_size=200_000, pop_infected=1000, pop_type='hybrid', n_days=60, verbose=0, rand_seed=1),
                            interventions=[('test_prob', dict(symp_prob=0.1, asymp_prob=0.01, start_day=10)),
                                           ('contact_tracing', dict(trace_probs=0.3, start_day=15))])
    for name in names:
        spec = specs[name]
        sim = cv.Sim(**scenarios.build(cv, spec), partition=True, pop_exact=False)
        sim.initialize()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        sim.run()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        log = sim.infection_log
        gathered = {}
        for k in sim.people.keys():
            gathered[k] = sim._comm.gather_objects(sim.people.to_numpy(k))
        if rank == 0:
            ref = cv.Sim(**scenarios.build(cv, spec), pop_exact=False)
            ref.run()
            for k in ref.people.keys():
                whole = ref.people.to_numpy(k)
                parts = np.concatenate(gathered[k], axis=-1)
                assert np.array_equal(whole, parts, equal_nan=(whole.dtype.kind == 'f')), f'{name}: People.{k} differs'
            for k in ref.result_keys():
                assert np.allclose(sim.results[k].values, ref.results[k].values, rtol=1e-6 if k == 'r_eff' else 1e-12, atol=0, equal_nan=True), f'{name}: result {k} differs'
            for k in ref.result_keys('variant'):
                assert np.array_equal(sim.results['variant'][k].values, ref.results['variant'][k].values), f'{name}: variant/{k} differs'
            want = ref.infection_log
            for k in ('source', 'target', 'date', 'layer', 'variant'):
                assert np.array_equal(log[k], want[k]), f'{name}: infection log {k} differs'
            print(f'OK {name} [exchange: {"peer memory" if sim._peer is not None else "ncclAllGather"}]: {world} ranks == single GPU (N={ref.n}, cum_infections={ref.summary["cum_infections"]:.0f}, partitioned run {el:.2f} s)', flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
