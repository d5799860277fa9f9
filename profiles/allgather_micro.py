'''
Floor of the per-day code exchange of agent-partitioned runs: back-to-back all-gathers of `--bytes` per rank, timed with CUDA events
(no compute in between, so no rank skew).   torchrun --nproc-per-node N profiles/allgather_micro.py --bytes 2000000
'''
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--bytes', type=int, default=2_000_000)
ap.add_argument('--iters', type=int, default=300)
args = ap.parse_args()
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
src = torch.full((args.bytes,), rank, dtype=torch.uint8, device='cuda')
dst = torch.empty(args.bytes * world, dtype=torch.uint8, device='cuda')
for _ in range(20):
    dist.all_gather_into_tensor(dst, src)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.iters):
    dist.all_gather_into_tensor(dst, src)
b.record()
torch.cuda.synchronize()
us = 1e3 * a.elapsed_time(b) / args.iters
t = torch.tensor([us], device='cuda', dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(world=world, bytes_per_rank=args.bytes, us_per_allgather=float(t.item()), env={k: v for k, v in os.environ.items() if k.startswith('NCCL_')})), flush=True)
dist.destroy_process_group()
