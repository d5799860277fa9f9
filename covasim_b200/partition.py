'''
Agent partition of ONE large simulation over several GPUs (BASELINE config 4).

The reference runs a simulation in one process on one core (sim.py:558-685).  Here a population too large or
too slow for one GPU is split into contiguous agent ranges, one per rank.  Each rank

* holds the People arrays of its own agents only;
* evaluates the transmissions (and contact-tracing notifications) whose TARGET it owns, from an adjacency
  with one row per GLOBAL source holding the edges that end in a local target;
* needs one byte per agent per day about the rest of the population (``transmit_code``: variant, symptomatic,
  isolated, quarantined, early viral load, breakthrough), exchanged with ONE fixed-size all-gather
  (``ncclAllGather`` over NVLink through torch.distributed), and a second all-gather of a 1-bit-per-agent
  case bitmap on days a contact_tracing intervention is active.

Because every random draw is a pure function of (seed, purpose, day, GLOBAL agent or edge id), a partitioned
run reproduces the single-GPU run of the same simulation bit for bit (tests/test_gpu_partition.py), for any
number of ranks.

Two communicators implement the exchanges: ``DistComm`` (torch.distributed: NCCL on GPUs, gloo in the CPU
tests of the host logic) and ``LocalComm`` (several ranks of one process on ONE GPU, one thread per rank --
how the single-GPU test box exercises the partitioned kernels).
'''
import threading

import numpy as np
import torch

__all__ = ['plan', 'build_partition_adjacency', 'build_partition_adjacency_chunks', 'DistComm', 'LocalComm', 'run_local']


def plan(n_global, world):
    ''' Chunk size (a multiple of 32) and the agent range [lo, hi) of every rank; every rank but the last owns exactly ``chunk`` agents '''
    n_global, world = int(n_global), int(world)
    chunk = -(-n_global // world)
    chunk = -(-chunk // 32) * 32
    ranges = [(min(r * chunk, n_global), min((r + 1) * chunk, n_global)) for r in range(world)]
    if any(hi <= lo for lo, hi in ranges):
        raise ValueError(f'{n_global} agents cannot be split over {world} ranks in chunks of {chunk}: a rank would own no agents')
    return chunk, ranges


def build_partition_adjacency(layers, layer_ids, lo, hi, n_slots, device):
    '''
    Rows = GLOBAL source id (``n_slots + 1`` row pointers), entries = the edges that end in a LOCAL target
    ``lo <= target < hi`` as int32[M, 4] = (target - lo, edge index within its layer, (layer << 1) | direction,
    beta bits) -- the entry format of ``cvb_bind_adjacency`` (direction 0: source is the edge's p1).
    ``layers`` is a list of dicts of 1-D tensors / arrays (p1, p2, beta) holding the WHOLE population's edges.
    Pure torch, device independent (the CPU tests check it against a brute-force loop).
    '''
    def chunks():
        for l, layer in zip(layer_ids, layers):
            yield l, torch.as_tensor(layer['p1']), torch.as_tensor(layer['p2']), torch.as_tensor(layer['beta']), 0
    return build_partition_adjacency_chunks(chunks(), lo, hi, n_slots, device)


def build_partition_adjacency_chunks(chunks, lo, hi, n_slots, device):
    '''
    The same from a stream of edge chunks ``(layer id, p1, p2, beta or None, index of the chunk's first edge within its
    layer)`` -- what the device-side population generator yields -- so that the whole population's edge lists never have
    to exist at once: only the entries with a local target are kept.
    '''
    dev = torch.device(device)
    src, cols = [], []
    for l, p1, p2, beta, e0 in chunks:
        p1 = p1.to(dev, torch.int64)
        p2 = p2.to(dev, torch.int64)
        E = p1.numel()
        if e0 + E >= 2 ** 31:
            raise ValueError('a layer with 2^31 or more edges cannot be indexed by the adjacency')
        wbits = (torch.ones(E, dtype=torch.float32, device=dev) if beta is None else beta.to(dev, torch.float32)).view(torch.int32)
        e = torch.arange(e0, e0 + E, dtype=torch.int32, device=dev)
        for d, (a, b) in enumerate(((p1, p2), (p2, p1))):
            keep = (b >= lo) & (b < hi)
            k = int(keep.sum())
            ent = torch.empty((k, 4), dtype=torch.int32, device=dev)
            ent[:, 0] = (b[keep] - lo).to(torch.int32)
            ent[:, 1] = e[keep]
            ent[:, 2] = (l << 1) | d
            ent[:, 3] = wbits[keep]
            src.append(a[keep])
            cols.append(ent)
    src = torch.cat(src) if src else torch.zeros(0, dtype=torch.int64, device=dev)
    M = int(src.numel())
    adj = torch.empty((max(M, 1), 4), dtype=torch.int32, device=dev)
    ptr = torch.zeros(n_slots + 1, dtype=torch.int64, device=dev)
    if M:
        # rows in ascending source order; within a row the order of (layer, direction, edge) as streamed (stable sort)
        order = torch.sort(src, stable=True).indices
        torch.cumsum(torch.bincount(src, minlength=n_slots), 0, out=ptr[1:])
        del src
        ent = torch.cat(cols)
        del cols
        torch.index_select(ent, 0, order, out=adj[:M])
    return ptr, adj, M


class DistComm:
    ''' The exchanges of a partitioned run over torch.distributed (one process per GPU) '''

    def __init__(self, group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('partition=True needs an initialised torch.distributed process group (torchrun)')
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, out, inp):
        ''' out[r * len(inp) : (r + 1) * len(inp)] = rank r's inp (one ncclAllGather; stream-ordered, no host sync) '''
        self.dist.all_gather_into_tensor(out, inp, group=self.group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def peer_exchange(self, device, sections):
        '''
        A PeerExchange over this group, or None when peer memory is not available (another backend, several nodes, a driver / container
        that does not allow the handle exchange) or switched off (COVASIM_B200_PEER_EXCHANGE=0): every rank must succeed, else all fall
        back to the library all-gather.
        '''
        import os
        ex, err = None, None
        if os.environ.get('COVASIM_B200_PEER_EXCHANGE', '1') != '0' and self.world <= 16 and self.dist.get_backend(self.group) == 'nccl':
            try:
                ex = PeerExchange(self, device, sections)
            except Exception as e:          # noqa: BLE001 -- any failure means "not available here"
                err = f'{type(e).__name__}: {e}'
        ok = self.gather_objects(ex is not None)
        if not all(ok):
            if ex is not None or err:
                import warnings
                warnings.warn(f'peer-memory exchange not available on every rank ({err or "another rank failed"}); using ncclAllGather')
            return None
        return ex


class PeerExchange:
    '''
    The per-day exchanges of a partitioned run over PEER MEMORY instead of a library collective: the ranks of one NVLink / NVSwitch node
    map one exchange buffer each into every other rank's address space (torch symmetric memory), every rank stores its chunk straight
    into all of them (one kernel, cvb_peer_push: the remote stores travel over NVLink) and a signal barrier separates the stores from the
    reads.  Two buffers per exchange alternate, so one barrier per exchange suffices: the buffer a rank overwrites at exchange k + 2 was
    last read before that rank's peers entered the barrier of exchange k + 1.  ``sections`` maps a name to the bytes one rank contributes.
    Measured on 4 B200s, 2 MB per rank: ncclAllGather 50 us (LL128: 29 us) -- profiles/allgather_micro.py; this exchange: see
    profiles/r2/README.md.
    '''

    def __init__(self, comm, device, sections, timeout_ms=60000):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.comm, self.world, self.rank = comm, comm.world, comm.rank
        self.timeout_ms = int(timeout_ms)
        self.offsets, self.sizes, self.parity = {}, {}, {}
        total = 0
        for name, nbytes in sections.items():
            span = (self.world * int(nbytes) + 255) // 256 * 256
            self.offsets[name] = (total, total + span)
            self.sizes[name] = int(nbytes)
            self.parity[name] = 0
            total += 2 * span
        group = comm.group if comm.group is not None else comm.dist.group.WORLD
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(ptrs) != self.world or not all(ptrs):
            raise RuntimeError('symmetric memory rendezvous returned no peer pointers')
        self.ptrs = (C.c_uint64 * self.world)(*ptrs)
        self.buf.zero_()
        self.hdl.barrier(channel=0, timeout_ms=self.timeout_ms)

    def all_gather(self, name, inp, stream_ptr=None):
        ''' Every rank's ``inp`` side by side in rank order; returns the device address of the filled buffer on this rank '''
        from . import _capi
        half = self.parity[name]
        self.parity[name] = half ^ 1
        nbytes = self.sizes[name]
        if inp.numel() * inp.element_size() != nbytes:
            raise ValueError(f'exchange "{name}" was sized for {nbytes} bytes per rank, got {inp.numel() * inp.element_size()}')
        off = self.offsets[name][half]
        _capi.call('cvb_peer_push', inp.data_ptr(), nbytes, self.ptrs, self.world, off + self.rank * nbytes, stream_ptr)
        self.hdl.barrier(channel=0, timeout_ms=self.timeout_ms)
        return self.buf.data_ptr() + off

    def view(self, name, half, dtype):
        off, n = self.offsets[name][half], self.world * self.sizes[name]
        return self.buf[off:off + n].view(dtype)


class _LocalShared:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class LocalComm:
    '''
    Several ranks inside one process on one GPU (one thread per rank, see ``run_local``).  Every rank issues its
    kernels on the same (legacy default) stream, so copies issued between two thread barriers are ordered against
    every rank's kernels exactly as an all-gather on that stream would be.
    '''

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    @staticmethod
    def make(world):
        shared = _LocalShared(world)
        return [LocalComm(shared, r) for r in range(world)]

    def _exchange(self, value):
        s = self.shared
        s.slots[self.rank] = value
        s.barrier.wait()
        vals = list(s.slots)
        s.barrier.wait()
        return vals

    def all_gather(self, out, inp):
        parts = self._exchange(inp)
        n = inp.numel()
        for r, p in enumerate(parts):
            out[r * n:(r + 1) * n].copy_(p)
        self.shared.barrier.wait()          # nobody overwrites its input before every rank has issued its copies

    def all_reduce_sum(self, t):
        parts = self._exchange(t.clone())
        t.zero_()
        for p in parts:
            t.add_(p.to(t.device))
        self.shared.barrier.wait()

    def gather_objects(self, obj):
        return self._exchange(obj)


def run_local(sims, fn=None):
    '''
    Run the ranks of a LocalComm-partitioned simulation, one thread per rank; ``fn(sim)`` defaults to
    ``sim.run()``.  Returns the list of results; re-raises the first exception of any rank.
    '''
    fn = fn or (lambda sim: sim.run())
    out, errs = [None] * len(sims), []

    def work(k):
        try:
            if torch.cuda.is_available():
                torch.cuda.set_device(sims[k].device)
            out[k] = fn(sims[k])
        except BaseException as e:          # noqa: BLE001 -- reported to the caller below
            errs.append(e)
            try:
                sims[k]._comm.shared.barrier.abort()
            except Exception:
                pass
    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(sims))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errs:
        real = [e for e in errs if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or errs)[0]
    return out
