"""Members of a group on different GPUs: the cross-shard RPC router.

Placement: member (group g, slot s) lives on shard (g + s) mod N, at local group index g div N.
A record from slot s to slot t of the same group therefore always goes to shard
(shard + t - s) mod N and to the same local row index there, so routing needs no table.

Per step and shard:  raft_step kernels  ->  records for other shards land in dense per-destination
buckets in HBM (ra_engine_set_outbox)  ->  ONE all-to-all of the bucket counts and ONE all-to-all
of the buckets (NCCL over NVLink/NVSwitch, `torch.distributed.all_to_all_single`)  ->
ra_engine_deliver scatters the received records into the mailboxes of the next step.
No other collective exists on the data path.  Everything is enqueued on one CUDA stream per
shard (the engine is switched to torch's current stream), nothing synchronises with the host.

`LocalTransport` runs all shards in one process on one device (plain device copies instead of
NCCL): the single-GPU parity tests use it, so the sharded kernels and ABI are covered without
a second GPU.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import abi
from .engine import Engine, lib

import ctypes as C


def shard_of(group: int, slot: int, n_shards: int) -> int:
    return (group + slot) % n_shards


def local_group(group: int, n_shards: int) -> int:
    return group // n_shards


def global_group(local_q: int, slot: int, shard: int, n_shards: int) -> int:
    return n_shards * local_q + ((shard - slot) % n_shards)


def _declare(l):
    l.ra_engine_set_stream.restype = C.c_int
    l.ra_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    l.ra_engine_set_outbox.restype = C.c_int
    l.ra_engine_set_outbox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    l.ra_engine_deliver.restype = C.c_int
    l.ra_engine_deliver.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]


class Shard:
    """One shard = one engine + its exchange buffers (torch tensors on the engine's device)."""

    def __init__(self, groups_local: int, members: int, n_shards: int, shard: int, device: int = 0,
                 cap: int | None = None, **kw):
        _declare(lib())
        self.n_shards, self.shard, self.dev = n_shards, shard, device
        self.eng = Engine(groups_local, members, device=device, route_on_device=True, n_shards=n_shards,
                          shard=shard, **kw)
        rows = groups_local * members
        # per destination and step: ~3.1 records per row on average in a flood, spread over N shards
        self.cap = cap or max(1024, (rows * 5) // n_shards)
        tdev = torch.device("cuda", device)
        self.outbox = torch.zeros((n_shards, self.cap, 64), dtype=torch.uint8, device=tdev)
        self.inbox = torch.zeros((n_shards, self.cap, 64), dtype=torch.uint8, device=tdev)
        self.out_cnt = torch.zeros(n_shards, dtype=torch.int32, device=tdev)
        self.in_cnt = torch.zeros(n_shards, dtype=torch.int32, device=tdev)
        with torch.cuda.device(tdev):
            stream = torch.cuda.current_stream(tdev).cuda_stream
        e = self.eng
        e._check(lib().ra_engine_set_stream(e._h, C.c_void_p(stream)), "set_stream")
        e._check(lib().ra_engine_set_outbox(e._h, C.c_void_p(self.outbox.data_ptr()),
                                            C.c_void_p(self.out_cnt.data_ptr()), self.cap), "set_outbox")

    def deliver(self) -> None:
        e = self.eng
        e._check(lib().ra_engine_deliver(e._h, C.c_void_p(self.inbox.data_ptr()),
                                         C.c_void_p(self.in_cnt.data_ptr()), self.cap), "deliver")

    def bootstrap_events(self) -> List[abi.RaEvent]:
        """election_timeout for the slot-0 member of every group that lives here."""
        return [abi.ev_simple(self.eng.row_of(q, 0), abi.EV_ELECTION_TIMEOUT) for q in range(self.eng.n_groups)]

    def global_row(self, local_row: int, total_groups: int) -> int:
        slot, q = divmod(local_row, self.eng.n_groups)
        return slot * total_groups + global_group(q, slot, self.shard, self.n_shards)


class LocalTransport:
    """All shards in this process, same device: the exchange is device-to-device copies."""

    def __init__(self, shards: Sequence[Shard]):
        self.shards = list(shards)

    def exchange(self) -> None:
        for b in self.shards:
            for a in self.shards:
                if a is b:
                    b.in_cnt[a.shard] = 0
                    continue
                b.inbox[a.shard].copy_(a.outbox[b.shard])
                b.in_cnt[a.shard] = a.out_cnt[b.shard]
        for s in self.shards:
            s.deliver()


class NcclTransport:
    """One shard per rank: two all_to_all_single calls per step (counts, then equal-size buckets)."""

    def __init__(self, shard: Shard):
        import torch.distributed as dist
        self.dist = dist
        self.shard = shard
        self.shards = [shard]

    def exchange(self) -> None:
        s = self.shard
        self.dist.all_to_all_single(s.in_cnt, s.out_cnt)
        self.dist.all_to_all_single(s.inbox.view(s.n_shards, -1), s.outbox.view(s.n_shards, -1))
        s.deliver()


class ShardedFlood:
    """The flood of ra_engine_flood over sharded members: step, exchange, deliver, repeat."""

    def __init__(self, transport):
        self.t = transport

    def bootstrap(self) -> None:
        for s in self.t.shards:
            s.eng.reset_empty()
            s.eng.step(s.bootstrap_events())
        self.t.exchange()

    def run(self, n_steps: int, cmds: int = 1, permille: int = 0, seed: int = 1) -> None:
        for _ in range(n_steps):
            for s in self.t.shards:
                s.eng.flood(1, cmds, permille, seed, sync=False)
            self.t.exchange()

    def sync(self) -> None:
        for s in self.t.shards:
            torch.cuda.synchronize(s.dev)

    def counters(self) -> dict:
        tot: dict = {}
        for s in self.t.shards:
            for k, v in s.eng.counters().items():
                tot[k] = tot.get(k, 0) + v
        return tot
