"""`-m gpu`: members of a group spread over shards (SURVEY §8e / BASELINE config 4 placement).

N engines in one process on one GPU, exchanging their cross-shard RPC buckets by device copies
(`LocalTransport`); the NCCL transport moves the same buckets with all_to_all_single and is
exercised by tests/test_sharded_nccl.py on a multi-GPU box.  The sharded run must leave every
member bit-identical to the UNSHARDED oracle run of the same groups.
"""
import pytest

from oracle_lib import Oracle
from ra_b200 import abi

pytestmark = pytest.mark.gpu

CASES = [
    # shards, members, local groups per slot, steps, election permille, commands per leader and step
    (2, 5, 300, 80, 10, 1),
    (4, 3, 128, 60, 0, 1),
    (8, 5, 96, 80, 20, 1),
    (3, 7, 100, 60, 10, 1),
    (5, 5, 200, 50, 10, 1),   # N == M: every member of a group on its own shard
    (8, 5, 256, 70, 10, 1),   # the shape of tests/test_sharded_nccl.py at world 8
    (8, 5, 64, 60, 10, 64),   # 64-entry pipelined AppendEntries, 8 shards
]


@pytest.mark.parametrize("transport", ["buckets", "peer"])
@pytest.mark.parametrize("n,m,gl,steps,permille,cmds", CASES)
def test_sharded_flood_equals_unsharded_oracle(n, m, gl, steps, permille, cmds, transport):
    from ra_b200.sharded import LocalPeerTransport, LocalTransport, Shard, ShardedFlood
    shards = [Shard(gl, m, n, k, buckets=(transport == "buckets")) for k in range(n)]
    fl = ShardedFlood(LocalTransport(shards) if transport == "buckets" else LocalPeerTransport(shards))
    fl.bootstrap()
    fl.run(steps, cmds, permille, seed=77)
    fl.sync()
    g = n * gl
    o = Oracle(g, m, route_on_device=True)
    o.reset_empty()
    o.step([abi.ev_simple(o.row_of(i, 0), abi.EV_ELECTION_TIMEOUT) for i in range(g)])
    o.flood(steps, cmds, permille, seed=77, threads=8)
    want = {r.row: r.key()[1:] for r in o.read_rows(range(o.n_rows))}
    seen = 0
    for s in shards:
        for r in s.eng.read_rows(range(s.eng.n_rows)):
            assert r.key()[1:] == want[s.global_row(r.row, g)], (s.shard, r.row)
            seen += 1
    assert seen == g * m
    co, ce = o.counters(), fl.counters()
    for k in ("events", "commits", "applied", "msgs_out", "msgs_dropped", "elections_won", "fatal_rows"):
        assert ce[k] == co[k], k
    assert co["commits"] > 0 and co["msgs_dropped"] == 0


def test_config4_at_stated_size():
    """BASELINE.json configs[3] at its stated size: 100,000 groups x 5 members spread over 8 shards (12,500 local
    groups per slot and shard), 64-entry commands -> pipelined 64-entry AppendEntries, peer-store transport.
    Counters equal the unsharded oracle's; every 13th row of every shard is diffed field by field."""
    from ra_b200.sharded import LocalPeerTransport, Shard, ShardedFlood
    n, m, gl, steps, permille, cmds = 8, 5, 12_500, 24, 10, 64
    shards = [Shard(gl, m, n, k, buckets=False) for k in range(n)]
    fl = ShardedFlood(LocalPeerTransport(shards))
    fl.bootstrap()
    fl.run(steps, cmds, permille, seed=0xA04)
    fl.sync()
    g = n * gl
    o = Oracle(g, m, route_on_device=True)
    o.reset_empty()
    o.step([abi.ev_simple(o.row_of(i, 0), abi.EV_ELECTION_TIMEOUT) for i in range(g)])
    o.flood(steps, cmds, permille, seed=0xA04, threads=8)
    co, ce = o.counters(), fl.counters()
    for k in ("events", "commits", "applied", "msgs_out", "msgs_dropped", "elections_won", "fatal_rows",
              "aer_received_follower", "aer_replies_success"):
        assert ce[k] == co[k], k
    assert co["commits"] >= g * 64 * (steps - 8) and co["msgs_dropped"] == 0
    checked = 0
    for s in shards:
        ids = list(range(0, s.eng.n_rows, 13))
        want = {r.row: r.key()[1:] for r in o.read_rows([s.global_row(i, g) for i in ids])}
        for r in s.eng.read_rows(ids):
            assert r.key()[1:] == want[s.global_row(r.row, g)], (s.shard, r.row)
            checked += 1
    assert checked > 30_000


def test_placement_math():
    from ra_b200.sharded import global_group, local_group, shard_of
    for n in (1, 2, 3, 8):
        for g in range(50):
            for s in range(5):
                k = shard_of(g, s, n)
                assert global_group(local_group(g, n), s, k, n) == g


def test_undersized_buckets_are_counted_not_silent():
    """The bucket transport has a capacity per destination and step; overflowing it is a counted
    drop (`msgs_dropped`, RA_ST_MSG_DROPPED), like the reference's nosuspend send -- never silent."""
    from ra_b200.sharded import LocalTransport, Shard, ShardedFlood
    n, m, gl = 8, 5, 256
    shards = [Shard(gl, m, n, k, cap=300) for k in range(n)]
    fl = ShardedFlood(LocalTransport(shards))
    fl.bootstrap()
    fl.run(40, 1, 10, seed=31)
    fl.sync()
    assert fl.counters()["msgs_dropped"] > 0
