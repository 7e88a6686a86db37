import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle_lib import Oracle
from ra_b200 import abi
from ra_b200.sharded import LocalPeerTransport, LocalTransport, Shard, ShardedFlood
n, m, gl, permille = 2, 5, 300, 10
for steps in (0, 1, 2, 3, 5, 8, 12, 20):
    shards = [Shard(gl, m, n, k, buckets=False) for k in range(n)]
    fl = ShardedFlood(LocalPeerTransport(shards)); fl.bootstrap(); fl.run(steps, 1, permille, seed=77); fl.sync()
    g = n * gl
    o = Oracle(g, m, route_on_device=True); o.reset_empty()
    o.step([abi.ev_simple(o.row_of(i, 0), abi.EV_ELECTION_TIMEOUT) for i in range(g)])
    o.flood(steps, 1, permille, seed=77, threads=1)
    want = {r.row: r for r in o.read_rows(range(o.n_rows))}
    bad = 0
    for s in shards:
        for r in s.eng.read_rows(range(s.eng.n_rows)):
            w = want[s.global_row(r.row, g)]
            if r.key()[1:] != w.key()[1:]:
                if bad < 3:
                    print("steps", steps, "shard", s.shard, "local row", r.row, "slot", r.self_slot)
                    print("  got ", r.key()[1:])
                    print("  want", w.key()[1:])
                bad += 1
    print("steps", steps, "bad rows", bad, fl.counters()["events"], o.counters()["events"])
    if bad: break
