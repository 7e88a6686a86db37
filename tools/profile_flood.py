"""Small driver for ncu: settle a flood, then run a few steps (the launches ncu captures)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_b200 import abi
from ra_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--groups", type=int, default=100_000)
ap.add_argument("--members", type=int, default=5)
ap.add_argument("--settle", type=int, default=40)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--permille", type=int, default=10)
a = ap.parse_args()
e = Engine(a.groups, a.members, route_on_device=True)
e.reset_empty()
e.step([abi.ev_simple(e.row_of(g, 0), abi.EV_ELECTION_TIMEOUT) for g in range(a.groups)])
e.flood(a.settle, 1, a.permille, seed=0xA00)
e.flood(a.steps, 1, a.permille, seed=0xA00)
print(e.counters(), e.last_kernel_ms())
names = ["none", "AER", "AER_REPLY", "REQ_VOTE", "REQ_VOTE_RES", "PRE_VOTE", "PRE_VOTE_RES", "WRITTEN", "COMMAND",
         "ELECTION_TMO", "AWAIT_TMO", "PIPELINE", "TICK"]
roles = ["follower", "candidate", "pre_vote", "leader", "await_condition"]
h = e.stall_histogram()
tot = sum(h.values())
print("events that left the fast kernel: %d of %d (%.2f%%)" % (tot, e.counters()["events"], 100.0 * tot / max(1, e.counters()["events"])))
for (ro, ty), v in sorted(h.items(), key=lambda kv: -kv[1]):
    print("  %-16s %-14s %10d" % (roles[ro] if ro < 5 else ro, names[ty] if ty < 13 else ty, v))
