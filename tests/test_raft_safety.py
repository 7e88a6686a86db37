"""Raft's safety properties (Ongaro & Ousterhout, fig. 3) checked on long closed-loop runs with message
loss and frequent elections -- independent of the oracle: parity tests show that engine and oracle
agree, these show that what they agree on is Raft.  The network here drops records but keeps each
sender -> receiver channel FIFO and free of duplicates, like the Erlang distribution ra runs on (ra_server
itself asserts on a re-ordered stale empty AppendEntries, `?assert` at ra_server.erl:1294-1303; the trace
parity tests do feed such input, and both backends flag the row fatal exactly like the reference would crash).

  Election Safety        at most one leader per term and group
  Log Matching           same (index, term) in two logs => identical up to that index
  Leader Completeness    an entry committed in some term is in the log of every later leader
  State Machine Safety   no two members apply different entries at one index; last_applied never goes back
(no forged RPCs here: Byzantine senders are outside Raft's fault model)."""
import pytest

import trace_gen
from ra_b200 import abi
from ra_suite import make_backend

BACKENDS = ["oracle", "emu", pytest.param("engine", marks=pytest.mark.gpu)]


def log_of(r: abi.RaRowState) -> dict:
    out = {}
    for k in range(r.n_runs):
        end = r.run_start[k + 1] - 1 if k + 1 < r.n_runs else r.last_index
        for i in range(r.run_start[k], end + 1):
            out[i] = r.run_term[k]
    return out


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("g,m,steps,seed,knobs", [
    (6, 3, 400, 1, dict(p_drop=0.08, p_timeout=0.03)),
    (4, 5, 500, 2, dict(p_drop=0.05, p_timeout=0.02, max_cmd=5)),
    (3, 7, 400, 3, dict(p_drop=0.1, p_timeout=0.04)),
])
def test_raft_safety_properties(backend, g, m, steps, seed, knobs):
    b = make_backend(backend, g, m)
    cl = trace_gen.Cluster(b, seed, p_adversarial=0.0, p_dup=0.0, p_delay=0.0, p_withhold_written=0.0, **knobs)
    leaders = {}                    # (group, term) -> slot
    committed = {}                  # group -> {index: term} as first seen committed on a leader
    applied_prev = {}
    n_leader_obs = n_commit = 0
    for t in range(steps):
        batch = cl.next_batch()
        msgs, notes = b.step(batch)
        cl.absorb(msgs, notes)
        if t % 4:
            continue
        rows = b.read_rows(range(b.n_rows))
        assert all(not (r.flags & 4) for r in rows), "a member went fatal without forged input"
        for grp in range(g):
            mem = [rows[s * g + grp] for s in range(m)]
            logs = [log_of(r) for r in mem]
            for r in mem:
                if r.role == abi.LEADER:                                           # Election Safety
                    n_leader_obs += 1
                    assert leaders.setdefault((grp, r.current_term), r.self_slot) == r.self_slot
                assert r.last_applied >= applied_prev.get(r.row, 0)               # applied never goes back
                applied_prev[r.row] = r.last_applied
            for a in range(m):                                                     # Log Matching
                for c in range(a + 1, m):
                    common = [i for i in logs[a] if i in logs[c] and logs[a][i] == logs[c][i]]
                    if common:
                        top = max(common)
                        lo = max(mem[a].first_index, mem[c].first_index)
                        for i in range(lo, top + 1):
                            assert logs[a].get(i) == logs[c].get(i), (grp, a, c, i)
            cm = committed.setdefault(grp, {})
            for r, lg in zip(mem, logs):
                if r.role == abi.LEADER:
                    for i, tm in cm.items():                                       # Leader Completeness
                        if i >= r.first_index and leaders.get((grp, r.current_term)) == r.self_slot and r.current_term >= tm:
                            assert lg.get(i) == tm, (grp, r.self_slot, i, tm, lg.get(i))
                    for i in range(max(r.first_index, 1), r.commit_index + 1):
                        if i in lg and i not in cm:
                            cm[i] = lg[i]
                            n_commit += 1
                for i in range(max(r.first_index, 1), min(r.last_applied, r.last_index) + 1):   # State Machine Safety
                    if i in cm and i in lg:
                        assert lg[i] == cm[i], (grp, r.self_slot, i)
    assert n_leader_obs > steps // 8 and n_commit > 20          # the run really elected leaders and committed
