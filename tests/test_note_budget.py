"""The note budget (include/ra_engine.h: RA_NOTE_CAP, RA_NOTE_RESERVE, RA_ST_NOTE_OVERFLOW).

Round-1 finding: note() dropped the 8th note of a row's step after the state had already moved, so a legal
batch of RA_LOCAL_CAP pipelined AppendEntries lost its last APPLY and the entries never reached
ra_machine:apply/3.  Now (a) the cap covers such batches, (b) a row stops TAKING events while fewer than
RA_NOTE_RESERVE slots are free -- unreached host events are reported unconsumed, unreached mailbox records
count as dropped -- and (c) a residual overflow inside one event stops the row (fatal) instead of losing a note.
Bodies run on the oracle, the host build of the device logic and, on a GPU, the CUDA engine.
"""
import pytest

from ra_suite import *  # noqa: F401,F403

BACKENDS = ["oracle", "emu", pytest.param("engine", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def be(request):
    return request.param


def _follower(n=3):
    st = empty_state(n, N2)
    st.current_term = 1
    st.leader_slot = N1
    return st


def _aer(prev, n, commit, term=1, frm=N1):
    return ev_aer(0, frm, term, prev, term if prev else 0, commit, [term] * n)


def _applied_ranges(notes):
    return [(n.a, n.b) for n in notes_of(notes, NOTE_APPLY)]


def test_four_pipelined_aers_in_one_batch(be):
    """4 AERs x 2 entries, leader_commit = prev + 1 each: WAL,APPLY x 4 = 8 notes; none may be lost."""
    nd = Node(be, 3, pure=False)
    st = _follower()
    nd.b.load_rows([st])
    evs = []
    for k in range(4):
        e = _aer(2 * k, 2, 2 * k + 1)
        e.row = st.row
        evs.append(e)
    msgs, notes = nd.b.step(evs)
    out = nd.b.read_rows([st.row])[0]
    assert out.last_index == 8 and out.commit_index == 7 and out.last_applied == 7
    wal = [(n.a, n.b, n.c) for n in notes_of(notes, NOTE_WAL_APPEND)]
    assert wal == [(1, 2, 1), (3, 4, 1), (5, 6, 1), (7, 8, 1)]
    # every index up to last_applied reaches ra_machine:apply exactly once, in order
    ap = _applied_ranges(notes)
    assert ap == [(1, 1), (2, 3), (4, 5), (6, 7)]
    assert not (status(notes) & ST_NOTE_OVERFLOW)
    assert out.flags & 4 == 0


def test_budget_refuses_host_events_and_reports_them(be):
    """note_cap 6 -> a row takes an event only while n_notes <= 1: the 2nd AER of the batch is left unconsumed."""
    nd = Node(be, 3, pure=False, note_cap=6)
    st = _follower()
    nd.b.load_rows([st])
    evs = []
    for k in range(3):
        e = _aer(2 * k, 2, 2 * k + 1)
        e.row = st.row
        evs.append(e)
    msgs, notes = nd.b.step(evs)
    out = nd.b.read_rows([st.row])[0]
    # only the first AER was evaluated (it left WAL + APPLY = 2 notes > budget for another event)
    assert out.last_index == 2 and out.commit_index == 1 and out.last_applied == 1
    (stn,) = notes_of(notes, NOTE_STATUS)
    assert stn.aux & ST_NOTE_OVERFLOW
    assert (stn.c >> 8) & 0xff == 2                    # the last two events of the run were not consumed
    assert stn.c & 0xff == 0 and out.flags & 4 == 0    # not fatal
    # the host submits them again: same result as an unbounded step
    msgs2, notes2 = nd.b.step(evs[1:2])
    msgs3, notes3 = nd.b.step(evs[2:3])
    out = nd.b.read_rows([st.row])[0]
    assert out.last_index == 6 and out.last_applied == 5
    assert _applied_ranges(notes) + _applied_ranges(notes2) + _applied_ranges(notes3) == [(1, 1), (2, 3), (4, 5)]


def test_budget_drops_unreached_mailbox_records(be):
    """routed mode, note_cap 6: a leader that finds two success replies in its mailbox takes the first (COMMIT +
    APPLY = 2 notes) and does not reach the second: it counts as a dropped record, exactly like a full
    transport.  Raft tolerates the loss: the flood keeps committing, nobody goes fatal, and the three
    backends agree on every row."""
    from oracle_lib import Oracle
    def run(name):
        b = make_backend(name, 4, 3, route_on_device=True, note_cap=6)
        b.reset_empty()
        b.step([ev_simple(b.row_of(g, 0), EV_ELECTION_TIMEOUT) for g in range(4)])
        if name == "oracle":
            b.flood(60, 1, 0, seed=5, threads=1)
        else:
            b.flood(60, 1, 0, seed=5)
        return b
    b = run(be)
    c = b.counters()
    assert c["msgs_dropped"] > 0 and c["fatal_rows"] == 0
    assert c["commits"] >= 4 * 20
    rows = [r.key() for r in b.read_rows(range(b.n_rows))]
    if be != "oracle":
        o = run("oracle")
        assert rows == [r.key() for r in o.read_rows(range(o.n_rows))]
        assert c == o.counters()


def test_residual_overflow_stops_the_row(be):
    """One event whose notes exceed the reserve (a SEND_SNAPSHOT per peer): the row goes fatal with
    RA_FATAL_NOTE_OVERFLOW instead of silently losing a note."""
    n = 8
    nd = Node(be, n, pure=False, note_cap=6)
    st = base_state(n)
    st.role = LEADER
    install_snapshot(st, 3, 5)                           # every peer needs entries below the snapshot
    st.commit_index = 3
    st.last_applied = 3
    for p in range(n):
        st.peers[p].next_index = 2
        st.peers[p].match_index = 0
    role, out, msgs, notes = nd.handle(LEADER, ev_simple(0, EV_PIPELINE_RPCS), st)
    (stn,) = notes_of(notes, NOTE_STATUS)
    assert stn.aux & ST_NOTE_OVERFLOW and stn.aux & ST_FATAL
    assert stn.c & 0xff == FATAL_NOTE_OVERFLOW
    assert out.flags & 4
