"""Golden vectors: the reference's own pure-function tests of the hot path, restated.

Each test cites the case it restates in /root/reference/test/ra_server_SUITE.erl (or the
in-module eunit of src/ra_server.erl).  The same bodies run against the CPU oracle (pins
the oracle, `-m "not gpu"`), against the engine's device logic compiled for the host (`emu`,
tests/emu/, also CPU) and against the CUDA engine through the C ABI (`-m gpu`).
Assertions on machine_state / payloads are dropped: ra_machine:apply/3 stays on the host.
"""
import pytest

from ra_suite import *  # noqa: F401,F403

BACKENDS = ["oracle", "emu", pytest.param("engine", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def be(request):
    return request.param


def test_agreed_commit():
    """agreed_commit_test, src/ra_server.erl:4198-4211."""
    from oracle_lib import agreed_commit
    assert agreed_commit([4]) == 4
    assert agreed_commit([4, 3]) == 3
    assert agreed_commit([4, 4, 4]) == 4
    assert agreed_commit([4, 4, 3]) == 4
    assert agreed_commit([3, 4, 4]) == 4
    assert agreed_commit([4, 2, 3]) == 3


def test_election_timeout(be):
    """election_timeout/1, :333-379."""
    nd = Node(be, 3)
    st = base_state(3)
    ev = ev_simple(0, EV_ELECTION_TIMEOUT)
    # follower
    role, s, msgs, notes = nd.handle_follower(ev, st)
    assert role == PRE_VOTE and s.current_term == 5 and s.votes == 0
    token = s.pre_vote_token
    (ne,) = next_events(msgs)
    assert ne.type == EV_PRE_VOTE_RES and ne.term == 5 and ne.c == token and ne.d == 1
    reqs = of_type(msgs, EV_PRE_VOTE)
    assert [r.row for r in reqs] == [N2, N3]
    assert all(r.term == 5 and r.c == token and r.a == 3 and r.b == 5 and r.from_slot == N1 for r in reqs)
    # non-voters ignore election_timeout
    nv = clone(st)
    nv.membership = PROMOTABLE
    role, s2, msgs2, _ = nd.handle_follower(ev, nv)
    assert role == FOLLOWER and msgs2 == [] and s2.key() == nv.key()
    role, s2, msgs2, _ = nd.handle_await_condition(ev, nv)
    assert role == AWAIT_CONDITION and msgs2 == []
    # pre_vote
    role, s3, msgs3, _ = nd.handle_pre_vote(ev, s)
    assert role == PRE_VOTE and s3.current_term == 5 and s3.votes == 0
    assert s3.pre_vote_token != token                      # tokens are not the same
    # candidate
    role, s4, msgs4, _ = nd.handle_candidate(ev, st)
    assert role == CANDIDATE and s4.current_term == 6 and s4.votes == 0
    (ne,) = next_events(msgs4)
    assert ne.type == EV_REQUEST_VOTE_RES and ne.term == 6 and ne.d == 1
    reqs = of_type(msgs4, EV_REQUEST_VOTE)
    assert [(r.row, r.term, r.a, r.b, r.from_slot) for r in reqs] == [(N2, 6, 3, 5, N1), (N3, 6, 3, 5, N1)]


def test_follower_aer_1(be):
    """follower_aer_1/1, :381-451."""
    nd = Node(be, 3)
    init = empty_state(3, N1)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, 0, [1]), init)
    assert (role, s1.leader_slot, s1.current_term, s1.commit_index, s1.last_applied) == (FOLLOWER, N1, 1, 0, 0)
    role, s2, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 1, 1, 1, [1]), s1)
    assert (role, s2.commit_index, s2.last_applied) == (FOLLOWER, 1, 1)
    role, s3, msgs, _ = nd.handle_follower(ev_written(0, 1, 1, 1), s2)
    (r,) = msgs
    assert reply_fields(r) == dict(to=N1, from_=N1, term=1, success=True, next_index=3, last_index=1, last_term=1)
    assert (s3.commit_index, s3.last_applied) == (1, 1)
    role, s4, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 2, 1, 3, [1]), s3)
    assert (role, s4.commit_index, s4.last_applied) == (FOLLOWER, 3, 3)
    role, s5, msgs, _ = nd.handle_follower(ev_written(0, 1, 2, 2), s4)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["next_index"], f["last_term"], f["last_index"]) == (4, 1, 2)
    # empty AER before {written, 3}
    role, s6, msgs, _ = nd.handle_follower(ev_aer(0, N1, 1, 3, 1, 3, []), s5)
    f = reply_fields(msgs[0])
    assert (role, f["next_index"], f["last_term"], f["last_index"]) == (FOLLOWER, 4, 1, 2)
    assert (s6.commit_index, s6.last_applied) == (3, 3)
    role, s7, msgs, _ = nd.handle_follower(ev_written(0, 1, 3, 3), s6)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["next_index"], f["last_term"], f["last_index"]) == (4, 1, 3)


def test_follower_aer_2(be):
    """follower_aer_2/1, :453-485."""
    nd = Node(be, 3)
    init = empty_state(3, N2)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, 0, [1]), init)
    assert (s1.leader_slot, s1.current_term, s1.commit_index, s1.last_applied) == (N1, 1, 0, 0)
    role, s2, msgs, _ = nd.handle_follower(ev_written(0, 1, 1, 1), s1)
    (r,) = msgs
    assert reply_fields(r) == dict(to=N1, from_=N2, term=1, success=True, next_index=2, last_index=1, last_term=1)
    assert (s2.commit_index, s2.last_applied) == (0, 0)
    role, s3, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 1, 1, 1, []), s2)
    assert (role, s3.commit_index, s3.last_applied) == (FOLLOWER, 1, 1)


def test_follower_aer_3(be):
    """follower_aer_3/1, :487-557: gap -> await_condition, catch-up batch, resend."""
    nd = Node(be, 3)
    init = empty_state(3, N2)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, 1, [1]), init)
    assert (s1.commit_index, s1.last_applied) == (1, 1)
    role, s2, msgs, _ = nd.handle_follower(ev_written(0, 1, 1, 1), s1)
    f = reply_fields(msgs[0])
    assert (f["next_index"], f["last_index"], f["last_term"]) == (2, 1, 1)
    # AER with index [3] -> missing
    role, s3, msgs, notes = nd.handle_follower(ev_aer(0, N1, 1, 2, 1, 3, [1]), s2)
    assert role == AWAIT_CONDITION
    (r,) = msgs
    assert reply_fields(r) == dict(to=N1, from_=N2, term=1, success=False, next_index=2, last_index=1, last_term=1)
    assert status(notes) & ST_LEADER_MSG                          # {record_leader_msg, N1}
    assert (s3.commit_index, s3.last_applied) == (1, 1)
    # AER with index [2,3,4], commit_index = 3
    aer3 = ev_aer(0, N1, 1, 1, 1, 3, [1, 1, 1])
    role, s3b, msgs, _ = nd.handle_await_condition(aer3, s3)
    assert role == FOLLOWER and len(next_events(msgs)) == 1       # {next_event, AER3}
    role, s4, _, _ = nd.handle_follower(aer3, s3b)
    assert (role, s4.commit_index, s4.last_applied) == (FOLLOWER, 3, 3)
    role, s5, msgs, _ = nd.handle_follower(ev_written(0, 1, 4, 4), s4)
    f = reply_fields(msgs[0])
    assert (f["next_index"], f["success"], f["last_term"], f["last_index"]) == (5, True, 1, 4)
    # resend of [2,3,4] with commit_index = 4
    role, s6, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 1, 1, 4, [1, 1, 1]), s5)
    assert (role, s6.commit_index, s6.last_applied) == (FOLLOWER, 4, 4)


def test_follower_aer_4(be):
    """follower_aer_4/1, :559-584: commit_index := LeaderCommit with no min."""
    nd = Node(be, 3)
    init = empty_state(3, N2)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, 10, [1, 1, 1, 1]), init)
    assert (s1.commit_index, s1.last_applied) == (10, 4)
    role, s2, msgs, _ = nd.handle_follower(ev_written(0, 1, 4, 4), s1)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["next_index"], f["last_term"], f["last_index"]) == (5, 1, 4)
    assert (s2.commit_index, s2.last_applied) == (10, 4)


@pytest.mark.parametrize("leader_commit1", [2, 3])
def test_follower_aer_5_6(be, leader_commit1):
    """follower_aer_5/1 :586-617, follower_aer_6/1 :619-653: new-term empty AER truncates."""
    nd = Node(be, 3)
    init = empty_state(3, N2)
    _, s00, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, leader_commit1, [1, 1, 1, 1]), init)
    _, s0, _, _ = nd.handle_follower(ev_written(0, 1, 4, 4), s00)
    assert s0.last_applied == leader_commit1
    role, s1, msgs, notes = nd.handle_follower(ev_aer(0, N5, 2, 3, 1, 3, []), s0)
    assert role == FOLLOWER
    f = reply_fields(msgs[0])
    assert (f["to"], f["next_index"], f["last_term"], f["last_index"]) == (N5, 4, 1, 3)
    assert s1.last_index == 3 and log_entries(s1) == [(0, 0), (1, 1), (2, 1), (3, 1)]
    assert notes_of(notes, NOTE_TRUNCATE)


def test_follower_aer_7(be):
    """follower_aer_7/1, :655-695: overwrite in a new term."""
    nd = Node(be, 3)
    init = empty_state(3, N2)
    _, s00, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, 3, [1, 1, 1, 1]), init)
    _, s0, _, _ = nd.handle_follower(ev_written(0, 1, 4, 4), s00)
    assert s0.last_applied == 3
    _, s1, _, _ = nd.handle_follower(ev_aer(0, N5, 2, 3, 1, 4, [2]), s0)
    role, s2, msgs, _ = nd.handle_follower(ev_written(0, 2, 4, 4), s1)
    assert role == FOLLOWER and s2.last_applied == 4
    f = reply_fields(msgs[0])
    assert (f["to"], f["next_index"], f["last_term"], f["last_index"]) == (N5, 5, 2, 4)


def test_follower_aer_diverged(be):
    """follower_aer_diverged/1, :697-744."""
    nd = Node(be, 3)
    s0 = base_state(3)
    s0.last_applied = 2
    s0.commit_index = 2
    role, s1, msgs, _ = nd.handle_follower(ev_aer(0, N1, 6, 1, 1, 3, [3]), s0)
    assert role == FOLLOWER
    f = reply_fields(msgs[0])
    assert f["success"] is True and f["next_index"] == 3
    assert f["term"] == 5                       # reply carries CurTerm, the pre-update term (:1340)
    assert (s1.last_applied, s1.commit_index, s1.current_term) == (2, 2, 6)
    # leader will send empty rpc
    role, s2, msgs, _ = nd.handle_follower(ev_aer(0, N1, 6, 3, 6, 3, []), s1)
    assert role == AWAIT_CONDITION
    f = reply_fields(msgs[0])
    assert f["success"] is False and f["next_index"] == 3
    aer3 = ev_aer(0, N1, 6, 2, 3, 3, [6])
    role, s2b, msgs, _ = nd.handle_await_condition(aer3, s2)
    assert role == FOLLOWER and len(next_events(msgs)) == 1
    role, s3, msgs, notes = nd.handle_follower(aer3, s2b)
    assert (role, s3.last_applied, s3.commit_index) == (FOLLOWER, 3, 3)
    assert sent(msgs) == [] and status(notes) & ST_LEADER_MSG     # [{aux,eval},{record_leader_msg,_}]
    assert notes_of(notes, NOTE_APPLY)


def test_follower_aer_term_mismatch(be):
    """follower_aer_term_mismatch/1, :746-766."""
    nd = Node(be, 3)
    st = base_state(3)
    st.last_applied = 2
    st.commit_index = 3
    role, s, msgs, _ = nd.handle_follower(ev_aer(0, N1, 6, 3, 6, 3, []), st)
    assert role == AWAIT_CONDITION and s.condition == 2
    assert reply_fields(msgs[0]) == dict(to=N1, from_=N1, term=6, success=False, next_index=3,
                                         last_index=2, last_term=3)


def test_follower_aer_term_mismatch_at_snapshot(be):
    """follower_aer_term_mismatch_at_snapshot/1, :768-820."""
    nd = Node(be, 3)
    st = base_state(3)
    install_snapshot(st, 3, 5)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N1, 5, 3, 5, 3, [5, 5, 5]), st)
    assert (role, s1.last_applied, s1.commit_index) == (FOLLOWER, 3, 3)
    role, s2, msgs, _ = nd.handle_follower(ev_written(0, 5, 4, 6), s1)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["term"], f["success"], f["next_index"]) == (5, True, 7)
    # a new leader deposes the old one: truncate down to the snapshot index
    role, s3, msgs, _ = nd.handle_follower(ev_aer(0, N2, 6, 3, 5, 3, []), s2)
    assert role == FOLLOWER
    assert reply_fields(msgs[0]) == dict(to=N2, from_=N1, term=6, success=True, next_index=4,
                                         last_index=3, last_term=5)
    assert (s3.last_applied, s3.commit_index) == (3, 3)


def test_follower_aer_term_mismatch_snapshot(be):
    """follower_aer_term_mismatch_snapshot/1, :822-850."""
    nd = Node(be, 3)
    st = base_state(3)
    install_snapshot(st, 3, 5)
    role, s, msgs, _ = nd.handle_follower(ev_aer(0, N1, 6, 3, 6, 3, []), st)
    assert role == AWAIT_CONDITION
    assert reply_fields(msgs[0]) == dict(to=N1, from_=N1, term=6, success=False, next_index=4,
                                         last_index=3, last_term=5)


def test_follower_handles_append_entries_rpc(be):
    """follower_handles_append_entries_rpc/1, :852-930."""
    nd = Node(be, 3)
    st = base_state(3)
    st.commit_index = 1
    empty_ae = lambda **kw: ev_aer(0, N1, kw.get("term", 5), kw.get("prev", 3), kw.get("prev_term", 5), 3, [])
    role, s, _, _ = nd.handle_follower(empty_ae(), st)
    assert (role, s.leader_slot, s.current_term) == (FOLLOWER, N1, 5)
    # leader term is higher: reply term updated, empty rpc replied immediately
    role, s, msgs, _ = nd.handle_follower(empty_ae(term=6), st)
    assert (role, s.leader_slot, s.current_term) == (FOLLOWER, N1, 6)
    assert reply_fields(msgs[0]) == dict(to=N1, from_=N1, term=6, success=True, next_index=4,
                                         last_index=3, last_term=5)
    # reply false if term < current_term (5.1)
    role, s, msgs, _ = nd.handle_follower(empty_ae(term=4), st)
    (r,) = msgs
    f = reply_fields(r)
    assert (role, f["term"], f["success"]) == (FOLLOWER, 5, False)
    # no entry at prev_log_index
    role, s, msgs, notes = nd.handle_follower(empty_ae(prev=4), st)
    f = reply_fields(msgs[0])
    assert (role, f["term"], f["success"]) == (AWAIT_CONDITION, 5, False)
    assert status(notes) & ST_LEADER_MSG
    # entry with a different term
    role, s, msgs, notes = nd.handle_follower(empty_ae(prev_term=4), st)
    f = reply_fields(msgs[0])
    assert (role, f["term"], f["success"]) == (AWAIT_CONDITION, 5, False)
    assert status(notes) & ST_LEADER_MSG
    # truncate/overwrite a conflicting entry (5.3)
    st2 = clone(st)
    st2.last_applied = 1
    role, inter3, _, _ = nd.handle_follower(ev_aer(0, N1, 5, 1, 1, 2, [4]), st2)
    role, s, msgs, _ = nd.handle_follower(ev_written(0, 4, 2, 2), inter3)
    (r,) = msgs
    assert reply_fields(r) == dict(to=N1, from_=N1, term=5, success=True, next_index=3, last_index=2, last_term=4)
    assert log_entries(s) == [(0, 0), (1, 1), (2, 4)]
    # leader_commit beyond the last entry received
    st3 = clone(st)
    st3.commit_index = 1
    st3.last_applied = 1
    e = ev_aer(0, N1, 5, 3, 5, 5, [5])
    role, inter4, _, _ = nd.handle_follower(e, st3)
    role, s, msgs, _ = nd.handle_follower(ev_written(0, 5, 4, 4), inter4)
    (r,) = msgs
    f = reply_fields(r)
    assert (s.commit_index, s.last_applied) == (5, 4)
    assert (f["term"], f["success"], f["last_index"], f["last_term"]) == (5, True, 4, 5)


def test_follower_catchup_condition(be):
    """follower_catchup_condition/1, :932-999."""
    nd = Node(be, 3)
    st0 = base_state(3)
    st0.commit_index = 1
    empty_ae = lambda **kw: ev_aer(0, N1, kw.get("term", 5), kw.get("prev", 3), kw.get("prev_term", 5), 3, [])
    role, st, _, _ = nd.handle_follower(empty_ae(prev=4), st0)
    assert role == AWAIT_CONDITION and st.condition == 1
    # lower leader term does not enter await_condition
    role, _, msgs, _ = nd.handle_follower(empty_ae(term=4, prev=4), st)
    assert role == FOLLOWER and len(msgs) == 1
    # prev index exists with a different term
    role, _, msgs, notes = nd.handle_follower(empty_ae(term=6, prev_term=4), st)
    assert role == AWAIT_CONDITION and len(msgs) == 1 and status(notes) & ST_LEADER_MSG
    # still a gap: stay, no reply
    role, _, msgs, _ = nd.handle_await_condition(empty_ae(prev=4), st)
    assert role == AWAIT_CONDITION and msgs == []
    # success: back to follower and the AER is re-queued
    role, _, msgs, _ = nd.handle_await_condition(empty_ae(), st)
    assert role == FOLLOWER
    (ne,) = next_events(msgs)
    assert ne.type == EV_AER and (ne.term, ne.a, ne.b, ne.c, ne.from_slot) == (5, 3, 5, 3, N1)
    # log events are just forwarded
    role, _, msgs, _ = nd.handle_await_condition(ev_written(0, 99, 99, 99), st)
    assert role == AWAIT_CONDITION and msgs == []
    # request_vote_rpc: back to follower, state unchanged, re-queued
    rv = ev_request_vote(0, N2, 6, 3, 5)
    role, s, msgs, _ = nd.handle_await_condition(rv, st)
    assert role == FOLLOWER and len(next_events(msgs)) == 1
    want = clone(st)
    want.role = FOLLOWER
    assert s.key() == want.key()
    # timeout: the reply effect is repeated
    role, _, msgs, notes = nd.handle_await_condition(ev_simple(0, EV_AWAIT_COND_TIMEOUT), st)
    assert role == FOLLOWER
    (r,) = msgs
    f = reply_fields(r)
    assert (f["to"], f["success"], f["next_index"]) == (N1, False, 4)
    assert status(notes) & ST_LEADER_MSG
    role, _, _, _ = nd.handle_await_condition(ev_simple(0, EV_ELECTION_TIMEOUT), st)
    assert role == PRE_VOTE


def test_candidate_handles_append_entries_rpc(be):
    """candidate_handles_append_entries_rpc/1, :1146-1159."""
    nd = Node(be, 3)
    st = base_state(3)
    st.commit_index = 1
    role, _, msgs, _ = nd.handle_candidate(ev_aer(0, N1, 4, 3, 5, 3, []), st)
    (r,) = msgs
    f = reply_fields(r)
    assert (role, f["term"], f["success"], f["last_index"], f["last_term"]) == (CANDIDATE, 5, False, 3, 5)


def test_follower_aer_dupe(be):
    """follower_aer_dupe/1, :1250-1282."""
    nd = Node(be, 3)
    init = empty_state(3, N1)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N2, 1, 0, 0, 1, [1, 1, 1]), init)
    assert (s1.leader_slot, s1.current_term, s1.commit_index, s1.last_applied) == (N2, 1, 1, 1)
    role, s2, msgs, _ = nd.handle_follower(ev_aer(0, N2, 1, 1, 1, 1, [1]), s1)
    (r,) = msgs
    f = reply_fields(r)
    assert (s2.commit_index, s2.last_applied) == (1, 1)
    assert (f["to"], f["from_"], f["success"], f["next_index"], f["last_term"], f["last_index"]) == \
        (N2, N1, True, 3, 1, 2)


def test_follower_leader_change_before_written(be):
    """follower_leader_change_before_written/1, :1284-1326."""
    nd = Node(be, 3)
    init = empty_state(3, N3)
    role, s1, _, _ = nd.handle_follower(ev_aer(0, N1, 1, 0, 0, 1, [1, 1]), init)
    assert (s1.leader_slot, s1.current_term, s1.commit_index, s1.last_applied) == (N1, 1, 1, 1)
    # NB the reference test sends prev_log_index 0 with entries 2,3; with contiguous AER
    # records the same overwrite is prev_log_index 1 (index 1 is untouched either way)
    role, s2, _, _ = nd.handle_follower(ev_aer(0, N2, 2, 1, 1, 1, [2, 2]), s1)
    assert (s2.leader_slot, s2.current_term, s2.commit_index, s2.last_applied) == (N2, 2, 1, 1)
    role, s3, msgs, _ = nd.handle_follower(ev_written(0, 1, 1, 2), s2)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["to"], f["from_"], f["success"], f["term"], f["last_index"], f["last_term"]) == (N2, N3, True, 2, 1, 1)
    role, s4, msgs, _ = nd.handle_follower(ev_written(0, 2, 2, 3), s3)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["success"], f["term"], f["last_index"], f["last_term"]) == (True, 2, 3, 2)
    assert s4.last_applied == 1


def _set_peers(st, peers):
    for slot, (ni, mi, cis) in peers.items():
        st.peers[slot].next_index = ni
        st.peers[slot].match_index = mi
        st.peers[slot].commit_index_sent = cis


def test_append_entries_reply_success(be):
    """append_entries_reply_success/1, :1328-1375."""
    nd = Node(be, 3)
    st0 = base_state(3)
    st0.commit_index = 1
    st0.last_applied = 1
    _set_peers(st0, {N1: (5, 4, 0), N2: (1, 0, 3), N3: (2, 1, 0)})
    msg = ev_aer_reply(0, N2, 5, True, 4, 3, 5)
    role, s, msgs, notes = nd.handle_leader(msg, st0)
    assert role == LEADER and peer(s, N2)[:2] == (4, 3)
    assert (s.commit_index, s.last_applied) == (3, 3)
    (ne,) = msgs
    assert ne.flags & EVF_NEXT_EVENT and ne.type == EV_PIPELINE_RPCS        # {next_event, info, pipeline_rpcs}
    (cm,) = notes_of(notes, NOTE_COMMIT)                                      # {aux, eval}
    assert (cm.a, cm.b) == (1, 3)
    role, s2, msgs, _ = nd.handle_leader(ev_simple(0, EV_PIPELINE_RPCS), s)
    assert role == LEADER and peer(s2, N2)[:2] == (4, 3) and (s2.commit_index, s2.last_applied) == (3, 3)
    (rpc,) = msgs
    assert aer_fields(rpc) == dict(to=N3, leader=N1, term=5, prev_log_index=1, prev_log_term=1,
                                   leader_commit=3, entries=[(2, 3), (3, 5)])
    # §5.4.2: the entry's term is not the current term -> no commit
    st7 = clone(st0)
    st7.current_term = 7
    role, s3, _, _ = nd.handle_leader(ev_aer_reply(0, N2, 7, True, 4, 3, 5), st7)
    assert role == LEADER and peer(s3, N2)[:2] == (4, 3)
    assert (s3.commit_index, s3.last_applied, s3.current_term) == (1, 1, 7)


def test_append_entries_reply_no_success(be):
    """append_entries_reply_no_success/1, :1377-1404."""
    nd = Node(be, 3)
    st = base_state(3)
    st.commit_index = 1
    st.last_applied = 1
    _set_peers(st, {N1: (1, 0, 0), N2: (3, 0, 0), N3: (2, 1, 1)})
    role, s, msgs, _ = nd.handle_leader(ev_aer_reply(0, N2, 5, False, 2, 1, 1), st)
    assert role == LEADER and peer(s, N2)[:2] == (4, 1)
    assert (s.commit_index, s.last_applied) == (1, 1)
    rpcs = {m.row: aer_fields(m) for m in msgs}
    assert set(rpcs) == {N2, N3}
    assert rpcs[N3] == dict(to=N3, leader=N1, term=5, prev_log_index=1, prev_log_term=1, leader_commit=1,
                            entries=[(2, 3), (3, 5)])


def test_append_entries_reply_no_success_from_unknown_peer(be):
    """append_entries_reply_no_success_from_unknown_peer/1, :1406-1420."""
    nd = Node(be, 3)
    st = base_state(3)
    st.commit_index = 1
    st.last_applied = 1
    role, s, msgs, _ = nd.handle_leader(ev_aer_reply(0, 7, 5, False, 2, 1, 1), st)
    want = clone(st)
    want.role = LEADER
    assert role == LEADER and msgs == [] and s.key() == want.key()


def test_follower_request_vote(be):
    """follower_request_vote/1, :1422-1470."""
    nd = Node(be, 3)
    st = base_state(3)

    def rv(cand=N2, term=6, lli=3, llt=5):
        return ev_request_vote(0, cand, term, lli, llt)

    def res(msgs):
        (m,) = msgs
        assert m.type == EV_REQUEST_VOTE_RES
        return m.row, m.term, bool(m.d)

    role, s1, msgs, notes = nd.handle_follower(rv(), st)
    assert (role, s1.voted_for, s1.current_term) == (FOLLOWER, N2, 6) and res(msgs) == (N2, 6, True)
    assert status(notes) & ST_TERM_VOTE_CHANGED
    role, s2, msgs, _ = nd.handle_follower(rv(), s1)                  # same candidate again
    assert (s2.voted_for, s2.current_term) == (N2, 6) and res(msgs) == (N2, 6, True)
    role, s3, msgs, _ = nd.handle_follower(rv(cand=N3), s1)           # different candidate
    assert (s3.voted_for, s3.current_term) == (N2, 6) and res(msgs) == (N3, 6, False)
    role, s4, msgs, _ = nd.handle_follower(rv(term=4), st)            # lower term
    assert s4.current_term == 5 and res(msgs) == (N2, 5, False)
    role, s5, msgs, _ = nd.handle_follower(rv(llt=4), st)             # candidate log older: term still bumps
    assert s5.current_term == 6 and res(msgs) == (N2, 6, False)
    role, s6, msgs, _ = nd.handle_follower(rv(lli=4), st)             # same term, longer log
    assert s6.current_term == 6 and res(msgs) == (N2, 6, True)
    nv = clone(st)
    nv.membership = PROMOTABLE
    role, s7, msgs, _ = nd.handle_follower(rv(), nv)                  # non-voters ignore
    assert role == FOLLOWER and msgs == [] and s7.key() == nv.key()


def test_follower_pre_vote(be):
    """follower_pre_vote/1, :1472-1590."""
    nd = Node(be, 3)
    st = base_state(3)
    TOKEN = 0xABCDEF

    def pv(**kw):
        return ev_pre_vote(0, N2, kw.get("term", 5), TOKEN, kw.get("lli", 3), kw.get("llt", 5),
                           version=kw.get("version", 1), machine_version=kw.get("mv", 0))

    def res(msgs):
        m = msgs[0]
        assert m.type == EV_PRE_VOTE_RES and m.c == TOKEN and m.row == N2
        return m.term, bool(m.d)

    def with_cfg(eff, mac):
        s = clone(st)
        s.effective_machine_version = eff
        s.machine_version = mac
        return s

    role, s, msgs, _ = nd.handle_follower(pv(), st)
    assert (role, s.current_term) == (FOLLOWER, 5) and res(msgs) == (5, True) and len(msgs) == 1
    _, _, msgs, _ = nd.handle_follower(pv(version=2), st)               # higher protocol version
    assert res(msgs) == (5, False)
    _, _, msgs, _ = nd.handle_follower(pv(version=0), st)               # lower protocol version
    assert res(msgs) == (5, True)
    _, _, msgs, notes = nd.handle_follower(pv(mv=99), st)
    assert res(msgs) == (5, False) and status(notes) & ST_START_ELECTION_TMO
    _, _, msgs, _ = nd.handle_follower(pv(mv=1), with_cfg(1, 0))
    assert res(msgs) == (5, True)
    _, _, msgs, _ = nd.handle_follower(pv(mv=0), with_cfg(1, 1))
    assert res(msgs) == (5, False)
    _, _, msgs, _ = nd.handle_follower(pv(mv=2), with_cfg(3, 2))
    assert res(msgs) == (5, False)
    _, _, msgs, _ = nd.handle_follower(pv(mv=2), with_cfg(1, 3))
    assert res(msgs) == (5, True)
    _, _, msgs, _ = nd.handle_follower(pv(mv=0), with_cfg(0, 1))
    assert res(msgs) == (5, True)
    # same machine version (the reference changes only machine_version; effective stays 0,
    # 2 is within [effective, local])
    _, _, msgs, _ = nd.handle_follower(pv(mv=2), with_cfg(0, 2))
    assert res(msgs) == (5, True) and len(msgs) == 1
    role, s, msgs, _ = nd.handle_follower(pv(term=4), st)               # lower term
    assert s.current_term == 5 and res(msgs) == (5, False) and len(msgs) == 1
    role, s, msgs, notes = nd.handle_follower(pv(llt=4, term=6), st)    # better candidate here
    assert s.current_term == 6 and msgs == [] and status(notes) & ST_START_ELECTION_TMO
    role, s, msgs, _ = nd.handle_follower(pv(lli=4), st)
    assert s.current_term == 5 and res(msgs) == (5, True)
    nv = clone(st)
    nv.membership = PROMOTABLE
    role, s, msgs, _ = nd.handle_follower(pv(), nv)
    assert role == FOLLOWER and msgs == [] and s.key() == nv.key()


def test_pre_vote_does_not_set_voted_for(be):
    """pre_vote_does_not_set_voted_for/1, :1592-1622."""
    nd = Node(be, 3)
    st0 = base_state(3)
    role, s1, msgs, _ = nd.handle_follower(ev_pre_vote(0, N2, 5, 77, 3, 5), st0)
    assert msgs[0].type == EV_PRE_VOTE_RES and msgs[0].d == 1
    assert s1.voted_for == RA_NO_SLOT
    role, s2, msgs, _ = nd.handle_follower(ev_request_vote(0, N3, 5, 3, 5), s1)
    assert s2.voted_for == N3 and msgs[0].type == EV_REQUEST_VOTE_RES and msgs[0].d == 1 and msgs[0].term == 5


def test_pre_vote_and_await_condition_receive_pre_vote(be):
    """pre_vote_receives_pre_vote/1 :1624-1638, await_condition_receives_pre_vote/1 :1640-1656."""
    nd = Node(be, 3)
    st = base_state(3)
    pv = ev_pre_vote(0, N2, 5, 99, 3, 5)
    role, s, msgs, _ = nd.handle_pre_vote(pv, st)
    assert role == PRE_VOTE and s.current_term == 5
    (m,) = msgs
    assert (m.type, m.term, m.c, m.d) == (EV_PRE_VOTE_RES, 5, 99, 1)
    role, s, msgs, _ = nd.handle_await_condition(pv, st)
    assert role == AWAIT_CONDITION and s.current_term == 5
    (m,) = msgs
    assert (m.type, m.term, m.c, m.d) == (EV_PRE_VOTE_RES, 5, 99, 1)


def test_request_vote_rpc_with_lower_term(be):
    """request_vote_rpc_with_lower_term/1, :1658-1671."""
    nd = Node(be, 3)
    st = base_state(3)
    st.current_term = 6
    st.voted_for = N1
    rv = ev_request_vote(0, N2, 5, 3, 5)
    role, s, msgs, _ = nd.handle_candidate(rv, st)
    (m,) = msgs
    assert (role, s.voted_for, s.current_term) == (CANDIDATE, N1, 6)
    assert (m.type, m.row, m.term, m.d) == (EV_REQUEST_VOTE_RES, N2, 6, 0)
    role, s, msgs, _ = nd.handle_leader(rv, st)
    (m,) = msgs
    assert (role, s.current_term) == (LEADER, 6) and (m.type, m.term, m.d) == (EV_REQUEST_VOTE_RES, 6, 0)


def test_leader_does_not_abdicate_to_unknown_peer(be):
    """leader_does_not_abdicate_to_unknown_peer/1, :1673-1691 (install_snapshot_result omitted)."""
    nd = Node(be, 3)
    st = base_state(3)
    want = clone(st)
    want.role = LEADER
    role, s, msgs, _ = nd.handle_leader(ev_request_vote(0, 7, 6, 3, 5), st)
    assert role == LEADER and msgs == [] and s.key() == want.key()
    role, s, msgs, _ = nd.handle_leader(ev_aer_reply(0, 7, 6, False, 4, 3, 5), st)
    assert role == LEADER and msgs == [] and s.key() == want.key()


def test_leader_replies_to_append_entries_rpc_with_lower_term(be):
    """leader_replies_to_append_entries_rpc_with_lower_term/1, :1694-1708."""
    nd = Node(be, 3)
    st = base_state(3)
    role, _, msgs, _ = nd.handle_leader(ev_aer(0, N3, 4, 3, 5, 3, []), st)
    (r,) = msgs
    f = reply_fields(r)
    assert (role, f["to"], f["from_"], f["term"], f["success"]) == (LEADER, N3, N1, 5, False)


def test_higher_term_detected(be):
    """higher_term_detected/1, :1710-1752."""
    nd = Node(be, 3)
    st = base_state(3)
    rep = ev_aer_reply(0, N2, 6, False, 4, 3, 5)
    role, s, msgs, _ = nd.handle_leader(rep, st)
    assert (role, s.current_term, s.leader_slot) == (FOLLOWER, 6, RA_NO_SLOT) and msgs == []
    role, s, msgs, _ = nd.handle_follower(rep, st)
    assert (role, s.current_term) == (FOLLOWER, 6) and msgs == []
    role, s, msgs, _ = nd.handle_candidate(rep, st)
    assert (role, s.current_term) == (FOLLOWER, 6) and msgs == []
    aer = ev_aer(0, N3, 6, 3, 5, 3, [])
    role, s, msgs, _ = nd.handle_leader(aer, st)
    (ne,) = msgs
    assert (role, s.current_term, s.leader_slot) == (FOLLOWER, 6, RA_NO_SLOT)
    assert ne.flags & EVF_NEXT_EVENT and (ne.type, ne.term, ne.from_slot, ne.a, ne.b, ne.c) == (EV_AER, 6, N3, 3, 5, 3)
    role, s, msgs, _ = nd.handle_candidate(aer, st)
    assert (role, s.current_term) == (FOLLOWER, 6) and len(next_events(msgs)) == 1
    vote = ev_request_vote(0, N2, 6, 3, 5)
    role, s, msgs, _ = nd.handle_leader(vote, st)
    (ne,) = msgs
    assert (role, s.current_term, s.leader_slot) == (FOLLOWER, 6, RA_NO_SLOT)
    assert ne.flags & EVF_NEXT_EVENT and (ne.type, ne.term, ne.from_slot) == (EV_REQUEST_VOTE, 6, N2)
    assert log_entries(s) == log_entries(st)
    role, s, msgs, _ = nd.handle_candidate(vote, st)
    assert (role, s.current_term) == (FOLLOWER, 6) and len(next_events(msgs)) == 1


def test_leader_noop_then_commit(be):
    """leader_noop_operation_enables_cluster_change/1, :1754-1766: the index/term/commit part."""
    nd = Node(be, 3)
    st00 = base_state(3)
    role, s0, msgs, notes = nd.handle_leader(ev_command(0, 1, noop=True), st00)
    assert role == LEADER and (s0.last_index, s0.last_term) == (4, 5)
    (w,) = notes_of(notes, NOTE_WAL_APPEND)
    assert (w.a, w.b, w.c) == (4, 4, 5)
    role, s, _, _ = nd.handle_leader(ev_written(0, 5, 4, 4), s0)
    role, s2, _, notes = nd.handle_leader(ev_aer_reply(0, N2, 5, True, 5, 4, 5), s)
    assert (s2.commit_index, s2.last_applied) == (4, 4)
    assert [(n.a, n.b) for n in notes_of(notes, NOTE_COMMIT)] == [(3, 4)]


def test_command(be):
    """command/1 :2287-2304 and command_notify/1 :2352-2384."""
    nd = Node(be, 3)
    st = base_state(3)
    role, s1, msgs, notes = nd.handle_leader(ev_command(0, 1), st)
    assert role == LEADER
    want = dict(leader=N1, term=5, prev_log_index=3, prev_log_term=5, leader_commit=3, entries=[(4, 5)])
    got = {m.row: aer_fields(m) for m in msgs}
    assert got == {N2: dict(to=N2, **want), N3: dict(to=N3, **want)}
    (w,) = notes_of(notes, NOTE_WAL_APPEND)                # {Idx, Term} = {4, 5}
    assert (w.a, w.b, w.c) == (4, 4, 5)
    role, s, _, _ = nd.handle_leader(ev_written(0, 5, 4, 4), s1)
    role, s2, msgs, notes = nd.handle_leader(ev_aer_reply(0, N2, 5, True, 5, 4, 5), s)
    assert next_events(msgs) and notes_of(notes, NOTE_COMMIT) and notes_of(notes, NOTE_APPLY)
    assert (s2.commit_index, s2.last_applied) == (4, 4)


def test_candidate_election(be):
    """candidate_election/1, :2386-2431."""
    nd = Node(be, 5)
    st = base_state(5)
    st.current_term = 6
    st.votes = 1
    reply = ev_request_vote_result(0, 6, True)
    role, s1, msgs, _ = nd.handle_candidate(reply, st)
    assert (role, s1.votes, msgs) == (CANDIDATE, 2, [])
    role, s, msgs, _ = nd.handle_candidate(ev_request_vote_result(0, 6, False), s1)
    assert (role, s.votes, msgs) == (CANDIDATE, 2, [])
    role, s, msgs, _ = nd.handle_candidate(ev_request_vote_result(0, 7, False), s1)
    assert (role, s.current_term, msgs) == (FOLLOWER, 7, [])
    # quorum: candidate becomes leader, peers re-initialised, noop queued
    role, s2, msgs, _ = nd.handle_candidate(reply, s1)
    assert role == LEADER and s2.leader_slot == N1
    for p in (N2, N3, N4, N5):
        assert peer(s2, p) == (4, 0, 0)
    (ne,) = msgs
    assert ne.flags & EVF_NEXT_EVENT and ne.type == EV_COMMAND and ne.flags & EVF_NOOP
    noop = ev_command(0, 1, noop=True)
    role, s3, msgs, _ = nd.handle_leader(noop, s2)
    assert role == LEADER and sorted(m.row for m in msgs) == [N2, N3, N4, N5]
    assert all(m.type == EV_AER for m in msgs)


def test_pre_vote_election(be):
    """pre_vote_election/1 :2433-2460, pre_vote_election_non_voter/1 :2462-2470."""
    nd = Node(be, 5)
    TOKEN = 4242
    st = base_state(5)
    st.votes = 1
    st.pre_vote_token = TOKEN
    reply = ev_pre_vote_result(0, 5, TOKEN, True)
    role, s1, msgs, _ = nd.handle_pre_vote(reply, st)
    assert (role, s1.votes, msgs) == (PRE_VOTE, 2, [])
    role, s, msgs, _ = nd.handle_pre_vote(ev_pre_vote_result(0, 5, TOKEN + 1, True), st)
    assert (role, s.votes, msgs) == (PRE_VOTE, 1, [])                      # different token ignored
    role, s, msgs, _ = nd.handle_pre_vote(ev_pre_vote_result(0, 5, TOKEN, False), s1)
    assert (role, s.votes, msgs) == (PRE_VOTE, 2, [])
    role, s, msgs, _ = nd.handle_pre_vote(ev_pre_vote_result(0, 6, TOKEN, False), s1)
    assert (role, s.current_term, s.votes, msgs) == (FOLLOWER, 6, 0, [])
    role, s, msgs, _ = nd.handle_pre_vote(reply, s1)
    assert (role, s.current_term) == (CANDIDATE, 6)
    nv = clone(st)
    nv.membership = PROMOTABLE
    role, s, msgs, _ = nd.handle_pre_vote(reply, nv)
    assert (role, s.votes, msgs) == (PRE_VOTE, 1, [])


def test_pre_vote_election_reverts(be):
    """pre_vote_election_reverts/1, :2472-2501 (install_snapshot_rpc omitted)."""
    nd = Node(be, 5)
    st = base_state(5)
    st.votes = 1
    st.pre_vote_token = 1
    vote = ev_request_vote(0, N2, 6, 3, 5)
    role, s, msgs, _ = nd.handle_pre_vote(vote, st)
    assert (role, s.current_term, s.votes) == (FOLLOWER, 6, 0) and len(next_events(msgs)) == 1
    ae = ev_aer(0, N2, 5, 3, 5, 3, [])
    role, s, msgs, _ = nd.handle_pre_vote(ae, st)
    assert (role, s.current_term, s.votes) == (FOLLOWER, 5, 0) and len(next_events(msgs)) == 1
    ae6 = ev_aer(0, N2, 6, 3, 5, 3, [])
    role, s, msgs, _ = nd.handle_pre_vote(ae6, st)
    assert (role, s.current_term, s.votes) == (FOLLOWER, 6, 0) and len(next_events(msgs)) == 1


def test_snapshotted_follower_received_append_entries(be):
    """snapshotted_follower_received_append_entries/1, :3030-3077: prev = snapshot index."""
    nd = Node(be, 3)
    st = empty_state(3, N3)
    st.current_term = 2
    st.leader_slot = N1
    st.commit_index = 3
    st.last_applied = 3
    install_snapshot(st, 3, 2)
    role, inter, _, _ = nd.handle_follower(ev_aer(0, N1, 2, 3, 2, 4, [2]), st)
    assert role == FOLLOWER and (inter.last_index, inter.last_term, inter.last_applied) == (4, 2, 4)
    role, s, msgs, _ = nd.handle_follower(ev_written(0, 2, 4, 4), inter)
    (r,) = msgs
    f = reply_fields(r)
    assert (f["to"], f["from_"], f["success"]) == (N1, N3, True)


def test_leader_received_append_entries_reply_with_stale_last_index(be):
    """leader_received_append_entries_reply_with_stale_last_index/1, :3079-3132."""
    nd = Node(be, 3)
    st = empty_state(3, N1)
    st.role = LEADER
    st.leader_slot = N1
    st.current_term = 2
    st.commit_index = 3
    st.last_applied = 4
    set_log(st, [(0, 0), (1, 1), (2, 2), (3, 5)], last_written=(3, 5))
    _set_peers(st, {N1: (1, 0, 0), N2: (3, 0, 0), N3: (4, 3, 3)})
    role, s, msgs, _ = nd.handle_leader(ev_aer_reply(0, N2, 2, False, 3, 2, 1), st)
    assert role == LEADER and peer(s, N2)[0] == 4
    (rpc,) = msgs
    f = aer_fields(rpc)
    assert f["to"] == N2 and [i for i, _ in f["entries"]] == [2, 3]


def test_leader_saw_append_entries_rpc_in_same_term(be):
    """exit(leader_saw_append_entries_rpc_in_same_term), src/ra_server.erl:836-840."""
    nd = Node(be, 3)
    st = base_state(3)
    role, s, msgs, notes = nd.handle_leader(ev_aer(0, N2, 5, 3, 5, 3, []), st)
    assert status(notes) & ST_FATAL
    (n,) = notes_of(notes, NOTE_STATUS)
    assert n.c == FATAL_LEADER_SAW_AER_SAME_TERM


def test_candidate_receives_pre_vote(be):
    """candidate_receives_pre_vote/1, :2503-2525."""
    nd = Node(be, 5)
    st = base_state(5)
    st.votes = 1
    TOKEN = 777
    pv = ev_pre_vote(0, N1, 5, TOKEN, 3, 5)
    role, s, msgs, _ = nd.handle_candidate(pv, st)          # not a lower index: granted
    (m,) = msgs
    assert role == CANDIDATE and (m.type, m.c, m.d) == (EV_PRE_VOTE_RES, TOKEN, 1)
    role, s, msgs, _ = nd.handle_candidate(ev_pre_vote(0, N1, 5, TOKEN, 2, 5), st)   # lower index: refused
    (m,) = msgs
    assert role == CANDIDATE and (m.type, m.c, m.d) == (EV_PRE_VOTE_RES, TOKEN, 0)
    role, s, msgs, _ = nd.handle_candidate(ev_pre_vote(0, N1, 6, TOKEN, 3, 5), st)   # higher term: abdicates
    assert (role, s.current_term) == (FOLLOWER, 6)


def test_leader_receives_pre_vote(be):
    """leader_receives_pre_vote/1, :2527-2546: an rpc to every peer at once, abdication on a higher term."""
    nd = Node(be, 5)
    st = base_state(5)
    st.votes = 1
    pv = ev_pre_vote(0, N1, 5, 31337, 3, 5)
    role, s, msgs, _ = nd.handle_leader(pv, st)
    rpcs = of_type(msgs, EV_AER)
    assert role == LEADER and sorted(m.row for m in rpcs) == [N2, N3, N4, N5]
    assert all((m.term, m.a, m.b, m.c, m.n) == (5, 3, 5, 3, 0) for m in rpcs)
    role, s, msgs, _ = nd.handle_leader(ev_pre_vote(0, N1, 6, 31337, 3, 5), st)
    assert (role, s.current_term) == (FOLLOWER, 6)


def test_persist_last_applied_with_unwritten(be):
    """persist_last_applied_with_unwritten/1, :3771-3793, the part that is ra_server's: a follower
    applies a committed entry before its own written event (last_written stays {0,0});
    persisted_last_applied itself (ra_log_meta) stays on the host."""
    nd = Node(be, 3)
    st = empty_state(3, N1)
    aer = ev_aer(0, N1, 1, 0, 0, 1, [1])
    role, s, msgs, notes = nd.handle_follower(aer, st)
    assert (role, s.leader_slot, s.current_term, s.commit_index, s.last_applied) == (FOLLOWER, N1, 1, 1, 1)
    assert (s.last_written_index, s.last_written_term) == (0, 0)
    (ap,) = notes_of(notes, NOTE_APPLY)
    assert (ap.a, ap.b) == (1, 1)
    role, s2, msgs, notes = nd.handle_follower(ev_written(0, 1, 1, 1), s)
    assert (s2.last_written_index, s2.last_written_term, s2.last_applied) == (1, 1, 1)


def test_follower_state_resets_peer_status(be):
    """follower_state_resets_peer_status/1, :2321-2350: handle_state_enter(follower, leader, _) puts
    every peer status back to normal.  The engine runs become/3 as part of the role change (not in
    `pure` mode, where one call is exactly one handle_<state>/2 clause)."""
    nd = Node(be, 3, pure=False)
    st = base_state(3)
    st.peers[N2].status = PEER_SENDING_SNAPSHOT
    st.peers[N3].status = PEER_DISCONNECTED
    role, s, msgs, notes = nd.handle_leader(ev_request_vote(0, N2, 6, 3, 5), st)    # higher term: steps down
    assert (role, s.current_term) == (FOLLOWER, 6)
    assert [s.peers[p].status for p in (N2, N3)] == [PEER_NORMAL, PEER_NORMAL]
    assert status(notes) & ST_ROLE_CHANGED


def test_load_rows_rejects_inconsistent_log_views(be):
    """ra_row_state_valid (include/ra_engine.h): what load_rows requires of a row's log view."""
    b = make_backend(be, 1, 3)
    ok = base_state(3)
    b.load_rows([ok])

    def bad(mut):
        s = clone(ok)
        mut(s)
        with pytest.raises(abi.RaError) as ei:
            b.load_rows([s])
        assert ei.value.status == abi.RA_E_INVAL

    bad(lambda s: setattr(s, "n_runs", 0))                          # entries but no runs
    bad(lambda s: setattr(s, "last_term", 4))                       # last_term is not the last run's term
    bad(lambda s: s.run_start.__setitem__(0, 1))                    # first run does not start at first_index
    bad(lambda s: s.run_start.__setitem__(2, 1))                    # run starts not ascending
    bad(lambda s: s.run_term.__setitem__(2, 0))                     # terms decreasing
    bad(lambda s: setattr(s, "last_index", 1))                      # a run starts beyond last_index
    bad(lambda s: setattr(s, "n_members", 5))                       # not this engine's group size
    bad(lambda s: setattr(s, "row", 99))                            # no such row
    assert b.read_rows([ok.row])[0].key() == ok.key()               # the good row is untouched


def test_reference_counters(be):
    """The reference's own counters of this path (ra.hrl:324-343): where ra_server.erl increments them
    (:528, :590, :1278, :1290, :1418, :2856, :2878, :3026)."""
    nd = Node(be, 3)
    st = base_state(3)
    c0 = nd.b.counters()

    def delta():
        c = nd.b.counters()
        return {k: c[k] - c0[k] for k in ("aer_received_follower", "aer_received_follower_empty", "aer_replies_success",
                                           "aer_replies_failed", "elections", "pre_vote_elections",
                                           "term_and_voted_for_updates") if c[k] != c0[k]}
    nd.handle_follower(ev_aer(0, N2, 5, 3, 5, 3, []), st)                 # up to date, no entries
    assert delta() == dict(aer_received_follower=1, aer_received_follower_empty=1)
    nd.handle_follower(ev_aer(0, N2, 4, 3, 5, 3, []), st)                 # stale term: counted too (:1418)
    assert delta() == dict(aer_received_follower=2, aer_received_follower_empty=1)
    nd.handle_follower(ev_aer(0, N2, 5, 3, 5, 3, [5]), st)                # one new entry: not "empty"
    assert delta()["aer_received_follower"] == 3 and delta()["aer_received_follower_empty"] == 1
    nd.handle_leader(ev_aer_reply(0, N2, 5, True, 4, 3, 5), st)
    nd.handle_leader(ev_aer_reply(0, N2, 5, False, 4, 3, 5), st)
    d = delta()
    assert (d["aer_replies_success"], d["aer_replies_failed"]) == (1, 1)
    nd.handle_follower(ev_simple(0, EV_ELECTION_TIMEOUT), st)             # pre-vote round: votes for itself
    d = delta()
    assert d["pre_vote_elections"] == 1 and d["term_and_voted_for_updates"] == 1 and "elections" not in d
    pv = clone(st)
    pv.votes = 1
    pv.pre_vote_token = 9
    role, s, _, _ = nd.handle_pre_vote(ev_pre_vote_result(0, 5, 9, True), pv)   # quorum of 3: becomes candidate
    assert role == CANDIDATE
    d = delta()
    assert d["elections"] == 1 and d["term_and_voted_for_updates"] == 2


def test_leader_pre_vote_sends_rpc_to_backoff_peer(be):
    """leader_pre_vote_sends_snapshot_to_backoff_peer/1, :2548-2574: make_all_rpcs/1 (:2337-2350) cancels the
    snapshot retry timer of a peer in snapshot_backoff and sends it an rpc as well."""
    nd = Node(be, 3)
    st = base_state(3)
    st.votes = 1
    st.peers[N2].status = PEER_SNAPSHOT_BACKOFF
    role, s, msgs, notes = nd.handle_leader(ev_pre_vote(0, N1, 5, 77, 3, 5), st)
    assert role == LEADER
    assert [(n.slot, n.a) for n in notes_of(notes, NOTE_CANCEL_SNAPSHOT_RETRY)] == [(N2, N2)]
    assert sorted(m.row for m in of_type(msgs, EV_AER)) == [N2, N3]
    # a tick (make_rpcs/1 over stale_peers/1) still leaves the backoff peer alone
    st.peers[N2].match_index = 1
    st.peers[N3].match_index = 1
    role, s, msgs, notes = nd.handle_leader(ev_simple(0, EV_TICK), st)
    assert [m.row for m in of_type(msgs, EV_AER)] == [N3] and notes_of(notes, NOTE_CANCEL_SNAPSHOT_RETRY) == []
