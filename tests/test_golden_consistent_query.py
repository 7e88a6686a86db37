"""Golden vectors for the consistent-query heartbeat round: the heartbeat cases of
/root/reference/test/ra_server_SUITE.erl (:3232-3725, :3805-3840) restated.  The query refs and the
`pending_consistent_queries` list stay on the host (it submits a query only while cluster changes are
permitted); the engine owns the indexes: RA_NOTE_QUERY_INDEX tells the host which index its query got,
RA_NOTE_QUERY_AGREED which index a quorum has confirmed (the host applies every waiting query <= it; the
reference emits the replies itself and stays silent when nothing waits)."""
import pytest

from ra_suite import *  # noqa: F401,F403

BACKENDS = ["oracle", "emu", pytest.param("engine", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def be(request):
    return request.param


def run(nd, role, ev, st, query_index=0, peers=None, agreed=0):
    s = clone(st)
    s.role = role
    ev.row = s.row
    nd.b.load_rows([s])
    q = abi.RaQueryState(row=s.row, query_index=query_index, agreed_index=agreed)
    for p, v in (peers or {}).items():
        q.peer_query_index[p] = v
    nd.b.load_query_state([q])
    msgs, notes = nd.b.step([ev])
    out = nd.b.read_rows([s.row])[0]
    return out.role, out, msgs, notes, nd.b.read_query_state([s.row])[0]


def hb_reply(m):
    assert m.type == abi.EV_HEARTBEAT_REPLY
    return (m.row, m.from_slot, m.term, m.a)


def test_follower_heartbeat(be):
    """follower_heartbeat/1 :3232-3271."""
    nd = Node(be, 3)
    st = base_state(3)                        # n1, term 5; the rpc names n1 as leader (as the reference does)
    role, s, msgs, _, _ = run(nd, FOLLOWER, abi.ev_heartbeat_rpc(0, N1, 4, 1), st)      # lower term
    assert role == FOLLOWER and s.current_term == 5 and [hb_reply(m) for m in msgs] == [(N1, N1, 5, 1)]
    role, s, msgs, _, _ = run(nd, FOLLOWER, abi.ev_heartbeat_rpc(0, N1, 5, 1), st)      # same term
    assert role == FOLLOWER and s.current_term == 5 and [hb_reply(m) for m in msgs] == [(N1, N1, 5, 1)]
    role, s, msgs, _, _ = run(nd, FOLLOWER, abi.ev_heartbeat_rpc(0, N2, 6, 1), st)      # higher term
    assert (role, s.current_term, s.voted_for, s.leader_slot) == (FOLLOWER, 6, abi.RA_NO_SLOT, N2)
    assert [hb_reply(m) for m in msgs] == [(N2, N1, 6, 1)]


def test_follower_heartbeat_reply(be):
    """follower_heartbeat_reply/1 :3273-3289."""
    nd = Node(be, 3)
    st = base_state(3)
    for term in (5, 4):
        role, s, msgs, _, _ = run(nd, FOLLOWER, abi.ev_heartbeat_reply(0, N2, term, 2), st)
        assert (role, s.current_term, msgs) == (FOLLOWER, 5, [])
    role, s, msgs, _, _ = run(nd, FOLLOWER, abi.ev_heartbeat_reply(0, N2, 6, 2), st)
    assert (role, s.current_term, s.voted_for, msgs) == (FOLLOWER, 6, abi.RA_NO_SLOT, [])


@pytest.mark.parametrize("role", [CANDIDATE, PRE_VOTE])
def test_candidate_and_pre_vote_heartbeat(be, role):
    """candidate_heartbeat/1 :3291-3323, pre_vote_heartbeat/1 :3371-3403: same or higher term -> follower and
    the rpc is handled again there; lower term -> reply with the own term."""
    nd = Node(be, 3)
    st = base_state(3)
    st.votes = 1
    votes = 1 if role == CANDIDATE else 0           # only handle_pre_vote resets the votes (:1186)
    r, s, msgs, _, _ = run(nd, role, abi.ev_heartbeat_rpc(0, N2, 5, 1), st)
    assert (r, s.current_term, s.votes) == (FOLLOWER, 5, votes) and len(next_events(msgs)) == 1
    r, s, msgs, _, _ = run(nd, role, abi.ev_heartbeat_rpc(0, N2, 6, 1), st)
    assert (r, s.current_term, s.votes, s.voted_for) == (FOLLOWER, 6, votes, abi.RA_NO_SLOT) and len(next_events(msgs)) == 1
    r, s, msgs, _, _ = run(nd, role, abi.ev_heartbeat_rpc(0, N2, 4, 1), st)
    assert (r, s.current_term) == (role, 5) and [hb_reply(m) for m in sent(msgs)] == [(N2, N1, 5, 1)]


@pytest.mark.parametrize("role", [CANDIDATE, PRE_VOTE])
def test_candidate_and_pre_vote_heartbeat_reply(be, role):
    """candidate_heartbeat_reply/1 :3325-3369, pre_vote_heartbeat_reply/1 :3405-3429."""
    nd = Node(be, 3)
    st = base_state(3)
    for term in (5, 4):
        r, s, msgs, _, _ = run(nd, role, abi.ev_heartbeat_reply(0, N2, term, 2), st)
        assert (r, s.current_term, msgs) == (role, 5, [])
    r, s, msgs, _, _ = run(nd, role, abi.ev_heartbeat_reply(0, N2, 6, 2), st)
    assert (r, s.current_term, s.voted_for, msgs) == (FOLLOWER, 6, abi.RA_NO_SLOT, [])


def test_leader_heartbeat(be):
    """leader_heartbeat/1 :3431-3469."""
    nd = Node(be, 3)
    st = base_state(3)
    r, s, msgs, notes, _ = run(nd, LEADER, abi.ev_heartbeat_rpc(0, N2, 5, 1), st)       # same term: the reference exits
    assert status(notes) & ST_FATAL
    (sn,) = notes_of(notes, NOTE_STATUS)
    assert sn.c == abi.FATAL_LEADER_SAW_HEARTBEAT_SAME_TERM
    r, s, msgs, _, _ = run(nd, LEADER, abi.ev_heartbeat_rpc(0, N2, 6, 1), st)           # higher term
    assert (r, s.current_term, s.leader_slot, s.voted_for) == (FOLLOWER, 6, abi.RA_NO_SLOT, abi.RA_NO_SLOT)
    assert len(next_events(msgs)) == 1
    r, s, msgs, _, _ = run(nd, LEADER, abi.ev_heartbeat_rpc(0, N2, 4, 1), st)           # lower term
    assert (r, s.current_term) == (LEADER, 5) and [hb_reply(m) for m in msgs] == [(N2, N1, 5, 1)]


def test_leader_heartbeat_reply_node_size_5(be):
    """leader_heartbeat_reply_node_size_5/1 :3471-3494: one reply of four peers is no quorum, two are."""
    nd = Node(be, 5)
    st = base_state(5)
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 5, 2), st, query_index=2)
    assert r == LEADER and notes_of(notes, NOTE_QUERY_AGREED) == [] and q.peer_query_index[N2] == 2
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N3, 5, 2), st, query_index=2, peers={N2: 2})
    (ag,) = notes_of(notes, NOTE_QUERY_AGREED)
    assert ag.a == 2 and q.agreed_index == 2


def test_leader_heartbeat_reply_same_term(be):
    """leader_heartbeat_reply_same_term/1 :3496-3579."""
    nd = Node(be, 3)
    st = base_state(3)
    QI = 2
    # the reply updates the peer's query index; a single reply in a 3-member group is a quorum for QI
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 5, QI), st, query_index=QI + 1)
    assert r == LEADER and q.peer_query_index[N2] == QI
    assert [n.a for n in notes_of(notes, NOTE_QUERY_AGREED)] == [QI]
    # unknown peer: ignored (no peer cell, quorum unchanged)
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, 7, 5, QI), st, query_index=QI + 1)
    assert r == LEADER and list(q.peer_query_index)[:3] == [0, 0, 0] and notes_of(notes, NOTE_QUERY_AGREED) == []
    # a lower index does not confirm the waiting query (index QI)
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 5, QI - 1), st, query_index=QI + 1)
    assert q.peer_query_index[N2] == QI - 1 and [n.a for n in notes_of(notes, NOTE_QUERY_AGREED)] == [QI - 1]
    # two waiting queries (QI, QI+1): the reply for QI confirms one, the reply for QI+1 both
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 5, QI + 1), st, query_index=QI + 1)
    assert [n.a for n in notes_of(notes, NOTE_QUERY_AGREED)] == [QI + 1] and q.agreed_index == QI + 1
    # nothing new is confirmed twice
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 5, QI), st, query_index=QI + 1,
                               peers={N2: QI}, agreed=QI)
    assert notes_of(notes, NOTE_QUERY_AGREED) == []


def test_leader_heartbeat_reply_other_terms(be):
    """leader_heartbeat_reply_lower_term/1 :3805-3820, leader_heartbeat_reply_higher_term/1 :3822-3840."""
    nd = Node(be, 3)
    st = base_state(3)
    for qi in (0, 1):
        r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 4, qi), st)
        assert (r, s.current_term, msgs) == (LEADER, 5, []) and q.peer_query_index[N2] == 0
        r, s, msgs, notes, q = run(nd, LEADER, abi.ev_heartbeat_reply(0, N2, 6, qi), st)
        assert (r, s.current_term, s.voted_for, s.leader_slot, msgs) == (FOLLOWER, 6, abi.RA_NO_SLOT, abi.RA_NO_SLOT, [])


def test_leader_consistent_query(be):
    """leader_consistent_query/1 :3635-3672: each query takes the next query index and a heartbeat round."""
    nd = Node(be, 3)
    st = base_state(3)
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_consistent_query(0), st)
    assert r == LEADER and q.query_index == 1
    assert [(m.type, m.row, m.from_slot, m.term, m.a) for m in msgs] == [(abi.EV_HEARTBEAT_RPC, N2, N1, 5, 1),
                                                                        (abi.EV_HEARTBEAT_RPC, N3, N1, 5, 1)]
    (qn,) = notes_of(notes, NOTE_QUERY_INDEX)
    assert (qn.a, qn.b) == (1, 3)                                   # its index, the commit index it reads at
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_consistent_query(0), st, query_index=1)
    assert q.query_index == 2 and [(m.row, m.a) for m in msgs] == [(N2, 2), (N3, 2)]
    # a peer that already confirmed this index, or is not `normal`, gets no rpc (:3756-3771)
    st2 = clone(st)
    st2.peers[N3].status = PEER_DISCONNECTED
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_consistent_query(0), st2, query_index=1, peers={N2: 2})
    assert msgs == []
    # not the leader: the host redirects
    for role in (FOLLOWER, CANDIDATE, PRE_VOTE, AWAIT_CONDITION):
        r, s, msgs, notes, q = run(nd, role, abi.ev_consistent_query(0), st)
        assert r == role and msgs == [] and len(notes_of(notes, NOTE_NOT_LEADER)) == 1 and q.query_index == 0


def test_single_member_applies_at_once(be):
    """make_heartbeat_rpc_effects/2 with no peers :3729-3731."""
    nd = Node(be, 1)
    st = base_state(1)
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_consistent_query(0), st)
    assert msgs == [] and len(notes_of(notes, NOTE_QUERY_APPLY)) == 1 and q.query_index == 0


def test_await_condition_drops_heartbeats(be):
    """await_condition_heartbeat_dropped/1 :3684-3702, ..._reply_dropped/1 :3704-3722."""
    nd = Node(be, 3)
    st = base_state(3)
    st.condition = 1
    for term in (5, 6, 4):
        for ev in (abi.ev_heartbeat_rpc(0, N2, term, 1), abi.ev_heartbeat_reply(0, N2, term, 1)):
            r, s, msgs, notes, q = run(nd, AWAIT_CONDITION, ev, st)
            assert (r, s.current_term, msgs) == (AWAIT_CONDITION, 5, [])


def test_term_change_resets_peer_query_indexes(be):
    """reset_query_index/1 :3743-3747 from update_term_and_voted_for/3 :3029; tick re-sends the heartbeats
    of the current index to peers that have not confirmed it (update_heartbeat_rpc_effects/1 :3704-3720)."""
    nd = Node(be, 3)
    st = base_state(3)
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_request_vote(0, N2, 6, 3, 5), st, query_index=4, peers={N2: 4, N3: 3})
    assert r == FOLLOWER and (q.query_index, list(q.peer_query_index)[:3]) == (4, [0, 0, 0])
    r, s, msgs, notes, q = run(nd, LEADER, abi.ev_simple(0, abi.EV_TICK), st, query_index=4, peers={N2: 4, N3: 3})
    assert [(m.row, m.a) for m in of_type(msgs, abi.EV_HEARTBEAT_RPC)] == [(N3, 4)]
