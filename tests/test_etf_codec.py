"""AppendEntries wire codec (include/ra_etf.h, ra_b200/csrc/ra_etf.c): Erlang external term format of
#append_entries_rpc{} / {Peer, #append_entries_reply{}} (rabbitmq/ra src/ra.hrl:122-141) <-> ra_event.

No OTP toolchain exists here, so the vectors are derived BY HAND from the documented format (erts "External Term
Format": 131 = version, 104 = SMALL_TUPLE_EXT arity, 119 = SMALL_ATOM_UTF8_EXT len bytes, 97 = SMALL_INTEGER_EXT,
98 = INTEGER_EXT 4 bytes big endian, 110 = SMALL_BIG_EXT n sign little-endian bytes, 108 = LIST_EXT len(4) elems tail,
106 = NIL_EXT, 109 = BINARY_EXT len(4) bytes, 116 = MAP_EXT arity(4) pairs): what term_to_binary/1 of OTP 26+ emits.
A small independent Python encoder (below) produces the randomised round-trip inputs."""
import ctypes as C
import os
import random
import struct

import pytest

from ra_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EtfId(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("node", C.c_char * 256)]


class EtfEntry(C.Structure):
    _fields_ = [("index", C.c_uint64), ("term", C.c_uint64), ("cmd_off", C.c_uint32), ("cmd_len", C.c_uint32)]


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "ra_b200", "csrc", "libra_etf.so")
    assert os.path.exists(so), "build first: make -C ra_b200/csrc"
    l = C.CDLL(so)
    l.ra_etf_decode_aer.restype = C.c_int
    l.ra_etf_decode_aer.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(abi.RaEvent), C.POINTER(EtfId),
                                    C.POINTER(EtfEntry), C.c_size_t, C.POINTER(C.c_size_t)]
    l.ra_etf_encode_aer.restype = C.c_size_t
    l.ra_etf_encode_aer.argtypes = [C.POINTER(abi.RaEvent), C.POINTER(EtfId), C.POINTER(C.c_char_p),
                                    C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
    l.ra_etf_decode_aer_reply.restype = C.c_int
    l.ra_etf_decode_aer_reply.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(abi.RaEvent), C.POINTER(EtfId)]
    l.ra_etf_encode_aer_reply.restype = C.c_size_t
    l.ra_etf_encode_aer_reply.argtypes = [C.POINTER(abi.RaEvent), C.POINTER(EtfId), C.c_char_p, C.c_size_t]
    l.ra_etf_term_size.restype = C.c_size_t
    l.ra_etf_term_size.argtypes = [C.c_char_p, C.c_size_t]
    return l


# ---- the independent reference encoder (subset of term_to_binary/1) ------------------------------------------
class Atom(str):
    pass


def enc(t) -> bytes:
    if isinstance(t, Atom):
        b = t.encode()
        return bytes([119, len(b)]) + b
    if isinstance(t, bool):
        return enc(Atom("true" if t else "false"))
    if isinstance(t, int):
        if 0 <= t < 256:
            return bytes([97, t])
        if -(1 << 31) <= t < (1 << 31):
            return b"\x62" + struct.pack(">i", t)
        n = (t.bit_length() + 7) // 8
        return bytes([110, n, 0]) + t.to_bytes(n, "little")
    if isinstance(t, tuple):
        return bytes([104, len(t)]) + b"".join(enc(x) for x in t)
    if isinstance(t, list):
        if not t:
            return b"\x6a"
        return b"\x6c" + struct.pack(">I", len(t)) + b"".join(enc(x) for x in t) + b"\x6a"
    if isinstance(t, bytes):
        return b"\x6d" + struct.pack(">I", len(t)) + t
    if isinstance(t, dict):
        return b"\x74" + struct.pack(">I", len(t)) + b"".join(enc(k) + enc(v) for k, v in t.items())
    raise TypeError(t)


def t2b(t) -> bytes:
    return b"\x83" + enc(t)


def aer(term, leader, commit, prev_idx, prev_term, entries):
    return (Atom("append_entries_rpc"), term, (Atom(leader[0]), Atom(leader[1])), commit, prev_idx, prev_term, entries)


def usr(data: bytes):
    """{'$usr', Meta, Data, ReplyMode} (ra_server.erl command())"""
    return (Atom("$usr"), {Atom("ts"): 1700000000000}, data, Atom("noreply"))


# ---- hand-derived known answers ---------------------------------------------------------------------------------
HAND_EMPTY_AER = bytes([
    131, 104, 7,                                                   # version, 7-tuple
    119, 18]) + b"append_entries_rpc" + bytes([                    # record name
    97, 5,                                                         # term = 5
    104, 2, 119, 2]) + b"n1" + bytes([119, 3]) + b"a@b" + bytes([  # leader_id = {n1, 'a@b'}
    97, 3,                                                         # leader_commit = 3
    98, 0, 0, 1, 44,                                               # prev_log_index = 300 (INTEGER_EXT)
    97, 4,                                                         # prev_log_term = 4
    106])                                                          # entries = []

HAND_ONE_ENTRY = bytes([
    131, 104, 7, 119, 18]) + b"append_entries_rpc" + bytes([
    97, 7,                                                         # term 7
    104, 2, 119, 2]) + b"n2" + bytes([119, 3]) + b"x@y" + bytes([
    97, 9,                                                         # leader_commit 9
    97, 9,                                                         # prev_log_index 9
    97, 6,                                                         # prev_log_term 6
    108, 0, 0, 0, 1,                                               # list of 1
    104, 3, 97, 10, 97, 7,                                         # {10, 7,
    109, 0, 0, 0, 2, 104, 105,                                     #  <<"hi">>}
    106])

HAND_REPLY = bytes([
    131, 104, 2,
    104, 2, 119, 2]) + b"n3" + bytes([119, 3]) + b"a@b" + bytes([  # {n3, 'a@b'}
    104, 6, 119, 20]) + b"append_entries_reply" + bytes([
    97, 5,                                                         # term
    119, 4]) + b"true" + bytes([                                   # success
    110, 5, 0, 1, 0, 0, 0, 1,                                      # next_index = 2^32 + 1 (SMALL_BIG_EXT, 5 bytes)
    97, 200,                                                       # last_index
    97, 5])                                                        # last_term


def test_reference_encoder_matches_hand_vectors():
    assert t2b(aer(5, ("n1", "a@b"), 3, 300, 4, [])) == HAND_EMPTY_AER
    assert t2b(aer(7, ("n2", "x@y"), 9, 9, 6, [(10, 7, b"hi")])) == HAND_ONE_ENTRY
    assert t2b(((Atom("n3"), Atom("a@b")), (Atom("append_entries_reply"), 5, True, (1 << 32) + 1, 200, 5))) == HAND_REPLY


def _decode(lib, blob, max_entries=256):
    ev, ld = abi.RaEvent(), EtfId()
    ents = (EtfEntry * max_entries)()
    n = C.c_size_t(0)
    rc = lib.ra_etf_decode_aer(blob, len(blob), C.byref(ev), C.byref(ld), ents, max_entries, C.byref(n))
    return rc, ev, ld, list(ents[: n.value])


def test_decode_hand_vectors(lib):
    rc, ev, ld, ents = _decode(lib, HAND_EMPTY_AER)
    assert rc == 0 and (ev.type, ev.term, ev.a, ev.b, ev.c, ev.n, ev.n1) == (abi.EV_AER, 5, 300, 4, 3, 0, 0)
    assert (ld.name, ld.node) == (b"n1", b"a@b") and ents == []
    rc, ev, ld, ents = _decode(lib, HAND_ONE_ENTRY)
    assert rc == 0 and (ev.term, ev.a, ev.b, ev.c, ev.n, ev.n1, ev.d, ev.e) == (7, 9, 6, 9, 1, 0, 7, 0)
    (e,) = ents
    assert (e.index, e.term) == (10, 7) and HAND_ONE_ENTRY[e.cmd_off:e.cmd_off + e.cmd_len] == enc(b"hi")
    ev, peer = abi.RaEvent(), EtfId()
    assert lib.ra_etf_decode_aer_reply(HAND_REPLY, len(HAND_REPLY), C.byref(ev), C.byref(peer)) == 0
    assert (ev.type, ev.term, ev.d, ev.a, ev.b, ev.c) == (abi.EV_AER_REPLY, 5, 1, (1 << 32) + 1, 200, 5)
    assert (peer.name, peer.node) == (b"n3", b"a@b")


def test_encode_reproduces_the_bytes(lib):
    for blob in (HAND_EMPTY_AER, HAND_ONE_ENTRY):
        rc, ev, ld, ents = _decode(lib, blob)
        cmds = (C.c_char_p * max(1, len(ents)))(*[blob[e.cmd_off:e.cmd_off + e.cmd_len] for e in ents])
        lens = (C.c_uint32 * max(1, len(ents)))(*[e.cmd_len for e in ents])
        need = lib.ra_etf_encode_aer(C.byref(ev), C.byref(ld), cmds, lens, None, 0)
        out = C.create_string_buffer(need)
        assert lib.ra_etf_encode_aer(C.byref(ev), C.byref(ld), cmds, lens, out, need) == need
        assert out.raw == blob
    ev, peer = abi.RaEvent(), EtfId()
    lib.ra_etf_decode_aer_reply(HAND_REPLY, len(HAND_REPLY), C.byref(ev), C.byref(peer))
    out = C.create_string_buffer(256)
    n = lib.ra_etf_encode_aer_reply(C.byref(ev), C.byref(peer), out, 256)
    assert out.raw[:n] == HAND_REPLY


def test_round_trip_random_batches(lib):
    rnd = random.Random(5)
    for _ in range(300):
        prev = rnd.choice([0, 1, 255, 256, 70000, (1 << 31) - 1, 1 << 31, (1 << 40) + 3, (1 << 63) + 9])
        term = rnd.choice([1, 7, 255, 256, 1 << 33])
        n = rnd.randrange(0, 40)
        n1 = rnd.randrange(0, n) if n > 1 and rnd.random() < 0.4 else 0
        t_old = max(1, term - rnd.randrange(1, 3)) if n1 else term
        ents = [(prev + 1 + i, t_old if (n1 and i < n1) else term, usr(bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 50)))))
                for i in range(n)]
        blob = t2b(aer(term, ("ra_%d" % rnd.randrange(100), "rabbit@host-%d" % rnd.randrange(9)), prev + rnd.randrange(0, 3), prev,
                       rnd.choice([0, term, t_old]), ents))
        rc, ev, ld, got = _decode(lib, blob)
        assert rc == 0
        assert ev.n == n and (ev.n1 == n1 or (n1 and t_old == term and ev.n1 == 0))
        for (idx, tm, cmd), e in zip(ents, got):
            assert (e.index, e.term) == (idx, tm) and blob[e.cmd_off:e.cmd_off + e.cmd_len] == enc(cmd)
            assert lib.ra_etf_term_size(blob[e.cmd_off:], len(blob) - e.cmd_off) == e.cmd_len
        cmds = (C.c_char_p * max(1, n))(*[blob[e.cmd_off:e.cmd_off + e.cmd_len] for e in got])
        lens = (C.c_uint32 * max(1, n))(*[e.cmd_len for e in got])
        need = lib.ra_etf_encode_aer(C.byref(ev), C.byref(ld), cmds, lens, None, 0)
        out = C.create_string_buffer(need)
        lib.ra_etf_encode_aer(C.byref(ev), C.byref(ld), cmds, lens, out, need)
        assert out.raw == blob


def test_rejects_what_the_engine_record_cannot_hold(lib):
    three_runs = t2b(aer(9, ("n1", "a@b"), 0, 0, 0, [(1, 1, b"a"), (2, 2, b"b"), (3, 3, b"c")]))
    assert _decode(lib, three_runs)[0] == -5                       # RA_ETF_E_RUNS: split the batch
    gap = t2b(aer(9, ("n1", "a@b"), 0, 0, 0, [(1, 1, b"a"), (3, 1, b"c")]))
    assert _decode(lib, gap)[0] == -5
    assert _decode(lib, HAND_ONE_ENTRY[:-4])[0] == -1              # truncated
    assert _decode(lib, t2b((Atom("request_vote_rpc"), 1, 2, 3, 4, 5, 6)))[0] == -2
    neg = t2b(aer(-1, ("n1", "a@b"), 0, 0, 0, []))
    assert _decode(lib, neg)[0] == -3                              # RA_ETF_E_RANGE
    assert _decode(lib, t2b(aer(1, ("n1", "a@b"), 0, 0, 0, [(i + 1, 1, b"") for i in range(9)])), max_entries=8)[0] == -4
