// Row-level pieces shared by the two step kernels (engine.cu) and by the host-side emulation
// that the CPU tests use to run this very device logic without a GPU (tests/emu/): a row's step
// context, loading / writing back its registers, its end of step, and the conversions between a
// row of the column store and the ABI's ra_row_state.
#pragma once
#include "raft_step.cuh"

struct StallCtx {                      // 4 x 16 bytes
    u32 row, flags, rem_mbox, rem_loc;
    u32 n_msgs_notes, status, sent_to, pn_type_slot_wk;
    u64 pn_a, pn_b;
    u64 pn_c, _pad;
};
#define STALL_PENDING 1u               // the deferred pipeline pass has not run yet

__device__ __forceinline__ void member_init(Member& m, const Cols& C, u32 r, ulonglong2 tc, ulonglong2 lg, ulonglong2 lw,
                                            ulonglong2 ap, u64 lrs, int cur, ulonglong2* sp)
{
    m.C = &C; m.row = r;
    {   // slot = r / groups through the precomputed reciprocal floor(2^32 / groups): off by at most one
        u32 q = __umulhi(r, C.groups_inv), rem = r - q * C.groups;
        if (rem >= C.groups) { q++; rem -= C.groups; }
        if (rem >= C.groups) { q++; rem -= C.groups; }
        m.slot = q; m.group = rem;
    }
    m.term = tc.x; m.commit = tc.y; m.last_idx = lg.x; m.last_term = lg.y;
    m.lw_idx = lw.x; m.lw_term = lw.y; m.applied = ap.x; m.meta = ap.y;
    m.cold = 0;
    m.lrs = lrs; m.lrs_ok = MT_NRUNS(ap.y) ? 1u : 0u;
    m.n_msgs = 0; m.n_notes = 0; m.status = MT_ROLE(ap.y) << 16; m.wk = 0;
    m.sent_to = 0; m.pn_type = RA_NOTE_NONE; m.pn_slot = 0; m.pn_a = m.pn_b = m.pn_c = 0;
    m.c_pack = 0; m.c_ref = 0; m.c_commits = m.c_applied = 0;
    m.nb = cur ^ 1;
#ifdef RA_HOST_EMU
    m.sp = sp;
#else
    m.sp = (u32)__cvta_generic_to_shared(sp);
#endif
    m.pstate = 0; m.pipe_clean = 0;
}

// The four hot pairs change on practically every step of an active row (commit_index,
// last_index, last_written, last_applied / meta), so they are stored unconditionally: keeping
// their loaded values around just to skip a store costs 16 registers per thread.
__device__ __forceinline__ void member_writeback(const Member& m, const Cols& C, u32 r)
{
    st2(&C.tc[r], m.term, m.commit);
    st2(&C.lg[r], m.last_idx, m.last_term);
    st2(&C.lw[r], m.lw_idx, m.lw_term);
    st2(&C.ap[r], m.applied, m.meta);
    lrs_writeback(m);
}

// ---- flood fault injection (include/ra_engine.h, ra_flood_faults): keyed by GLOBAL ids ---------------------
__device__ __forceinline__ void flood_ids(const Cols& C, const Member& m, u32 r, u64& gg, u64& gr)
{
    gg = m.group; gr = r;
    if (C.n_shards > 1) {
        gg = (u64)C.n_shards * m.group + (C.shard + 8u * C.n_shards - m.slot) % C.n_shards;
        gr = (u64)m.slot * C.groups * C.n_shards + gg;
    }
}
// is the mailbox record `e` lost before row r evaluates it in step F.step?
template <int MM>
__device__ __forceinline__ bool flood_lost(const FloodArgs& F, const Cols& C, const Member& m, u32 r, const Rec& e)
{
    if (!(F.drop | F.part)) return false;
    u64 gg, gr; flood_ids(C, m, r, gg, gr);
    const u32 from = R_from(e);
    if (F.part) {                                            // one member of the group is cut off in this window
        const u64 w = F.step / F.part_len;
        const u32 h = (u32)(mix64(F.seed ^ (w * 0xC2B2AE3D27D4EB4Full) ^ (gg * 0x165667B19E3779F9ull)) >> 32);
        if (h % 1000u < F.part) { const u32 p = (h / 1000u) % NMEM(C); if (p == m.slot || p == from) return true; }
    }
    if (F.drop && R_type(e) == RA_EV_AER) {
        const u32 h = (u32)(mix64(F.seed ^ (F.step * 0x9E3779B97F4A7C15ull) ^ (gr * 0xD6E8FEB86659FD93ull) ^ ((u64)(from + 1) << 56)) >> 32);
        if (h % 1000u < F.drop) return true;
    }
    return false;
}

// end of a row's step: publish mailbox counts, STATUS note, output counts, flood host model
// FAULTS = false compiles the fault injection out (the specialisations of the hot kernel that run the plain flood)
template <int MM, bool FAULTS = true>
__device__ __forceinline__ u32 row_end_of_step(Member& m, const Cols& C, u32 r, int cur, const FloodArgs& F)
{
    u32 fatal = 0;
    const bool routed = MTR == TR_RUNTIME ? (C.routed != 0) : (MTR != TR_HOST);
    if (routed) {
        const bool sharded = MTR == TR_RUNTIME ? (C.n_shards > 1) : (MTR == TR_PEER || MTR == TR_BUCKET);
        const bool peer = MTR == TR_RUNTIME ? (C.peer_mode != 0) : (MTR == TR_PEER);
        for (u32 s = 0; s < NMEM(C); s++) {
            if (s == m.slot) continue;
            u64* cnt = C.mbox_cnt[cur ^ 1];
            if (sharded) {
                const u32 ds = (C.shard + s + 8u * C.n_shards - m.slot) % C.n_shards;
                if (peer) cnt = C.peer_cnt[cur ^ 1][ds];                 // byte store over NVLink
                else if (ds != C.shard) continue;   // set when the records are delivered (deliver_kernel)
            }
            // (a count byte is zero when its buffer comes round again -- the owner clears the word when it consumes
            // it -- so only senders that sent something have to publish: 1 of a follower's 4 bytes in steady state)
            const u32 nsent = (m.sent_to >> (4 * s)) & 15u;
            if (nsent) reinterpret_cast<u8*>(&cnt[(size_t)s * C.groups + m.group])[m.slot] = (u8)nsent;
        }
    }
    // record_leader_msg alone (the steady state of a follower) does not get a STATUS note of its own: the
    // flags ride in the aux field of the row's last note of the step (include/ra_engine.h, RA_NOTE_STATUS)
    const u32 st16 = m.status & 0xffffu;
    const bool elide = st16 == RA_ST_LEADER_MSG && m.pn_type != RA_NOTE_NONE;
    note_flush(m, elide ? st16 : 0u);
    if (st16 && !elide) {
        u64 ld = MT_LEADER(m.meta), vf = MT_VOTED(m.meta);
        u64 b = (vf == SLOT_NONE ? 0xFFull : vf) | ((ld == SLOT_NONE ? 0xFFull : ld) << 8) |
                ((u64)((m.status >> 16) & 7u) << 16) | ((u64)MT_ROLE(m.meta) << 24);
        note_store(m, m.n_notes, RA_NOTE_STATUS, m.slot, m.status & 0xffffu, m.term, b,
                   ((m.status >> 20) & 0xffu) | ((u64)(m.status >> 28) << 8));
        m.n_notes++;
        if (m.status & RA_ST_FATAL) fatal = 1;
    }
    C.out_n[r] = m.n_msgs | (m.n_notes << 16);
    // flood: synthetic host (DESIGN.md "flood host model")
    if (F.on && !MT_FATAL(m.meta)) {
        u32 k = 0;
        bool held = false;                                  // a lagging fsync: this step's notifications are not produced
        if (FAULTS && F.withhold && (m.wk & 3u)) {
            u64 gg0, gr0; flood_ids(C, m, r, gg0, gr0);
            held = (u32)(mix64(F.seed ^ (F.step * 0xA0761D6478BD642Full) ^ (gr0 * 0xE7037ED1A0B428DBull)) >> 32) % 1000u < F.withhold;
        }
        if (held) m.wk = 0;
        // {written, Term, {From, To}} for the (last two) WAL_APPEND notes of this step: read back
        // from the row's own note slots instead of being carried in registers through the step
        if ((m.wk & 3u) == 2) {
            const ulonglong2* q = reinterpret_cast<const ulonglong2*>(&C.onote[(size_t)((m.wk >> 8) & 15u) * C.rows + r]);
            const ulonglong2 h = q[0], t = q[1];
            put_local(C.loc, C.tiles, k, r, RA_EV_WRITTEN, 0, t.y, h.y, t.x); k++;
        }
        if ((m.wk & 3u) >= 1) {
            const ulonglong2* q = reinterpret_cast<const ulonglong2*>(&C.onote[(size_t)((m.wk >> 4) & 15u) * C.rows + r]);
            const ulonglong2 h = q[0], t = q[1];
            put_local(C.loc, C.tiles, k, r, RA_EV_WRITTEN, 0, t.y, h.y, t.x); k++;
        }
        const u32 role = MT_ROLE(m.meta);
        if (role == RA_LEADER && F.cmds) { put_local(C.loc, C.tiles, k, r, RA_EV_COMMAND, F.cmds, 0, 0, 0); k++; }
        u32 idle = MT_IDLE(m.meta);
        if (role == RA_LEADER || (m.status & RA_ST_LEADER_MSG)) idle = 0;
        else if (idle < 15) idle++;
        bool fire = false;
        if (role != RA_LEADER) {
            // the model is keyed by GLOBAL group / row ids so that a sharded run equals the unsharded one
            u64 gg = m.group, gr = r;
            if (C.n_shards > 1) {
                gg = (u64)C.n_shards * m.group + (C.shard + 8u * C.n_shards - m.slot) % C.n_shards;
                gr = (u64)m.slot * C.groups * C.n_shards + gg;
            }
            if (F.permille) {
                const u32 h = (u32)(mix64(F.seed ^ (F.step * 0x9E3779B97F4A7C15ull) ^ (gg * 0xD1B54A32D192ED03ull)) >> 32);
                if ((h % 1000u) < F.permille && ((h / 1000u) % NMEM(C)) == m.slot) fire = true;
            }
            if (idle >= 8) {                                // the hash only matters from 8 idle steps on
                const u32 h2 = (u32)(mix64(F.seed ^ (gr * 0xA24BAED4963EE407ull) ^ F.step) >> 32);
                if (idle >= 8 + (h2 & 7u)) fire = true;
            }
        }
        if (fire) { put_local(C.loc, C.tiles, k, r, RA_EV_ELECTION_TIMEOUT, 0, 0, 0, 0); k++; idle = 0; }
        MT_SET(m.meta, 28, 4, idle);
        C.loc_n[r] = k;
    }
    return fatal;
}


// ---- row <-> ra_row_state (load_rows / read_rows / reset_empty) -------------------------------
__device__ __forceinline__ void reset_row(const Cols& C, const u32 r)
{
    // ra_server_SUITE:empty_state/2: term 0, log {0 => 0}, peers next=1 match=0, all voters
    u64 meta = 0;
    MT_SET(meta, 0, 3, RA_FOLLOWER); MT_SET(meta, 3, 4, SLOT_NONE); MT_SET(meta, 7, 4, SLOT_NONE);
    MT_SET(meta, 19, 4, 1); MT_SET(meta, 27, 1, 1);
    MT_SET(meta, 56, 8, (1u << C.members) - 1u);
    st2(&C.tc[r], 0, 0); st2(&C.lg[r], 0, 0); st2(&C.lw[r], 0, 0); st2(&C.ap[r], 0, meta);
    st2(&C.sn[r], 0, 0); st2(&C.tk[r], 0, 0); st2(&C.fm[r], 0, 0);
    st2(&C.cd[r], 0, 0); st2(&C.cd[(size_t)C.rows + r], 0, 0);
    for (u32 s = 0; s < C.members; s++) { st2(&C.pnm[(size_t)s * C.rows + r], 1, 0); C.pcs[(size_t)s * C.rows + r] = 0; }
    for (u32 k = 0; k < RA_MAX_RUNS; k++) st2(&C.run[(size_t)k * C.rows + r], 0, 0);
    C.lrs[r] = 0;
    C.qi[r] = 0; C.qa[r] = 0; C.wc[r] = 0;
    for (u32 s = 0; s < C.members; s++) C.pqi[(size_t)s * C.rows + r] = 0;
    C.loc_n[r] = 0; C.out_n[r] = 0;
    if (C.routed) { C.mbox_cnt[0][r] = 0; C.mbox_cnt[1][r] = 0; }
}

__device__ __forceinline__ void load_row(const Cols& C, const ra_row_state& s)
{
    const u32 r = s.row;
    u64 meta = 0;
    MT_SET(meta, 0, 3, s.role);
    MT_SET(meta, 3, 4, s.leader_slot == RA_NO_SLOT ? SLOT_NONE : s.leader_slot);
    MT_SET(meta, 7, 4, s.voted_for == RA_NO_SLOT ? SLOT_NONE : s.voted_for);
    MT_SET(meta, 11, 2, s.membership); MT_SET(meta, 13, 2, s.condition);
    MT_SET(meta, 15, 4, s.votes); MT_SET(meta, 19, 4, s.n_runs);
    MT_SET(meta, 23, 1, s.has_snapshot ? 1 : 0);
    MT_SET(meta, 24, 1, (s.flags & 1) ? 1 : 0); MT_SET(meta, 25, 1, (s.flags & 2) ? 1 : 0);
    MT_SET(meta, 26, 1, (s.flags & 4) ? 1 : 0);
    MT_SET(meta, 27, 1, s.machine_version >= s.effective_machine_version ? 1 : 0);
    u32 voters = 0;
    for (u32 p = 0; p < RA_MAX_MEMBERS; p++) {
        MT_SET(meta, 32 + 3 * p, 3, s.peers[p].status);
        if (s.peers[p].voter) voters |= 1u << p;
    }
    MT_SET(meta, 56, 8, voters);
    st2(&C.tc[r], s.current_term, s.commit_index);
    st2(&C.lg[r], s.last_index, s.last_term);
    st2(&C.lw[r], s.last_written_index, s.last_written_term);
    st2(&C.ap[r], s.last_applied, meta);
    st2(&C.sn[r], s.snapshot_index, s.snapshot_term);
    st2(&C.tk[r], s.pre_vote_token, s.token_counter);
    st2(&C.fm[r], s.first_index, (u64)s.machine_version | ((u64)s.effective_machine_version << 32));
    st2(&C.cd[r], s.cond_reply_term, s.cond_reply_next_index);
    st2(&C.cd[(size_t)C.rows + r], s.cond_reply_last_index, s.cond_reply_last_term);
    for (u32 p = 0; p < C.members; p++) {
        st2(&C.pnm[(size_t)p * C.rows + r], s.peers[p].next_index, s.peers[p].match_index);
        C.pcs[(size_t)p * C.rows + r] = s.peers[p].commit_index_sent;
    }
    for (u32 k = 0; k < RA_MAX_RUNS; k++)
        st2(&C.run[(size_t)k * C.rows + r], k < s.n_runs ? s.run_start[k] : 0, k < s.n_runs ? s.run_term[k] : 0);
    C.lrs[r] = s.n_runs ? s.run_start[s.n_runs - 1] : 0;
    C.qi[r] = 0; C.qa[r] = 0; C.wc[r] = 0;
    for (u32 p = 0; p < C.members; p++) C.pqi[(size_t)p * C.rows + r] = 0;
    C.loc_n[r] = 0;
}

__device__ __forceinline__ void read_row(const Cols& C, ra_row_state& s)
{
    const u32 r = s.row;
    const ulonglong2 tc = C.tc[r], lg = C.lg[r], lw = C.lw[r], ap = C.ap[r], sn = C.sn[r], tk = C.tk[r], fm = C.fm[r];
    const u64 meta = ap.y;
    s.role = MT_ROLE(meta); s.self_slot = r / C.groups; s.n_members = C.members;
    u32 l = MT_LEADER(meta), v = MT_VOTED(meta);
    s.leader_slot = l == SLOT_NONE ? RA_NO_SLOT : l; s.voted_for = v == SLOT_NONE ? RA_NO_SLOT : v;
    s.membership = MT_MEMBERSHIP(meta); s.condition = MT_COND(meta); s.has_snapshot = MT_HAS_SNAP(meta);
    s.votes = MT_VOTES(meta); s.machine_version = (u32)fm.y; s.effective_machine_version = (u32)(fm.y >> 32);
    s.n_runs = MT_NRUNS(meta);
    s.flags = MT_PIPE_PEND(meta) | (MT_COND_VALID(meta) << 1) | (MT_FATAL(meta) << 2);
    s.current_term = tc.x; s.commit_index = tc.y; s.last_applied = ap.x;
    s.pre_vote_token = tk.x; s.token_counter = tk.y;
    s.first_index = fm.x; s.last_index = lg.x; s.last_term = lg.y;
    s.last_written_index = lw.x; s.last_written_term = lw.y;
    s.snapshot_index = sn.x; s.snapshot_term = sn.y;
    for (u32 k = 0; k < RA_MAX_RUNS; k++) {
        ulonglong2 rr = C.run[(size_t)k * C.rows + r];
        s.run_start[k] = k < s.n_runs ? rr.x : 0; s.run_term[k] = k < s.n_runs ? rr.y : 0;
    }
    ulonglong2 c0 = C.cd[r], c1 = C.cd[(size_t)C.rows + r];
    s.cond_reply_term = c0.x; s.cond_reply_next_index = c0.y; s.cond_reply_last_index = c1.x; s.cond_reply_last_term = c1.y;
    for (u32 p = 0; p < RA_MAX_MEMBERS; p++) {
        ra_peer_init& pi = s.peers[p];
        if (p < C.members) {
            ulonglong2 nm = C.pnm[(size_t)p * C.rows + r];
            pi.next_index = nm.x; pi.match_index = nm.y; pi.commit_index_sent = C.pcs[(size_t)p * C.rows + r];
            pi.status = MT_PSTATUS(meta, p); pi.voter = MT_VOTER(meta, p);
        } else { pi.next_index = pi.match_index = pi.commit_index_sent = 0; pi.status = 0; pi.voter = 0; }
        for (int q = 0; q < 6; q++) pi._pad[q] = 0;
    }
}

// ---- bucket transport: one record another shard sent to a member of this engine ---------------
// -> mailbox plane of the next step.  The slot is fixed by the record itself (sender slot, k-th
// record of that sender for this row); byte `from` of the row's count word becomes
// max(old count, k + 1) | tail flag.
__device__ __forceinline__ void deliver_record(const Cols& C, const int buf, const Rec& r)
{
    const u32 row = R_row(r), from = R_from(r), k = (u32)(r.w0.y >> 32);
    if (row >= C.rows || from >= C.members || k >= RA_MBOX_DEPTH) return;
    const u32 tail = st_rec_plane(C.mbox[buf], C.tiles, from * RA_MBOX_DEPTH + k, row, r) ? 8u : 0u;
    u32* w = reinterpret_cast<u32*>(&C.mbox_cnt[buf][row]) + (from >> 2);
    const u32 sh = 8u * (from & 3u);
    u32 old = *w;
    for (;;) {
        const u32 ob = (old >> sh) & 0xffu;
        const u32 nb = ((ob & 7u) >= k + 1 ? (ob & 7u) : k + 1) | (ob & 8u) | tail;
        if (nb == ob) break;
        const u32 seen = atomicCAS(w, old, (old & ~(0xffu << sh)) | (nb << sh));
        if (seen == old) break;
        old = seen;
    }
}

// ---- consistent-query state of a row <-> ra_query_state ----------------------------------------
__device__ __forceinline__ void load_query_row(const Cols& C, const ra_query_state& q)
{
    const u32 r = q.row;
    *C.q_used = 1u;
    C.qi[r] = q.query_index; C.qa[r] = q.agreed_index;
    for (u32 p = 0; p < C.members; p++) C.pqi[(size_t)p * C.rows + r] = q.peer_query_index[p];
}
__device__ __forceinline__ void read_query_row(const Cols& C, ra_query_state& q)
{
    const u32 r = q.row;
    q._pad = 0; q.query_index = C.qi[r]; q.agreed_index = C.qa[r];
    for (u32 p = 0; p < RA_MAX_MEMBERS; p++) q.peer_query_index[p] = p < C.members ? C.pqi[(size_t)p * C.rows + r] : 0;
}
