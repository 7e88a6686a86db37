/* TEST INFRASTRUCTURE -- not part of the product, never loaded by ra_b200/.
 *
 * Host emulation of the engine's two step kernels: the DEVICE LOGIC headers of the product
 * (ra_b200/csrc/raft_step.cuh, raft_row.cuh -- fast_event, process_event, emit_msg, the record
 * head/tail codec, row_end_of_step, load_row / read_row ...) are compiled here as plain C++
 * through tests/emu/cuda_shim.h and driven one row at a time with the control flow of
 * raft_step_kernel (fast paths until a row stalls) and raft_general_kernel (resume from the
 * 64-byte stall context).  What is NOT covered is the kernels' own frame in engine.cu (TMA ring,
 * warp reductions, launch plumbing): that is what the `-m gpu` tests are for.  The point of this
 * library is that the CPU test tier diffs the very source the GPU runs against the oracle
 * (tests/test_emu_parity.py), so a logic regression shows up before any GPU time is spent.
 *
 * Same C ABI as include/ra_engine.h with the prefix ra_emu_.  Shards: the bucket transport
 * (ra_emu_set_outbox / ra_emu_deliver); the peer transport is the same stores through other
 * base pointers and is left to the GPU tests.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include "cuda_shim.h"
#include "../../ra_b200/csrc/raft_step.cuh"
#include "../../ra_b200/csrc/raft_row.cuh"

struct ra_emu {
    ra_engine_cfg cfg;
    Cols C;
    int cur;
    u64 step_no, steps;
    void* allocs[96]; int n_allocs;
    int narrow, narrow_mode;                          // as engine.cu: RA_STEP_WIDE unset = auto, 0 = always narrow, 1 = never
    int out_pending;                                  // a step's outputs did not fit: they wait in the row slots
    int sub_busy, sub_rc; size_t sub_nm, sub_nn;      // the one "submitted" call (ra_engine_submit_host shim)
};

template <typename T>
static int halloc(ra_emu* e, T** p, size_t count)
{
    void* q = calloc(count ? count : 1, sizeof(T));
    if (!q) return RA_E_NOMEM;
    e->allocs[e->n_allocs++] = q;
    *p = (T*)q;
    return RA_OK;
}

extern "C" void ra_emu_destroy(ra_emu* e)
{
    if (!e) return;
    for (int i = 0; i < e->n_allocs; i++) free(e->allocs[i]);
    free(e);
}

extern "C" int ra_emu_reset_empty(ra_emu* e)
{
    if (!e) return RA_E_INVAL;
    for (u32 r = 0; r < e->C.rows; r++) reset_row(e->C, r);
    memset(e->C.counters, 0, (8 + 8 * 16 + 8) * sizeof(u64));
    e->C.q_used[0] = 0;
    e->cur = 0; e->step_no = 0; e->steps = 0;
    e->narrow = e->narrow_mode != 2;
    return RA_OK;
}

extern "C" int ra_emu_create(const ra_engine_cfg* cfg, ra_emu** out)
{
    if (!cfg || !out || cfg->n_members < 1 || cfg->n_members > RA_MAX_MEMBERS || cfg->n_groups == 0) return RA_E_INVAL;
    if ((u64)cfg->n_groups * cfg->n_members > 0x7fffffffull) return RA_E_INVAL;
    if (cfg->n_shards > 1 && (!cfg->route_on_device || cfg->shard >= cfg->n_shards || cfg->n_shards > 64)) return RA_E_INVAL;
    ra_emu* e = (ra_emu*)calloc(1, sizeof(ra_emu));
    if (!e) return RA_E_NOMEM;
    e->cfg = *cfg;
    { const char* w = getenv("RA_STEP_WIDE"); e->narrow_mode = !(w && *w) ? 0 : (*w == '0' ? 1 : 2); e->narrow = e->narrow_mode != 2; }
    if (e->cfg.max_pipeline_count == 0) e->cfg.max_pipeline_count = 4096;
    if (e->cfg.max_aer_batch == 0) e->cfg.max_aer_batch = 128;
    int rc = RA_OK;
    Cols& C = e->C;
    const size_t R = (size_t)cfg->n_groups * cfg->n_members, M = cfg->n_members;
    C.rows = (u32)R; C.groups = cfg->n_groups; C.members = cfg->n_members;
    C.groups_inv = cfg->n_groups > 1 ? (u32)(0x100000000ull / cfg->n_groups) : 0xFFFFFFFFu;
    C.max_pipeline = e->cfg.max_pipeline_count; C.max_batch = e->cfg.max_aer_batch;
    C.routed = cfg->route_on_device ? 1 : 0; C.pure = cfg->pure ? 1 : 0;
    C.note_cap = cfg->note_cap ? cfg->note_cap : RA_NOTE_CAP;
    C.n_shards = cfg->n_shards > 1 ? cfg->n_shards : 1; C.shard = cfg->n_shards > 1 ? cfg->shard : 0;
    C.outbox = nullptr; C.out_cnt = nullptr; C.out_cap = 0;
    C.peer_mode = 0;
    for (int b = 0; b < 2; b++) for (int k = 0; k < 8; k++) { C.peer_mbox[b][k] = nullptr; C.peer_cnt[b][k] = nullptr; }
#define HA(p, n) if ((rc = halloc(e, &(p), (n))) != RA_OK) goto bad
    HA(C.tc, R); HA(C.lg, R); HA(C.lw, R); HA(C.ap, R); HA(C.sn, R); HA(C.tk, R); HA(C.fm, R);
    HA(C.cd, 2 * R); HA(C.pnm, M * R); HA(C.pcs, M * R); HA(C.run, RA_MAX_RUNS * R); HA(C.lrs, R);
    HA(C.qi, R); HA(C.qa, R); HA(C.pqi, M * R);
    C.tiles = (u32)((R + RT - 1) / RT);
    {
        const size_t PW = (size_t)C.tiles * 4 * RT;
        HA(C.loc, (size_t)RA_LOCAL_CAP * PW); HA(C.loc_n, R);
        HA(C.onote, (size_t)RA_NOTE_CAP * R); HA(C.out_n, R); HA(C.counters, 8 + 8 * 16 + 8); HA(C.q_used, 4); HA(C.wc, R); HA(C.wf, R);
        if (C.routed) {
            for (int b = 0; b < 2; b++) { HA(C.mbox[b], M * RA_MBOX_DEPTH * PW); HA(C.mbox_cnt[b], R); }
            HA(C.omsg, (size_t)RA_MSG_CAP * (C.pure ? R : 1));
        } else {
            C.mbox[0] = C.mbox[1] = nullptr; C.mbox_cnt[0] = C.mbox_cnt[1] = nullptr;
            HA(C.omsg, (size_t)RA_MSG_CAP * R);
        }
    }
#undef HA
    if ((rc = ra_emu_reset_empty(e)) != RA_OK) goto bad;
    *out = e;
    return RA_OK;
bad:
    ra_emu_destroy(e);
    return rc;
}

extern "C" int ra_emu_get_cfg(ra_emu* e, ra_engine_cfg* out)
{
    if (!e || !out) return RA_E_INVAL;
    *out = e->cfg;
    return RA_OK;
}

extern "C" int ra_emu_load_rows(ra_emu* e, const ra_row_state* rows, size_t n)
{
    if (!e || (!rows && n)) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++)
        if (rows[i].row >= e->C.rows || rows[i].n_members != e->C.members || !ra_row_state_valid(&rows[i])) return RA_E_INVAL;
    if (e->narrow_mode == 0 && n * 2 >= e->C.rows) {
        size_t wide = 0;
        for (size_t i = 0; i < n; i++) wide += row_state_is_wide(rows[i], e->C.members) ? 1 : 0;
        e->narrow = wide * 2 < n;
    }
    for (size_t i = 0; i < n; i++) load_row(e->C, rows[i]);
    return RA_OK;
}

extern "C" int ra_emu_read_rows(ra_emu* e, ra_row_state* rows, size_t n)
{
    if (!e || (!rows && n)) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++) if (rows[i].row >= e->C.rows) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++) read_row(e->C, rows[i]);
    return RA_OK;
}

extern "C" int ra_emu_load_query_state(ra_emu* e, const ra_query_state* q, size_t n)
{
    if (!e || (!q && n)) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++) if (q[i].row >= e->C.rows) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++) load_query_row(e->C, q[i]);
    return RA_OK;
}

extern "C" int ra_emu_read_query_state(ra_emu* e, ra_query_state* q, size_t n)
{
    if (!e || (!q && n)) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++) if (q[i].row >= e->C.rows) return RA_E_INVAL;
    for (size_t i = 0; i < n; i++) read_query_row(e->C, q[i]);
    return RA_OK;
}

// ---- one row through one step: raft_step_kernel, then raft_general_kernel if it stalled -------

struct Scratch { ulonglong2 nm[RA_MAX_MEMBERS]; u64 cs[RA_MAX_MEMBERS]; };   // the per-thread shared-memory columns

// test observability: rows stepped by the narrow pass, records it refused (a field >= 2^30), rows it left to the
// general kernel because their sticky `wide` byte is set
static unsigned long long g_narrow_stats[3];
extern "C" void ra_emu_narrow_stats(unsigned long long* out, int reset)
{ for (int i = 0; i < 3; i++) { out[i] = g_narrow_stats[i]; if (reset) g_narrow_stats[i] = 0; } }

namespace ra_wide {
#define RA_NARROW_PASS 0
#include "step_row.inc"
#undef RA_NARROW_PASS
}
namespace ra_narrow {
#define RA_NARROW_PASS 1
#include "step_row.inc"
#undef RA_NARROW_PASS
}
// a value left or entered the 32-bit pass out of range: the guards (Cols::wf, rec_decode) have a hole
extern "C" void ra_emu_narrow_violation(const char* what, unsigned long long v)
{
    fprintf(stderr, "ra_emu: narrow pass violation in %s(): value %llu\n", what, v);
    abort();
}

// the general kernel's body for one stall context (engine.cu: raft_general_kernel)
static void general_row(const Cols& C, int cur, const FloodArgs& F, const StallCtx& ctx)
{
    constexpr int MM = MK_MM(0, TR_RUNTIME);
    Scratch sc;
    const u32 r = ctx.row, flags = ctx.flags;
    u32 rem_mbox = ctx.rem_mbox, rem_loc = ctx.rem_loc;
    Member m;
    member_init(m, C, r, C.tc[r], C.lg[r], C.lw[r], C.ap[r], 0, cur, sc.nm);
    m.lrs_ok = 0;
    m.n_msgs = ctx.n_msgs_notes & 0xffffu; m.n_notes = (ctx.n_msgs_notes >> 16) & 0xffffu; m.status = ctx.status;
    m.sent_to = ctx.sent_to;
    { const u32 w = ctx.pn_type_slot_wk; m.pn_type = w & 0xffu; m.pn_slot = (w >> 8) & 0xffu; m.wk = w >> 16; }
    m.pn_a = ctx.pn_a; m.pn_b = ctx.pn_b; m.pn_c = ctx.pn_c;
    u64 big = 0;
    if (flags & STALL_PENDING) {
        MT_SET(m.meta, 24, 1, 0);
        process_event<MM>(m, mk_rec(r, RA_EV_PIPELINE_RPCS, RA_NO_SLOT, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
        m.cold &= ~8u;
    }
    while (rem_mbox) {
        const u32 p = __ffs(rem_mbox) - 1; rem_mbox &= rem_mbox - 1;
        const Rec e = ld_rec_plane(C.mbox[cur], C.tiles, p, r);
        big |= rec_magnitude(e);
        if (MT_FATAL(m.meta)) m.c_pack += 1u;
        else if (flood_lost<MM>(F, C, m, r, e)) m.c_pack += 1u << 20;
        else if (!note_budget_ok(m)) budget_drop_record(m);
        else if (C.pure || !fast_event<MM>(m, e)) { process_event<MM>(m, e); m.cold &= ~8u; }
    }
    while (rem_loc) {
        const u32 p = __ffs(rem_loc) - 1; rem_loc &= rem_loc - 1;
        const Rec e = ld_rec_plane(C.loc, C.tiles, p, r);
        big |= rec_magnitude(e);
        if (MT_FATAL(m.meta)) m.c_pack += 1u;
        else if (!note_budget_ok(m)) budget_refuse_local(m);
        else if (C.pure || !fast_event<MM>(m, e)) { process_event<MM>(m, e); m.cold &= ~8u; }
    }
    peers_writeback<MM>(m);
    const u32 k_fatal = row_end_of_step<MM>(m, C, r, cur, F);
    member_writeback(m, C, r);
    row_mark_wide(C, r, m, big);
    ra_wide::add_counters(C, m, k_fatal);
}

static int run_step(ra_emu* e, const FloodArgs& F)
{
    const Cols& C = e->C;
    if (C.n_shards > 1) {
        if (!C.outbox) return RA_E_INVAL;                       // ra_emu_set_outbox first
        memset(C.out_cnt, 0, C.n_shards * sizeof(u32));
    }
    // same choice of specialisation as launch_step() in engine.cu
    const int tr = !C.routed ? TR_HOST : (C.n_shards > 1 ? TR_BUCKET : TR_LOCAL);
    for (u32 r = 0; r < C.rows; r++) {
        StallCtx ctx;
        bool stalled;
        if (C.members == 5) {
            switch (tr) {
            // (narrow pass for the same specialisations as launch_step(); RA_STEP_WIDE=1 keeps the 64-bit pass)
            case TR_LOCAL:  stalled = e->narrow ? ra_narrow::step_row<MK_MM(5, TR_LOCAL)>(C, e->cur, F, r, ctx)
                                                : ra_wide::step_row<MK_MM(5, TR_LOCAL)>(C, e->cur, F, r, ctx); break;
            case TR_BUCKET: stalled = ra_wide::step_row<MK_MM(5, TR_BUCKET)>(C, e->cur, F, r, ctx); break;
            default:        stalled = e->narrow ? ra_narrow::step_row<MK_MM(5, TR_HOST)>(C, e->cur, F, r, ctx)
                                                : ra_wide::step_row<MK_MM(5, TR_HOST)>(C, e->cur, F, r, ctx); break;
            }
        } else stalled = ra_wide::step_row<MK_MM(0, TR_RUNTIME)>(C, e->cur, F, r, ctx);
        // the general kernel runs after the step kernel; rows only ever write to OTHER rows' mailboxes of
        // the NEXT step, so handling a stalled row right away is the same thing
        if (stalled) general_row(C, e->cur, F, ctx);
    }
    if (C.routed) e->cur ^= 1;
    e->steps++;
    return RA_OK;
}

extern "C" int ra_emu_set_outbox(ra_emu* e, void* outbox, uint32_t* counts, uint32_t cap)
{
    if (!e || e->C.n_shards < 2 || !outbox || !counts || !cap) return RA_E_INVAL;
    e->C.outbox = (ra_event*)outbox; e->C.out_cnt = counts; e->C.out_cap = cap;
    return RA_OK;
}

extern "C" int ra_emu_deliver(ra_emu* e, const void* inbox, const uint32_t* counts, uint32_t cap)
{
    if (!e || e->C.n_shards < 2 || !inbox || !counts || !cap) return RA_E_INVAL;
    const ra_event* in = (const ra_event*)inbox;
    // e->cur is the buffer the next step reads: the one the last step's senders wrote into
    for (u32 b = 0; b < e->C.n_shards; b++) {
        const u32 n = counts[b] < cap ? counts[b] : cap;
        for (u32 i = 0; i < n; i++) deliver_record(e->C, e->cur, ld_rec(&in[(size_t)b * cap + i]));
    }
    return RA_OK;
}

// scan + gather_out: per-row slots -> flat arrays ordered by (row, seq).  When they do not fit, nothing is
// written and the slots stay: RA_E_CAPACITY with the sizes needed, the caller fetches again (engine.cu)
extern "C" int ra_emu_fetch_output(ra_emu* e, ra_event* msgs, size_t msgs_cap, size_t* n_msgs,
                                   ra_note* notes, size_t notes_cap, size_t* n_notes)
{
    if (!e) return RA_E_INVAL;
    const Cols& C = e->C;
    const u32 R = C.rows;
    size_t tm = 0, tn = 0;
    for (u32 r = 0; r < R; r++) { tm += C.out_n[r] & 0xffffu; tn += C.out_n[r] >> 16; }
    if (n_msgs) *n_msgs = tm;
    if (n_notes) *n_notes = tn;
    if (tm > msgs_cap || tn > notes_cap) { e->out_pending = 1; return RA_E_CAPACITY; }
    e->out_pending = 0;
    size_t om = 0, on = 0;
    for (u32 r = 0; r < R; r++) {
        const u32 v = C.out_n[r];
        C.out_n[r] = 0;
        const u32 nm = v & 0xffffu, nn = v >> 16;
        for (u32 k = 0; k < nm; k++) st_rec(&msgs[om + k], ld_rec(&C.omsg[(size_t)k * C.rows + r]));
        for (u32 k = 0; k < nn; k++) notes[on + k] = C.onote[(size_t)k * C.rows + r];
        om += nm; on += nn;
    }
    return RA_OK;
}

extern "C" int ra_emu_step(ra_emu* e, const ra_event* ev, size_t n_ev,
                           ra_event* msgs, size_t msgs_cap, size_t* n_msgs,
                           ra_note* notes, size_t notes_cap, size_t* n_notes)
{
    if (!e || (!ev && n_ev) || n_ev > 0x7fffffffull) return RA_E_INVAL;
    if (e->out_pending) return RA_E_CAPACITY;                    // ra_emu_fetch_output first
    const Cols& C = e->C;
    const u32 R = C.rows;
    for (u32 r = 0; r < R; r++) C.loc_n[r] = 0;                  // clear_loc_kernel
    // ingest_kernel: flat batch -> per-row local slots
    u32 err = 0;
    for (size_t i = 0; i < n_ev; i++) {
        const u32 row = ev[i].row;
        if (row >= R || ev[i].type > RA_EV_CONSISTENT_QUERY || ev[i].type == RA_EV_NONE) { if (err < 3) err = 3; continue; }
        if (i > 0 && ev[i - 1].row == row) continue;
        u32 len = 1;
        while (i + len < n_ev && ev[i + len].row == row && len <= RA_LOCAL_CAP) len++;
        if (len > RA_LOCAL_CAP) { if (err < 2) err = 2; continue; }
        if (C.loc_n[row] != 0) { if (err < 1) err = 1; continue; }
        C.loc_n[row] = len;
        u32 tails = 0;
        for (u32 k = 0; k < len; k++)
            if (st_rec_plane(C.loc, C.tiles, k, row, ld_rec(&ev[i + k]))) tails |= 0x100u << k;
        C.loc_n[row] |= tails;
    }
    if (err) {
        for (u32 r = 0; r < R; r++) C.loc_n[r] = 0;
        return err == 1 ? RA_E_UNGROUPED : (err == 2 ? RA_E_CAPACITY : RA_E_INVAL);
    }
    FloodArgs F; memset(&F, 0, sizeof F);
    { const int rc = run_step(e, F); if (rc) return rc; }
    return ra_emu_fetch_output(e, msgs, msgs_cap, n_msgs, notes, notes_cap, n_notes);
}

extern "C" int ra_emu_step_host(ra_emu* e, const ra_host_event* ev, size_t n_ev,
                                ra_event* msgs, size_t msgs_cap, size_t* n_msgs,
                                ra_note* notes, size_t notes_cap, size_t* n_notes)
{
    if (!e || (!ev && n_ev)) return RA_E_INVAL;
    ra_event* w = (ra_event*)calloc(n_ev ? n_ev : 1, sizeof(ra_event));
    if (!w) return RA_E_NOMEM;
    for (size_t i = 0; i < n_ev; i++) {
        const u32 t = ev[i].type;
        if (!(t == RA_EV_WRITTEN || t == RA_EV_COMMAND || t == RA_EV_ELECTION_TIMEOUT || t == RA_EV_AWAIT_COND_TIMEOUT ||
              t == RA_EV_PIPELINE_RPCS || t == RA_EV_TICK || t == RA_EV_CONSISTENT_QUERY)) { free(w); return RA_E_INVAL; }
        w[i].row = ev[i].row; w[i].type = (uint8_t)t; w[i].from_slot = RA_NO_SLOT; w[i].flags = ev[i].flags; w[i].n = ev[i].n;
        w[i].term = ev[i].term; w[i].a = ev[i].a; w[i].b = ev[i].b;
    }
    const int rc = ra_emu_step(e, w, n_ev, msgs, msgs_cap, n_msgs, notes, notes_cap, n_notes);
    free(w);
    return rc;
}

extern "C" int ra_emu_flood_faults(ra_emu* e, uint32_t n_steps, uint32_t cmds_per_step, uint32_t election_permille,
                                   uint64_t seed, const ra_flood_faults* ff);
extern "C" int ra_emu_flood(ra_emu* e, uint32_t n_steps, uint32_t cmds_per_step, uint32_t election_permille,
                            uint64_t seed)
{ return ra_emu_flood_faults(e, n_steps, cmds_per_step, election_permille, seed, nullptr); }
extern "C" int ra_emu_flood_faults(ra_emu* e, uint32_t n_steps, uint32_t cmds_per_step, uint32_t election_permille,
                                   uint64_t seed, const ra_flood_faults* ff)
{
    if (!e || !e->C.routed) return RA_E_INVAL;
    if (ff && ff->partition_permille && !ff->partition_steps) return RA_E_INVAL;
    for (u32 t = 0; t < n_steps; t++) {
        FloodArgs F; memset(&F, 0, sizeof F);
        F.on = 1; F.cmds = cmds_per_step; F.permille = election_permille;
        if (ff) { F.drop = ff->drop_permille; F.withhold = ff->withhold_permille; F.part = ff->partition_permille; F.part_len = ff->partition_steps; }
        F.seed = seed; F.step = e->step_no + t;
        const int rc = run_step(e, F);
        if (rc) return rc;
    }
    e->step_no += n_steps;
    return RA_OK;
}

extern "C" int ra_emu_counters(ra_emu* e, ra_counters* out)
{
    if (!e || !out) return RA_E_INVAL;
    const u64* h = e->C.counters;
    out->events = h[0]; out->commits = h[1]; out->applied = h[2]; out->msgs_out = h[3];
    out->msgs_dropped = h[4]; out->elections_won = h[5]; out->fatal_rows = h[6]; out->steps = e->steps;
    out->aer_received_follower = h[136]; out->aer_received_follower_empty = h[137]; out->aer_replies_success = h[138];
    out->aer_replies_failed = h[139]; out->elections = h[140]; out->pre_vote_elections = h[141];
    out->term_and_voted_for_updates = h[142];
    return RA_OK;
}

extern "C" int ra_emu_stall_histogram(ra_emu* e, uint64_t* out128)
{
    if (!e || !out128) return RA_E_INVAL;
    memcpy(out128, e->C.counters + 8, 128 * sizeof(u64));
    return RA_OK;
}

/* ---- ra_hostsim (ra_b200/csrc/host_flood.cu, the host-side caller of the ABI used by the e2e
 * benchmark) over this emulation: the file is compiled into this library with RA_NO_CUDA and
 * reaches the "engine" through these three forwarding entry points. */
extern "C" int ra_engine_step(ra_engine* e, const ra_event* ev, size_t n_ev, ra_event* msgs, size_t msgs_cap,
                              size_t* n_msgs, ra_note* notes, size_t notes_cap, size_t* n_notes)
{ return ra_emu_step((ra_emu*)e, ev, n_ev, msgs, msgs_cap, n_msgs, notes, notes_cap, n_notes); }
extern "C" int ra_engine_get_cfg(ra_engine* e, ra_engine_cfg* out) { return ra_emu_get_cfg((ra_emu*)e, out); }

/* ---- the record-plane codec on its own: one ABI record through st_rec_plane / ld_rec_plane ---------- */
extern "C" int ra_emu_codec_roundtrip(const ra_event* in, ra_event* out, int* had_tail)
{
    if (!in || !out) return RA_E_INVAL;
    static thread_local ulonglong2 plane[4 * RT];              // one tile of one plane
    memset(plane, 0xA5, sizeof plane);                         // stale bytes must not leak into the result
    const u32 row = in->row & (RT - 1);
    const bool tail = st_rec_plane(plane, 1, 0, row, ld_rec(in));
    if (had_tail) *had_tail = tail ? 1 : 0;
    st_rec(out, ld_rec_plane(plane, 1, 0, row));
    return RA_OK;
}
extern "C" int ra_engine_step_host(ra_engine* e, const ra_host_event* ev, size_t n_ev, ra_event* msgs, size_t msgs_cap,
                                   size_t* n_msgs, ra_note* notes, size_t notes_cap, size_t* n_notes)
{ return ra_emu_step_host((ra_emu*)e, ev, n_ev, msgs, msgs_cap, n_msgs, notes, notes_cap, n_notes); }
/* the split-phase pair: the emulation evaluates at submit and hands the result over at collect */
extern "C" int ra_engine_submit_host(ra_engine* e, const ra_host_event* ev, size_t n_ev, ra_event* msgs, size_t msgs_cap,
                                     ra_note* notes, size_t notes_cap)
{
    ra_emu* m = (ra_emu*)e;
    if (!m || m->sub_busy) return RA_E_BUSY;
    m->sub_rc = ra_emu_step_host(m, ev, n_ev, msgs, msgs_cap, &m->sub_nm, notes, notes_cap, &m->sub_nn);
    m->sub_busy = 1;
    return RA_OK;
}
extern "C" int ra_engine_submit_host_segs(ra_engine* e, const ra_host_event_seg* segs, size_t n_segs, ra_event* msgs,
                                          size_t msgs_cap, ra_note* notes, size_t notes_cap)
{
    size_t n = 0;
    for (size_t k = 0; k < n_segs; k++) n += segs[k].n;
    ra_host_event* all = (ra_host_event*)malloc((n ? n : 1) * sizeof(ra_host_event));
    if (!all) return RA_E_NOMEM;
    size_t off = 0;
    for (size_t k = 0; k < n_segs; k++) { if (segs[k].n) memcpy(all + off, segs[k].ev, segs[k].n * sizeof(ra_host_event)); off += segs[k].n; }
    const int rc = ra_engine_submit_host(e, all, n, msgs, msgs_cap, notes, notes_cap);
    free(all);
    return rc;
}
extern "C" int ra_engine_collect(ra_engine* e, size_t* n_msgs, size_t* n_notes)
{
    ra_emu* m = (ra_emu*)e;
    if (!m || !m->sub_busy) return RA_E_INVAL;
    m->sub_busy = 0;
    if (n_msgs) *n_msgs = m->sub_nm;
    if (n_notes) *n_notes = m->sub_nn;
    return m->sub_rc;
}
