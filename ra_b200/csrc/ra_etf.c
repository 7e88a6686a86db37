/* ra_etf.c -- AppendEntries wire codec: Erlang external term format <-> 64-byte engine records
 * (include/ra_etf.h).  Host-only C.
 *
 * The format is the documented one ("External Term Format", erts): a message is 131 then one term; integers
 * are SMALL_INTEGER_EXT (97, 1 byte), INTEGER_EXT (98, 4 bytes big endian, signed) or SMALL_BIG_EXT (110, n,
 * sign, n little-endian bytes); atoms are [SMALL_]ATOM[_UTF8]_EXT; a record is a tuple whose first element
 * is the record name.  Records mirrored: rabbitmq/ra src/ra.hrl:122-141.
 */
#include <string.h>
#include "../../include/ra_etf.h"

enum {
    T_NEW_FLOAT = 70, T_BIT_BINARY = 77, T_NEW_PID = 88, T_NEW_PORT = 89, T_NEWER_REF = 90,
    T_SMALL_INT = 97, T_INT = 98, T_FLOAT = 99, T_ATOM = 100, T_REF = 101, T_PORT = 102, T_PID = 103,
    T_SMALL_TUPLE = 104, T_LARGE_TUPLE = 105, T_NIL = 106, T_STRING = 107, T_LIST = 108, T_BINARY = 109,
    T_SMALL_BIG = 110, T_LARGE_BIG = 111, T_EXPORT = 113, T_NEW_REF = 114, T_SMALL_ATOM = 115, T_MAP = 116,
    T_ATOM_UTF8 = 118, T_SMALL_ATOM_UTF8 = 119, T_V4_PORT = 120
};

typedef struct { const uint8_t* p; size_t len, off; int err; } rd_t;

static int need(rd_t* r, size_t n)
{
    if (r->err) return 0;
    if (r->len - r->off < n) { r->err = RA_ETF_E_TRUNCATED; return 0; }
    return 1;
}
static uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* size of the term at p[0..len), 0 on error */
static size_t term_size(const uint8_t* p, size_t len, int depth)
{
    if (len < 1 || depth > 64) return 0;
    size_t n, off, i;
    switch (p[0]) {
    case T_SMALL_INT: return len >= 2 ? 2 : 0;
    case T_INT: return len >= 5 ? 5 : 0;
    case T_FLOAT: return len >= 32 ? 32 : 0;
    case T_NEW_FLOAT: return len >= 9 ? 9 : 0;
    case T_ATOM: case T_ATOM_UTF8: if (len < 3) return 0; n = 3 + be16(p + 1); return n <= len ? n : 0;
    case T_SMALL_ATOM: case T_SMALL_ATOM_UTF8: if (len < 2) return 0; n = 2u + p[1]; return n <= len ? n : 0;
    case T_NIL: return 1;
    case T_STRING: if (len < 3) return 0; n = 3 + be16(p + 1); return n <= len ? n : 0;
    case T_BINARY: if (len < 5) return 0; n = 5 + (size_t)be32(p + 1); return n <= len ? n : 0;
    case T_BIT_BINARY: if (len < 6) return 0; n = 6 + (size_t)be32(p + 1); return n <= len ? n : 0;
    case T_SMALL_BIG: if (len < 3) return 0; n = 3u + p[1]; return n <= len ? n : 0;
    case T_LARGE_BIG: if (len < 6) return 0; n = 6 + (size_t)be32(p + 1); return n <= len ? n : 0;
    case T_SMALL_TUPLE: case T_LARGE_TUPLE: case T_LIST: case T_MAP: {
        size_t cnt;
        if (p[0] == T_SMALL_TUPLE) { if (len < 2) return 0; cnt = p[1]; off = 2; }
        else { if (len < 5) return 0; cnt = be32(p + 1); off = 5; }
        if (p[0] == T_MAP) cnt *= 2;
        if (p[0] == T_LIST) cnt += 1;                               /* the tail */
        for (i = 0; i < cnt; i++) {
            n = term_size(p + off, len - off, depth + 1);
            if (!n) return 0;
            off += n;
        }
        return off;
    }
    /* node atom, then fixed fields */
    case T_PID: case T_NEW_PID: case T_PORT: case T_NEW_PORT: case T_V4_PORT: case T_REF: case T_NEW_REF: case T_NEWER_REF: {
        off = 1;
        size_t idn = 0;
        if (p[0] == T_NEW_REF || p[0] == T_NEWER_REF) { if (len < 3) return 0; idn = be16(p + 1); off = 3; }
        n = term_size(p + off, len - off, depth + 1);               /* the node name */
        if (!n) return 0;
        off += n;
        switch (p[0]) {
        case T_PID: off += 9; break;                                /* id 4, serial 4, creation 1 */
        case T_NEW_PID: off += 12; break;                           /* id 4, serial 4, creation 4 */
        case T_PORT: off += 5; break;
        case T_NEW_PORT: off += 8; break;
        case T_V4_PORT: off += 12; break;
        case T_REF: off += 5; break;
        case T_NEW_REF: off += 1 + 4 * idn; break;
        default: off += 4 + 4 * idn; break;                         /* NEWER_REFERENCE_EXT */
        }
        return off <= len ? off : 0;
    }
    case T_EXPORT: {
        off = 1;
        for (i = 0; i < 3; i++) { n = term_size(p + off, len - off, depth + 1); if (!n) return 0; off += n; }
        return off;
    }
    default: return 0;
    }
}

size_t ra_etf_term_size(const uint8_t* p, size_t len) { return p ? term_size(p, len, 0) : 0; }

/* ---- readers ---------------------------------------------------------------------------------------- */
static uint64_t rd_uint(rd_t* r)
{
    if (!need(r, 1)) return 0;
    const uint8_t* p = r->p + r->off;
    if (p[0] == T_SMALL_INT) { if (!need(r, 2)) return 0; r->off += 2; return p[1]; }
    if (p[0] == T_INT) {
        if (!need(r, 5)) return 0;
        uint32_t v = be32(p + 1);
        if (v & 0x80000000u) { r->err = RA_ETF_E_RANGE; return 0; }
        r->off += 5; return v;
    }
    if (p[0] == T_SMALL_BIG) {
        if (!need(r, 3)) return 0;
        uint32_t n = p[1];
        if (!need(r, 3 + n)) return 0;
        if (p[2] != 0) { r->err = RA_ETF_E_RANGE; return 0; }       /* negative */
        uint64_t v = 0;
        for (uint32_t i = 0; i < n; i++) {
            if (i >= 8) { if (p[3 + i]) { r->err = RA_ETF_E_RANGE; return 0; } continue; }
            v |= (uint64_t)p[3 + i] << (8 * i);
        }
        r->off += 3 + n; return v;
    }
    r->err = RA_ETF_E_FORMAT;
    return 0;
}

/* atom -> NUL-terminated text (<= 255 bytes), returns length or -1 */
static int rd_atom(rd_t* r, char* out)
{
    if (!need(r, 1)) return -1;
    const uint8_t* p = r->p + r->off;
    size_t n, hdr;
    if (p[0] == T_ATOM || p[0] == T_ATOM_UTF8) { if (!need(r, 3)) return -1; n = be16(p + 1); hdr = 3; }
    else if (p[0] == T_SMALL_ATOM || p[0] == T_SMALL_ATOM_UTF8) { if (!need(r, 2)) return -1; n = p[1]; hdr = 2; }
    else { r->err = RA_ETF_E_FORMAT; return -1; }
    if (!need(r, hdr + n)) return -1;
    if (n > 255) { r->err = RA_ETF_E_FORMAT; return -1; }
    if (out) { memcpy(out, p + hdr, n); out[n] = 0; }
    r->off += hdr + n;
    return (int)n;
}
static int rd_atom_is(rd_t* r, const char* want)
{
    char buf[256];
    if (rd_atom(r, buf) < 0) return 0;
    if (strcmp(buf, want) != 0) { r->err = RA_ETF_E_FORMAT; return 0; }
    return 1;
}
static uint32_t rd_tuple(rd_t* r)
{
    if (!need(r, 2)) return 0;
    const uint8_t* p = r->p + r->off;
    if (p[0] == T_SMALL_TUPLE) { r->off += 2; return p[1]; }
    if (p[0] == T_LARGE_TUPLE) { if (!need(r, 5)) return 0; r->off += 5; return be32(p + 1); }
    r->err = RA_ETF_E_FORMAT;
    return 0;
}
static void rd_id(rd_t* r, ra_etf_id* id)
{
    if (rd_tuple(r) != 2) { if (!r->err) r->err = RA_ETF_E_FORMAT; return; }
    rd_atom(r, id ? id->name : NULL);
    rd_atom(r, id ? id->node : NULL);
}

int ra_etf_decode_aer(const uint8_t* msg, size_t len, ra_event* ev, ra_etf_id* leader,
                      ra_etf_entry* entries, size_t max_entries, size_t* n_entries)
{
    if (!msg || !ev || len < 2 || msg[0] != 131) return RA_ETF_E_FORMAT;
    rd_t r = { msg, len, 1, 0 };
    if (rd_tuple(&r) != 7 || !rd_atom_is(&r, "append_entries_rpc")) return r.err ? r.err : RA_ETF_E_FORMAT;
    memset(ev, 0, sizeof *ev);
    ev->type = RA_EV_AER;
    ev->from_slot = RA_NO_SLOT;                                     /* the caller maps `leader` to a slot */
    ev->term = rd_uint(&r);
    rd_id(&r, leader);
    ev->c = rd_uint(&r);                                            /* leader_commit  */
    ev->a = rd_uint(&r);                                            /* prev_log_index */
    ev->b = rd_uint(&r);                                            /* prev_log_term  */
    if (r.err) return r.err;
    size_t n = 0;
    if (!need(&r, 1)) return r.err;
    if (msg[r.off] == T_NIL) r.off += 1;
    else if (msg[r.off] == T_LIST) {
        if (!need(&r, 5)) return r.err;
        const size_t cnt = be32(msg + r.off + 1);
        r.off += 5;
        uint64_t t0 = 0, t1 = 0; uint32_t n1 = 0; int runs = 0;
        for (size_t i = 0; i < cnt; i++) {
            if (rd_tuple(&r) != 3) return r.err ? r.err : RA_ETF_E_FORMAT;
            const uint64_t idx = rd_uint(&r), term = rd_uint(&r);
            if (r.err) return r.err;
            const size_t cs = term_size(msg + r.off, len - r.off, 0);
            if (!cs) return RA_ETF_E_TRUNCATED;
            if (idx != ev->a + 1 + i) return RA_ETF_E_RUNS;         /* entries are prev+1 .. prev+n */
            if (runs == 0) { t0 = term; runs = 1; }
            else if (runs == 1 && term != t0) { t1 = term; n1 = (uint32_t)i; runs = 2; }
            else if ((runs == 1 && term != t0) || (runs == 2 && term != t1)) return RA_ETF_E_RUNS;
            if (i >= max_entries || i >= 0xFFFF) return RA_ETF_E_CAPACITY;
            if (entries) { entries[i].index = idx; entries[i].term = term; entries[i].cmd_off = (uint32_t)r.off; entries[i].cmd_len = (uint32_t)cs; }
            r.off += cs;
        }
        if (!need(&r, 1) || msg[r.off] != T_NIL) return r.err ? r.err : RA_ETF_E_FORMAT;   /* proper list */
        r.off += 1;
        n = cnt;
        ev->n = (uint16_t)cnt; ev->n1 = (uint16_t)n1; ev->d = t0; ev->e = runs == 2 ? t1 : 0;
    } else return RA_ETF_E_FORMAT;
    if (r.off != len) return RA_ETF_E_FORMAT;
    if (n_entries) *n_entries = n;
    return RA_ETF_OK;
}

int ra_etf_decode_aer_reply(const uint8_t* msg, size_t len, ra_event* ev, ra_etf_id* peer)
{
    if (!msg || !ev || len < 2 || msg[0] != 131) return RA_ETF_E_FORMAT;
    rd_t r = { msg, len, 1, 0 };
    if (rd_tuple(&r) != 2) return r.err ? r.err : RA_ETF_E_FORMAT;
    rd_id(&r, peer);
    if (rd_tuple(&r) != 6 || !rd_atom_is(&r, "append_entries_reply")) return r.err ? r.err : RA_ETF_E_FORMAT;
    memset(ev, 0, sizeof *ev);
    ev->type = RA_EV_AER_REPLY;
    ev->from_slot = RA_NO_SLOT;
    ev->term = rd_uint(&r);
    char b[256];
    if (rd_atom(&r, b) < 0) return r.err;
    if (!strcmp(b, "true")) ev->d = 1; else if (!strcmp(b, "false")) ev->d = 0; else return RA_ETF_E_FORMAT;
    ev->a = rd_uint(&r);                                            /* next_index */
    ev->b = rd_uint(&r);                                            /* last_index */
    ev->c = rd_uint(&r);                                            /* last_term  */
    if (r.err) return r.err;
    return r.off == len ? RA_ETF_OK : RA_ETF_E_FORMAT;
}

/* ---- writers (out == NULL: size only) ----------------------------------------------------------------- */
typedef struct { uint8_t* p; size_t cap, off; } wr_t;
static void wr_bytes(wr_t* w, const void* b, size_t n)
{
    if (w->p && w->off + n <= w->cap) memcpy(w->p + w->off, b, n);
    w->off += n;
}
static void wr_u8(wr_t* w, uint8_t v) { wr_bytes(w, &v, 1); }
static void wr_uint(wr_t* w, uint64_t v)
{   /* what term_to_binary emits: the smallest of SMALL_INTEGER / INTEGER / SMALL_BIG */
    if (v < 256) { wr_u8(w, T_SMALL_INT); wr_u8(w, (uint8_t)v); return; }
    if (v < 0x80000000ull) {
        uint8_t b[5] = { T_INT, (uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v };
        wr_bytes(w, b, 5); return;
    }
    uint8_t n = 0; uint64_t t = v;
    while (t) { n++; t >>= 8; }
    wr_u8(w, T_SMALL_BIG); wr_u8(w, n); wr_u8(w, 0);
    for (uint8_t i = 0; i < n; i++) wr_u8(w, (uint8_t)(v >> (8 * i)));
}
static void wr_atom(wr_t* w, const char* a)
{
    const size_t n = strlen(a);
    wr_u8(w, T_SMALL_ATOM_UTF8); wr_u8(w, (uint8_t)n); wr_bytes(w, a, n);
}
static void wr_id(wr_t* w, const ra_etf_id* id)
{
    wr_u8(w, T_SMALL_TUPLE); wr_u8(w, 2); wr_atom(w, id->name); wr_atom(w, id->node);
}

size_t ra_etf_encode_aer(const ra_event* ev, const ra_etf_id* leader,
                         const uint8_t* const* cmds, const uint32_t* cmd_len, uint8_t* out, size_t cap)
{
    if (!ev || !leader || ev->type != RA_EV_AER || (ev->n && (!cmds || !cmd_len))) return 0;
    wr_t w = { out, cap, 0 };
    wr_u8(&w, 131); wr_u8(&w, T_SMALL_TUPLE); wr_u8(&w, 7);
    wr_atom(&w, "append_entries_rpc");
    wr_uint(&w, ev->term); wr_id(&w, leader);
    wr_uint(&w, ev->c); wr_uint(&w, ev->a); wr_uint(&w, ev->b);
    if (ev->n == 0) wr_u8(&w, T_NIL);
    else {
        uint8_t h[5] = { T_LIST, 0, 0, (uint8_t)(ev->n >> 8), (uint8_t)ev->n };
        wr_bytes(&w, h, 5);
        for (uint32_t i = 0; i < ev->n; i++) {
            const uint64_t term = (ev->n1 == 0 || i < ev->n1) ? ev->d : ev->e;
            wr_u8(&w, T_SMALL_TUPLE); wr_u8(&w, 3);
            wr_uint(&w, ev->a + 1 + i); wr_uint(&w, term);
            wr_bytes(&w, cmds[i], cmd_len[i]);
        }
        wr_u8(&w, T_NIL);
    }
    return (out && w.off > cap) ? 0 : w.off;
}

size_t ra_etf_encode_aer_reply(const ra_event* ev, const ra_etf_id* peer, uint8_t* out, size_t cap)
{
    if (!ev || !peer || ev->type != RA_EV_AER_REPLY) return 0;
    wr_t w = { out, cap, 0 };
    wr_u8(&w, 131); wr_u8(&w, T_SMALL_TUPLE); wr_u8(&w, 2);
    wr_id(&w, peer);
    wr_u8(&w, T_SMALL_TUPLE); wr_u8(&w, 6);
    wr_atom(&w, "append_entries_reply");
    wr_uint(&w, ev->term);
    wr_atom(&w, ev->d ? "true" : "false");
    wr_uint(&w, ev->a); wr_uint(&w, ev->b); wr_uint(&w, ev->c);
    return (out && w.off > cap) ? 0 : w.off;
}
