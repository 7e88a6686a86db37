/*
 * ra_etf.h -- AppendEntries wire codec: Erlang external term format
 * (erlang:term_to_binary/1, version magic 131) of
 *
 *     #append_entries_rpc{term, leader_id, leader_commit, prev_log_index, prev_log_term, entries}
 *                                                                        rabbitmq/ra src/ra.hrl:122-128
 *     {PeerId, #append_entries_reply{term, success, next_index, last_index, last_term}}   :130-141
 *
 * <-> the engine's 64-byte ra_event records (include/ra_engine.h) plus, for the rpc, one
 * {offset, length} iovec per entry pointing at the entry's command term INSIDE the message, so that a
 * transport can move RPCs between the distribution socket and the GPU inbox without building Erlang terms:
 * the Raft fields go to the engine, the payload bytes go to the WAL / the follower untouched.
 *
 * Host-only C (no CUDA); built as ra_b200/csrc/libra_etf.so.  Byte vectors derived by hand from the
 * documented format pin it in tests/test_etf_codec.py (no OTP toolchain exists in the build image).
 *
 * Tags handled (erts/doc "External Term Format"): SMALL_INTEGER_EXT 97, INTEGER_EXT 98, SMALL_BIG_EXT 110,
 * ATOM_EXT 100, SMALL_ATOM_EXT 115, ATOM_UTF8_EXT 118, SMALL_ATOM_UTF8_EXT 119, SMALL_TUPLE_EXT 104,
 * LARGE_TUPLE_EXT 105, NIL_EXT 106, LIST_EXT 108; and, only to find where an entry's command ends:
 * STRING_EXT 107, BINARY_EXT 109, BIT_BINARY_EXT 77, MAP_EXT 116, NEW_FLOAT_EXT 70, FLOAT_EXT 99,
 * LARGE_BIG_EXT 111, PID_EXT 103, NEW_PID_EXT 88, PORT_EXT 102, NEW_PORT_EXT 89, V4_PORT_EXT 120,
 * REFERENCE_EXT 101, NEW_REFERENCE_EXT 114, NEWER_REFERENCE_EXT 90, EXPORT_EXT 113.
 */
#ifndef RA_ETF_H
#define RA_ETF_H

#include <stddef.h>
#include <stdint.h>
#include "ra_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

/* a ra_server_id() = {Name :: atom(), Node :: atom()} as it appears on the wire */
typedef struct ra_etf_id {
    char     name[256], node[256];       /* NUL-terminated UTF-8 */
} ra_etf_id;

/* one log_entry() = {Index, Term, Command}: where its Command term sits in the message */
typedef struct ra_etf_entry {
    uint64_t index, term;
    uint32_t cmd_off, cmd_len;           /* byte range of the command's external term (no version byte) */
} ra_etf_entry;

enum ra_etf_status {
    RA_ETF_OK = 0,
    RA_ETF_E_TRUNCATED = -1,             /* message ends inside a term                        */
    RA_ETF_E_FORMAT = -2,                /* not the expected record / unsupported tag         */
    RA_ETF_E_RANGE = -3,                 /* an integer does not fit 64 bits / is negative     */
    RA_ETF_E_CAPACITY = -4,              /* more entries than `max_entries`, or output buffer too small */
    RA_ETF_E_RUNS = -5                   /* entries span more than two term runs or are not consecutive:
                                            split the batch (engine contract: an AER record spans <= 2 runs) */
};

/*
 * Decode term_to_binary(#append_entries_rpc{}).  Fills ev (type RA_EV_AER; row and from_slot are the
 * caller's business: it maps `leader` to a slot), the leader id and one ra_etf_entry per entry.
 * ev->n / n1 / d / e describe the entries' term runs as the engine expects them.
 */
int ra_etf_decode_aer(const uint8_t* msg, size_t len, ra_event* ev, ra_etf_id* leader,
                      ra_etf_entry* entries, size_t max_entries, size_t* n_entries);

/*
 * Encode an engine AER record back into term_to_binary(#append_entries_rpc{}).  cmds[i] / cmd_len[i] are the
 * external terms (without version byte) of the ev->n commands, in index order, as read from the log.
 * Returns the number of bytes written, or 0 when `cap` is too small (call with out = NULL to size).
 */
size_t ra_etf_encode_aer(const ra_event* ev, const ra_etf_id* leader,
                         const uint8_t* const* cmds, const uint32_t* cmd_len, uint8_t* out, size_t cap);

/* {PeerId, #append_entries_reply{}} <-> RA_EV_AER_REPLY */
int    ra_etf_decode_aer_reply(const uint8_t* msg, size_t len, ra_event* ev, ra_etf_id* peer);
size_t ra_etf_encode_aer_reply(const ra_event* ev, const ra_etf_id* peer, uint8_t* out, size_t cap);

/* length of the external term starting at p (no version byte), 0 on error: exported for tests */
size_t ra_etf_term_size(const uint8_t* p, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* RA_ETF_H */
