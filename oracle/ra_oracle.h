/*
 * ra_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's Raft core for the hot path
 * (rabbitmq/ra v3.1.6, src/ra_server.erl + the ra_log facade), exposed with the
 * same shape as the engine's C ABI (include/ra_engine.h) so that parity is a
 * backend diff.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (ra_b200/) never does.
 */
#ifndef RA_ORACLE_H
#define RA_ORACLE_H
#include "../include/ra_engine.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ra_oracle ra_oracle;

int  ra_oracle_create(const ra_engine_cfg* cfg, ra_oracle** out);
void ra_oracle_destroy(ra_oracle* o);
int  ra_oracle_load_rows(ra_oracle* o, const ra_row_state* rows, size_t n);
int  ra_oracle_reset_empty(ra_oracle* o);
int  ra_oracle_read_rows(ra_oracle* o, ra_row_state* rows, size_t n);
int  ra_oracle_load_query_state(ra_oracle* o, const ra_query_state* q, size_t n);
int  ra_oracle_read_query_state(ra_oracle* o, ra_query_state* q, size_t n);
int  ra_oracle_step(ra_oracle* o, const ra_event* ev, size_t n_ev,
                    ra_event* msgs, size_t msgs_cap, size_t* n_msgs,
                    ra_note* notes, size_t notes_cap, size_t* n_notes);
/* threads > 1: groups are sharded statically over that many pthreads */
int  ra_oracle_flood(ra_oracle* o, uint32_t n_steps, uint32_t cmds_per_step,
                     uint32_t election_permille, uint64_t seed, uint32_t threads);
int  ra_oracle_flood_faults(ra_oracle* o, uint32_t n_steps, uint32_t cmds_per_step, uint32_t election_permille,
                            uint64_t seed, uint32_t threads, const ra_flood_faults* faults);
int  ra_oracle_step_host(ra_oracle* o, const ra_host_event* ev, size_t n_ev, ra_event* msgs, size_t msgs_cap,
                         size_t* n_msgs, ra_note* notes, size_t notes_cap, size_t* n_notes);
int  ra_oracle_counters(ra_oracle* o, ra_counters* out);
/* this oracle's group g stands for global group offset + g * stride of a flood over total_groups groups (the
   flood host model is keyed by global group / row ids): the checker of runs too big to replay in full */
int  ra_oracle_set_sample(ra_oracle* o, uint32_t stride, uint32_t offset, uint32_t total_groups);

/* in-module KAT of the reference: agreed_commit/1, src/ra_server.erl:3657-3661 */
uint64_t ra_oracle_agreed_commit(const uint64_t* indexes, size_t n);

#ifdef __cplusplus
}
#endif
#endif
