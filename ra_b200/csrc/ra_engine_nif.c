/* ra_engine_nif.c -- dirty-NIF shim over include/ra_engine.h.
 * erl_nif.h does not exist in this image, so the body is compiled only when the build defines
 * RA_HAVE_ERL_NIF on a box with an OTP toolchain; here it is only syntax-checked against a stub of
 * the erl_nif API (tests/nif_stub/erl_nif.h, tests/test_abi_exports.py) -- it is not built or run.
 * Records travel as binaries of the ABI structs (ra_event 64 B, ra_note 32 B, ra_row_state), which
 * the Erlang side builds / matches with bit syntax. */
#ifdef RA_HAVE_ERL_NIF
#include <erl_nif.h>
#include <string.h>
#include "../../include/ra_engine.h"

static ErlNifResourceType* ENG;

static void eng_dtor(ErlNifEnv* env, void* obj) { (void)env; ra_engine_destroy(*(ra_engine**)obj); }

static int load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info)
{
    (void)priv; (void)info;
    ENG = enif_open_resource_type(env, NULL, "ra_engine", eng_dtor, ERL_NIF_RT_CREATE, NULL);
    return ENG ? 0 : 1;
}

/* new(#{groups, members, device}) */
static ERL_NIF_TERM new_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[])
{
    unsigned groups, members; int device;
    if (argc != 3 || !enif_get_uint(env, argv[0], &groups) || !enif_get_uint(env, argv[1], &members) ||
        !enif_get_int(env, argv[2], &device)) return enif_make_badarg(env);
    ra_engine_cfg cfg; memset(&cfg, 0, sizeof cfg);
    cfg.n_groups = groups; cfg.n_members = members; cfg.device = device;
    ra_engine* e = NULL;
    int rc = ra_engine_create(&cfg, &e);
    if (rc) return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_int(env, rc));
    ra_engine** r = (ra_engine**)enif_alloc_resource(ENG, sizeof(ra_engine*));
    *r = e;
    ERL_NIF_TERM t = enif_make_resource(env, r);
    enif_release_resource(r);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), t);
}

/* step(Ref, EventsBin) -> {MsgsBin, NotesBin} | {error, Code} */
static ERL_NIF_TERM step_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[])
{
    ra_engine** e; ErlNifBinary ev;
    if (argc != 2 || !enif_get_resource(env, argv[0], ENG, (void**)&e) ||
        !enif_inspect_binary(env, argv[1], &ev) || ev.size % sizeof(ra_event)) return enif_make_badarg(env);
    size_t n = ev.size / sizeof(ra_event), nm = 0, nn = 0;
    size_t mc = n * RA_MSG_CAP + 64, nc = n * RA_NOTE_CAP + 64;
    ErlNifBinary m, t;
    if (!enif_alloc_binary(mc * sizeof(ra_event), &m)) return enif_make_badarg(env);
    if (!enif_alloc_binary(nc * sizeof(ra_note), &t)) { enif_release_binary(&m); return enif_make_badarg(env); }
    int rc = ra_engine_step(*e, (const ra_event*)ev.data, n, (ra_event*)m.data, mc, &nm, (ra_note*)t.data, nc, &nn);
    if (rc) {
        enif_release_binary(&m); enif_release_binary(&t);
        return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_int(env, rc));
    }
    enif_realloc_binary(&m, nm * sizeof(ra_event));
    enif_realloc_binary(&t, nn * sizeof(ra_note));
    return enif_make_tuple2(env, enif_make_binary(env, &m), enif_make_binary(env, &t));
}

static ERL_NIF_TERM status_term(ErlNifEnv* env, int rc)
{
    return rc ? enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_int(env, rc)) : enif_make_atom(env, "ok");
}

/* load_rows(Ref, RowsBin) -> ok | {error, Code}: RowsBin = << <<Row:sizeof(ra_row_state)/binary>> ... >> */
static ERL_NIF_TERM load_rows_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[])
{
    ra_engine** e; ErlNifBinary rows;
    if (argc != 2 || !enif_get_resource(env, argv[0], ENG, (void**)&e) ||
        !enif_inspect_binary(env, argv[1], &rows) || rows.size % sizeof(ra_row_state)) return enif_make_badarg(env);
    return status_term(env, ra_engine_load_rows(*e, (const ra_row_state*)rows.data, rows.size / sizeof(ra_row_state)));
}

/* read_rows(Ref, RowIdsBin) -> RowsBin | {error, Code}: RowIdsBin = << <<Row:32/little>> ... >> */
static ERL_NIF_TERM read_rows_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[])
{
    ra_engine** e; ErlNifBinary ids, out;
    if (argc != 2 || !enif_get_resource(env, argv[0], ENG, (void**)&e) ||
        !enif_inspect_binary(env, argv[1], &ids) || ids.size % 4) return enif_make_badarg(env);
    const size_t n = ids.size / 4;
    if (!enif_alloc_binary(n * sizeof(ra_row_state), &out)) return enif_make_badarg(env);
    memset(out.data, 0, out.size);
    for (size_t i = 0; i < n; i++) memcpy(&((ra_row_state*)out.data)[i].row, ids.data + 4 * i, 4);
    int rc = ra_engine_read_rows(*e, (ra_row_state*)out.data, n);
    if (rc) { enif_release_binary(&out); return status_term(env, rc); }
    return enif_make_binary(env, &out);
}

/* reset_empty(Ref) -> ok | {error, Code} */
static ERL_NIF_TERM reset_empty_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[])
{
    ra_engine** e;
    if (argc != 1 || !enif_get_resource(env, argv[0], ENG, (void**)&e)) return enif_make_badarg(env);
    return status_term(env, ra_engine_reset_empty(*e));
}

/* counters(Ref) -> #{events => _, commits => _, ...} | {error, Code} */
static ERL_NIF_TERM counters_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[])
{
    ra_engine** e; ra_counters c;
    if (argc != 1 || !enif_get_resource(env, argv[0], ENG, (void**)&e)) return enif_make_badarg(env);
    int rc = ra_engine_counters(*e, &c);
    if (rc) return status_term(env, rc);
    const char* names[] = {"events", "commits", "applied", "msgs_out", "msgs_dropped", "elections_won", "fatal_rows", "steps"};
    const uint64_t vals[] = {c.events, c.commits, c.applied, c.msgs_out, c.msgs_dropped, c.elections_won, c.fatal_rows, c.steps};
    ERL_NIF_TERM map = enif_make_new_map(env);
    for (int i = 0; i < 8; i++)
        enif_make_map_put(env, map, enif_make_atom(env, names[i]), enif_make_uint64(env, vals[i]), &map);
    return map;
}

static ErlNifFunc funcs[] = {
    {"new", 3, new_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"load_rows", 2, load_rows_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"read_rows", 2, read_rows_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"reset_empty", 1, reset_empty_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"step", 2, step_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"counters", 1, counters_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
};
ERL_NIF_INIT(ra_engine_nif, funcs, load, NULL, NULL, NULL)
#else
typedef int ra_engine_nif_not_built_without_erl_nif_h;
#endif
