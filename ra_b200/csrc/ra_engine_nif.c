/* ra_engine_nif.c -- dirty-NIF shim over include/ra_engine.h.
 * erl_nif.h does not exist in this image, so the body is compiled only when the build defines
 * RA_HAVE_ERL_NIF on a box with an OTP toolchain; it is not built or tested here. */
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

static ErlNifFunc funcs[] = {
    {"new", 3, new_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"step", 2, step_nif, ERL_NIF_DIRTY_JOB_CPU_BOUND},
};
ERL_NIF_INIT(ra_engine_nif, funcs, load, NULL, NULL, NULL)
#else
typedef int ra_engine_nif_not_built_without_erl_nif_h;
#endif
