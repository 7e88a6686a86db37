%% ra_log -- B1 harness delegate: the module literally named `ra_log` that ra_server.erl calls,
%% forwarding to the reference's own in-memory model test/ra_log_memory.erl.  This is what
%% ra_server_SUITE:setup_log/0 (test/ra_server_SUITE.erl:168-246) installs with meck; as a plain
%% module it needs no meck and no hex dependency.  Compile it INSTEAD of the reference's src/ra_log.erl
%% (see erlang/b1/Makefile).  SOURCE ONLY in the build image.
-module(ra_log).
-export([init/1, recover_snapshot/1, snapshot_state/1, set_snapshot_state/2, install_snapshot/4,
         snapshot_index_term/1, fold/5, fold/6, release_resources/3, overview/1, append_sync/2,
         write_config/2, next_index/1, has_pending/1, append/2, write/2, write_sparse/3,
         handle_event/2, last_written/1, last_index_term/1, set_last_index/2, fetch_term/2,
         exists/2, update_release_cursor/5, tick/2, close/1, can_write/1, needs_cache_flush/1]).

init(C) -> ra_log_memory:init(C).
recover_snapshot(L) -> ra_log_memory:recover_snapshot(L).
snapshot_state(L) -> ra_log_memory:snapshot_state(L).
set_snapshot_state(S, L) -> ra_log_memory:set_snapshot_state(S, L).
install_snapshot(A, B, C, D) -> ra_log_memory:install_snapshot(A, B, C, D).
snapshot_index_term(L) -> ra_log_memory:snapshot_index_term(L).
fold(A, B, C, D, E) -> ra_log_memory:fold(A, B, C, D, E).
fold(A, B, C, D, E, _) -> ra_log_memory:fold(A, B, C, D, E).
release_resources(A, B, C) -> ra_log_memory:release_resources(A, B, C).
overview(L) -> ra_log_memory:overview(L).
append_sync({Idx, Term, _} = E, L0) ->
    L1 = ra_log_memory:append(E, L0),
    {L, _} = ra_log_memory:handle_event({written, Term, [Idx]}, L1),
    L.
write_config(C, L) -> ra_log_memory:write_config(C, L).
next_index(L) -> ra_log_memory:next_index(L).
has_pending(_) -> false.
append(E, L) -> ra_log_memory:append(E, L).
write(Es, L) -> ra_log_memory:write(Es, L).
write_sparse(A, B, C) -> ra_log_memory:write_sparse(A, B, C).
handle_event(E, L) -> ra_log_memory:handle_event(E, L).
last_written(L) -> ra_log_memory:last_written(L).
last_index_term(L) -> ra_log_memory:last_index_term(L).
set_last_index(I, L) -> ra_log_memory:set_last_index(I, L).
fetch_term(I, L) -> ra_log_memory:fetch_term(I, L).
exists({Idx, Term}, L) ->
    case ra_log_memory:fetch_term(Idx, L) of
        {Term, Log} -> {true, Log};
        {_, Log} -> {false, Log}
    end.
update_release_cursor(A, B, C, D, E) -> ra_log_memory:update_release_cursor(A, B, C, D, E).
tick(_, L) -> L.
close(_) -> ok.
can_write(_) -> true.
needs_cache_flush(_) -> false.
