%% ra_engine_nif -- Erlang face of the dirty-NIF shim over include/ra_engine.h
%% (ra_b200/csrc/ra_engine_nif.c).  Every function is ERL_NIF_DIRTY_JOB_CPU_BOUND: a call enqueues a
%% host->device copy, the step kernels and the write-back of the outputs and waits for one CUDA event,
%% far beyond the 1 ms reduction budget of a normal scheduler.
%%
%% Records travel as binaries of the ABI structs; ra_engine_codec builds / matches them.
%% SOURCE ONLY in the build image (no OTP toolchain there).
-module(ra_engine_nif).

-export([new/3, load_rows/2, read_rows/2, reset_empty/1, step/2, step_host/2,
         submit/2, submit_host/2, collect/1, counters/1]).
-on_load(init/0).

-type engine() :: reference().
-type status() :: ok | {error, integer()}.      %% enum ra_status of include/ra_engine.h

init() ->
    Dir = case code:priv_dir(ra_engine) of
              {error, _} -> "priv";
              D -> D
          end,
    erlang:load_nif(filename:join(Dir, "ra_engine_nif"), 0).

%% ra_engine_create: Groups x Members rows on CUDA device Device
-spec new(pos_integer(), 1..8, non_neg_integer()) -> {ok, engine()} | {error, integer()}.
new(_Groups, _Members, _Device) -> erlang:nif_error(nif_not_loaded).

%% ra_engine_load_rows: RowsBin = << <<(ra_engine_codec:encode_row(R))/binary>> ... >>
-spec load_rows(engine(), binary()) -> status().
load_rows(_Eng, _RowsBin) -> erlang:nif_error(nif_not_loaded).

%% ra_engine_read_rows: RowIdsBin = ra_engine_codec:row_ids(Rows) -> 544 bytes per row
-spec read_rows(engine(), binary()) -> binary() | {error, integer()}.
read_rows(_Eng, _RowIdsBin) -> erlang:nif_error(nif_not_loaded).

-spec reset_empty(engine()) -> status().
reset_empty(_Eng) -> erlang:nif_error(nif_not_loaded).

%% ra_engine_step: one batch of 64-byte events (events of one row adjacent, mailbox order kept)
%% -> {RPC records (64 B each, addressed), host notes (32 B each)}, both ordered by (row, seq).
%% {error, -5} with outputs pending is handled inside the shim (it fetches again with buffers of
%% the size the engine reported), so a caller never sees it for output capacity.
-spec step(engine(), binary()) -> {binary(), binary()} | {error, integer()}.
step(_Eng, _EventsBin) -> erlang:nif_error(nif_not_loaded).

%% the same for batches of host-origin events only: 32-byte records (ra_engine_step_host)
-spec step_host(engine(), binary()) -> {binary(), binary()} | {error, integer()}.
step_host(_Eng, _HostEventsBin) -> erlang:nif_error(nif_not_loaded).

%% split-phase form (ra_engine_submit / ra_engine_collect): up to two batches in flight per engine;
%% one batcher process can keep several engines busy: submit to each, then collect each
-spec submit(engine(), binary()) -> status().
submit(_Eng, _EventsBin) -> erlang:nif_error(nif_not_loaded).
-spec submit_host(engine(), binary()) -> status().
submit_host(_Eng, _HostEventsBin) -> erlang:nif_error(nif_not_loaded).
-spec collect(engine()) -> {binary(), binary()} | {error, integer()}.
collect(_Eng) -> erlang:nif_error(nif_not_loaded).

%% ra_engine_counters: the aggregate counters and the reference's per-path ones (ra.hrl:324-343)
-spec counters(engine()) -> map() | {error, integer()}.
counters(_Eng) -> erlang:nif_error(nif_not_loaded).
