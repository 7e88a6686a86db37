%% ra_bench_config1 -- BASELINE.json configs[0]: 64 ra clusters x 3 members on one BEAM node, machine
%% {simple, fun erlang:'+'/2, 0} (as test/ra_server_SUITE.erl:294), 5 client processes per cluster
%% pipelining integers with a window of 500 (src/ra_bench.erl:18,40), 60 s: committed entries/s through
%% the REAL ra (WAL, segments, gen_statem) -- the reference's own plumbing number.  Needs a full ra build
%% (rebar3 + hex deps gen_batch_server 0.10.0, aten 0.6.0, seshat 1.0.1).  SOURCE ONLY in the build image.
%%
%%   erl -pa _build/default/lib/*/ebin -noshell -eval 'ra_bench_config1:run()' -s init stop
-module(ra_bench_config1).
-export([run/0, run/1]).

-define(CLUSTERS, 64).
-define(MEMBERS, 3).
-define(CLIENTS, 5).
-define(PIPE, 500).

run() -> run(60).

run(Seconds) ->
    {ok, _} = application:ensure_all_started(ra),
    _ = ra_system:start_default(),
    Leaders =
        [begin
             Name = list_to_atom("c" ++ integer_to_list(C)),
             Ids = [{list_to_atom("c" ++ integer_to_list(C) ++ "_" ++ integer_to_list(M)), node()}
                    || M <- lists:seq(1, ?MEMBERS)],
             {ok, Started, []} = ra:start_cluster(default, Name, {simple, fun erlang:'+'/2, 0}, Ids),
             {ok, _, Leader} = ra:members(hd(Started)),
             Leader
         end || C <- lists:seq(1, ?CLUSTERS)],
    Counter = counters:new(1, [write_concurrency]),
    Deadline = erlang:monotonic_time(millisecond) + Seconds * 1000,
    Self = self(),
    Pids = [spawn_link(fun() -> client(L, Counter, Deadline, Self) end)
            || L <- Leaders, _ <- lists:seq(1, ?CLIENTS)],
    T0 = erlang:monotonic_time(millisecond),
    [receive {done, P} -> ok end || P <- Pids],
    Ms = erlang:monotonic_time(millisecond) - T0,
    N = counters:get(Counter, 1),
    io:format("config 1: ~b clusters x ~b members, ~b clients each, window ~b: ~b commits in ~b ms = ~.0f commits/s "
              "(~b schedulers online)~n",
              [?CLUSTERS, ?MEMBERS, ?CLIENTS, ?PIPE, N, Ms, N / (Ms / 1000), erlang:system_info(schedulers_online)]).

%% ra_bench's client (src/ra_bench.erl:89-136 shape): keep ?PIPE commands in flight with
%% ra:pipeline_command/4, count an entry when its applied notification comes back
client(Leader, Counter, Deadline, Parent) ->
    [ra:pipeline_command(Leader, 1, I, low) || I <- lists:seq(1, ?PIPE)],
    loop(Leader, Counter, Deadline, ?PIPE + 1),
    Parent ! {done, self()}.

loop(Leader, Counter, Deadline, Next) ->
    case erlang:monotonic_time(millisecond) >= Deadline of
        true -> ok;
        false ->
            receive
                {ra_event, _From, {applied, Corrs}} ->
                    N = length(Corrs),
                    counters:add(Counter, 1, N),
                    [ra:pipeline_command(Leader, 1, Next + I, low) || I <- lists:seq(0, N - 1)],
                    loop(Leader, Counter, Deadline, Next + N);
                {ra_event, _From, {rejected, {not_leader, NewLeader, _Corr}}} when NewLeader =/= undefined ->
                    loop(NewLeader, Counter, Deadline, Next);
                _ ->
                    loop(Leader, Counter, Deadline, Next)
            after 1000 ->
                    loop(Leader, Counter, Deadline, Next)
            end
    end.
