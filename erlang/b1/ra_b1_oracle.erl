%% ra_b1_oracle -- the TRUE-reference oracle and BEAM baseline harness (SURVEY.md 8c "When an OTP toolchain
%% is available", VERDICT round 1 item J3).  SOURCE ONLY in the build image (no erl / erlc there).
%%
%% It drives the reference's own ra_server:handle_<state>/2 (rabbitmq/ra v3.1.6 src/ra_server.erl) over the
%% reference's in-memory log model (test/ra_log_memory.erl, reached through the `ra_log` delegate of this
%% directory) with the SAME recorded inputs the engine and the C oracle are tested on:
%%
%%   tests/golden/<name>.events.z   zlib of  [ <<N:32/little, N x 64-byte ra_event>> per step ]
%%
%% and writes, for every member row, the ra_row_state record (544 bytes, include/ra_engine.h) the engine's
%% ra_engine_read_rows would return -- `tools/compare_rows.py` diffs that file against the engine / oracle.
%%
%%   erlc -o ebin -I <ra>/src  <ra>/src/{ra_server,ra_lib,ra_machine,ra_machine_simple,ra_seq,ra_range,ra_system,
%%                               ra_env,ra_counters,ra_flru}.erl <ra>/test/ra_log_memory.erl \
%%        erlang/b1/{ra_log,ra_log_meta,ra_b1_oracle}.erl erlang/src/ra_engine_codec.erl
%%   erl -noshell -pa ebin -eval 'ra_b1_oracle:main(["tests/golden/t5_mixed.events.z","16","5","rows.bin"])' -s init stop
%%
%% With "bench" as a 5th argument the replay is repeated and timed instead (events/s and commits/s of the BEAM
%% ra_server path on this box's cores: the number BASELINE.md's B1 row is waiting for).
%%
%% Mapping of the engine's row layout: member (group G, slot S) = row S * Groups + G; its ra_server_id() is
%% {list_to_atom("g<G>_n<S+1>"), node()}; every group is its own cluster of `Members` such ids.
-module(ra_b1_oracle).
-export([main/1]).

-include_lib("ra/src/ra.hrl").

main([File, GroupsS, MembersS, OutFile | Rest]) ->
    Groups = list_to_integer(GroupsS),
    Members = list_to_integer(MembersS),
    {ok, Z} = file:read_file(File),
    Steps = split_steps(zlib:uncompress(Z)),
    case Rest of
        ["bench" | _] ->
            {T, {_, NEv}} = timer:tc(fun() -> replay(Steps, init_all(Groups, Members), Groups, 0) end),
            io:format("B1 BEAM ra_server path: ~b events in ~.3f s = ~.0f events/s (~b schedulers online)~n",
                      [NEv, T / 1.0e6, NEv / (T / 1.0e6), erlang:system_info(schedulers_online)]);
        _ ->
            {States, _} = replay(Steps, init_all(Groups, Members), Groups, 0),
            Rows = [row_state(Row, maps:get(Row, States), Groups, Members)
                    || Row <- lists:seq(0, Groups * Members - 1)],
            ok = file:write_file(OutFile, [ra_engine_codec:encode_row(R) || R <- Rows]),
            io:format("wrote ~b rows to ~s~n", [length(Rows), OutFile])
    end.

split_steps(<<>>) -> [];
split_steps(<<N:32/little, Rest/binary>>) ->
    Sz = N * 64,
    <<Batch:Sz/binary, Tail/binary>> = Rest,
    [[E || <<E:64/binary>> <= Batch] | split_steps(Tail)].

%% ---- members ------------------------------------------------------------------------------------
id(G, S) -> {list_to_atom("g" ++ integer_to_list(G) ++ "_n" ++ integer_to_list(S + 1)), node()}.

%% ra_server_SUITE:empty_state/2 (test/ra_server_SUITE.erl:4022-4032): ra_server:init/1 + recover on a fresh log
init_all(Groups, Members) ->
    maps:from_list(
      [begin
           Ids = [id(G, S1) || S1 <- lists:seq(0, Members - 1)],
           {Name, _} = Id = id(G, S),
           St = ra_server:recover(
                  ra_server:init(#{cluster_name => list_to_atom("c" ++ integer_to_list(G)),
                                   id => Id,
                                   uid => atom_to_binary(Name, utf8),
                                   initial_members => Ids,
                                   log_init_args => #{uid => atom_to_binary(Name, utf8)},
                                   machine => {simple, fun(E, _) -> E end, <<>>}})),
           {S * Groups + G, {follower, St}}
       end || G <- lists:seq(0, Groups - 1), S <- lists:seq(0, Members - 1)]).

%% ---- replay --------------------------------------------------------------------------------------
replay([], States, _Groups, N) -> {States, N};
replay([Batch | T], States0, Groups, N0) ->
    {States, N} = lists:foldl(fun(Ev, {Ss, N1}) -> {event(Ev, Ss, Groups), N1 + 1} end, {States0, N0}, Batch),
    replay(T, States, Groups, N).

event(<<Row:32/little, _/binary>> = Ev, States, Groups) ->
    G = Row rem Groups,
    IdOf = fun(_R, Slot) -> id(G, Slot) end,
    {Row, _From, _Seq, _Flags, Msg0} = ra_engine_codec:decode_record(Ev, IdOf),
    {Role, St} = maps:get(Row, States),
    Msg = token_in(Row, to_ra_msg(Msg0, Ev, St), St),
    put(cur_row, Row),
    States#{Row => dispatch(Role, [Msg], St)}.

%% Pre-vote tokens: ra_server makes a reference per pre-vote election (call_for_election/3 :2873-2897), the engine
%% counts them (token = the member's n-th pre-vote election).  The recorded traces carry the engine's integers,
%% so the harness counts this member's pre-vote elections too and hands ra_server its own current token when a
%% result names the current election, a foreign reference otherwise.
token_in(Row, #pre_vote_result{token = I} = R, St) when is_integer(I) ->
    case get({pv_count, Row}) of
        I -> R#pre_vote_result{token = maps:get(pre_vote_token, St, make_ref())};
        _ -> R#pre_vote_result{token = make_ref()}
    end;
token_in(_Row, Msg, _St) -> Msg.

count_pre_votes(Effects) ->
    case [x || {send_vote_requests, [{_, #pre_vote_rpc{}} | _]} <- lists:flatten(Effects)] of
        [] -> ok;
        L ->
            Row = get(cur_row),
            put({pv_count, Row}, (case get({pv_count, Row}) of undefined -> 0; N -> N end) + length(L))
    end.

%% what ra_engine_codec hands back -> the ra_msg() ra_server takes
to_ra_msg({aer, Rpc, {From, To}, {N1, D, E}}, _Ev, _St) ->
    Entries = [{I, case N1 =:= 0 orelse I - From < N1 of true -> D; false -> E end, usr(I)}
               || I <- lists:seq(From, To)],
    Rpc#append_entries_rpc{entries = Entries};
to_ra_msg({command, N, true}, _Ev, _St) when N >= 1 -> {command, {noop, #{from => undefined, ts => 0}, 0}};
to_ra_msg({command, 1, false}, _Ev, _St) -> {command, usr(cmd)};
to_ra_msg({command, N, false}, _Ev, _St) -> {commands, [usr(cmd) || _ <- lists:seq(1, N)]};
to_ra_msg(tick, _Ev, _St) -> {tick, 0};                       %% leader tick -> make_rpcs (ra_server_proc.erl:610-613)
to_ra_msg({ra_log_event, {written, Term, {A, B}}}, _Ev, _St) -> {ra_log_event, {written, Term, [{A, B}]}};
to_ra_msg(Msg, _Ev, _St) -> Msg.               %% RPC records, election_timeout, await_condition_timeout, pipeline_rpcs

usr(Data) -> {'$usr', #{from => undefined, ts => 0}, Data, noreply}.

%% one mailbox turn of ra_server_proc: the message, then every {next_event, _} it produced, front first
%% (gen_statem semantics, ra_server_proc.erl:1574-1577); role changes take effect between events
dispatch(Role, [], St) -> {Role, St};
dispatch(Role, [Msg | Q], St0) ->
    Fun = case Role of
              leader -> handle_leader;
              follower -> handle_follower;
              candidate -> handle_candidate;
              pre_vote -> handle_pre_vote;
              await_condition -> handle_await_condition
          end,
    {Next, St1, Effects} =
        try ra_server:Fun(Msg, St0)
        catch throw:{N, S, E} when is_atom(N), is_map(S), is_list(E) -> {N, S, E}
        end,
    count_pre_votes(Effects),
    St = case Next =/= Role of
             true -> element(1, ra_server:handle_state_enter(Next, Role, St1));   %% become/3 (:2153-2177)
             false -> St1
         end,
    Nexts = [E || {next_event, E} <- lists:flatten(Effects)] ++
            [E || {next_event, _Type, E} <- lists:flatten(Effects)],
    case Next of
        R when R =:= leader; R =:= follower; R =:= candidate; R =:= pre_vote; R =:= await_condition ->
            dispatch(Next, Nexts ++ Q, St);
        _ -> {Next, St}                         %% terminating_* / receive_snapshot: outside the hot path
    end.

%% ---- ra_server_state() -> the map ra_engine_codec:encode_row/1 takes -----------------------------
row_state(Row, {Role, #{cfg := Cfg, current_term := Term, commit_index := CI, last_applied := LA,
                        cluster := Cluster, log := Log} = St}, Groups, Members) ->
    G = Row rem Groups, S = Row div Groups,
    SlotOf = fun(undefined) -> undefined;
                (Id) -> hd([Sl || Sl <- lists:seq(0, Members - 1), id(G, Sl) =:= Id] ++ [undefined])
             end,
    {LastIdx, LastTerm} = ra_log:last_index_term(Log),
    Ov = ra_log:overview(Log),
    First = maps:get(first_index, Ov, 0),
    Runs = runs(First, LastIdx, Log),
    Peers = [begin
                 P = maps:get(id(G, Sl), Cluster, #{}),
                 #{next_index => maps:get(next_index, P, 1), match_index => maps:get(match_index, P, 0),
                   commit_index_sent => maps:get(commit_index_sent, P, 0),
                   status => maps:get(status, P, normal),
                   voter => maps:get(membership, maps:get(voter_status, P, #{}), voter) =:= voter}
             end || Sl <- lists:seq(0, Members - 1)],
    #{row => Row, role => Role, self_slot => S, n_members => Members,
      leader_slot => SlotOf(maps:get(leader_id, St, undefined)),
      voted_for => SlotOf(maps:get(voted_for, St, undefined)),
      membership => maps:get(membership, St, voter), condition => 0,
      votes => maps:get(votes, St, 0),
      machine_version => element(#cfg.machine_version, Cfg),
      effective_machine_version => element(#cfg.effective_machine_version, Cfg),
      current_term => Term, commit_index => CI, last_applied => LA,
      first_index => First, last_index => LastIdx, last_term => LastTerm,
      last_written => ra_log:last_written(Log),
      snapshot => ra_log:snapshot_index_term(Log),
      runs => Runs, peers => Peers}.

%% index -> term knowledge as (start, term) runs, the engine's log view
runs(First, Last, _Log) when First > Last -> [];
runs(First, Last, Log) ->
    {Rs, _} = lists:foldl(fun(I, {Acc, Prev}) ->
                                  {T, _} = ra_log:fetch_term(I, Log),
                                  case T =:= Prev of
                                      true -> {Acc, Prev};
                                      false -> {[{I, T} | Acc], T}
                                  end
                          end, {[], make_ref()}, lists:seq(First, Last)),
    lists:reverse(Rs).
