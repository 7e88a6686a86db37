%% ra_engine_batcher -- the effect-interpreter adapter between ra_server_proc and the GPU engine
%% (SURVEY.md section 8f-1).
%%
%% Seam replaced (rabbitmq/ra v3.1.6): ra_server_proc:handle_leader/2 (src/ra_server_proc.erl:1353-1377)
%% and handle_raft_state/3 (:1379-1394) call ra_server:handle_<state>(Msg, ServerState) and get
%% {NextState, ServerState, Effects} back.  With the engine, the per-member Raft state lives in one row
%% of the GPU's struct-of-arrays, so a ra_server_proc instead does
%%
%%     {NextState, Effects} = ra_engine_batcher:handle(Batcher, ServerId, Msg)
%%
%% and runs Effects through its unmodified handle_effects/5 (:1527-1857).  One batcher serves all
%% members that live on one GPU: it collects what arrived while the previous batch was on the device
%% (continuous batching), keeps the events of one member adjacent and in mailbox order (engine contract
%% item 1: at most RA_LOCAL_CAP = 4 per member per batch, the rest waits for the next batch), calls the
%% dirty NIF once, and turns the outputs back into effects, per member:
%%
%%   RPC records (addressed to the DESTINATION row, from_slot = the sender's slot)  -> effects of the SENDER:
%%     AER              {send_rpc, Peer, #append_entries_rpc{entries = read From..To from ra_log}}   (:2401-2418)
%%     AER_REPLY        {cast, Leader, {Self, #append_entries_reply{}}}                              (:3597-3604)
%%     REQUEST_VOTE /   {send_vote_requests, [{Peer, #request_vote_rpc{} | #pre_vote_rpc{}}]}        (:2853-2897)
%%     PRE_VOTE
%%     *_VOTE_RES       {reply, #request_vote_result{} | #pre_vote_result{}}                         (:1467-1513, :2899-2956)
%%     HEARTBEAT_RPC    {send_rpc, Peer, #heartbeat_rpc{}};  HEARTBEAT_REPLY  {cast, Leader, {Self, #heartbeat_reply{}}}
%%   host notes (what the engine cannot do itself):
%%     WAL_APPEND       {engine_log, {append | write, From, To, Term}}   ra_log:append / ra_log:write of the payloads
%%                                                                       the proc kept for the commands / AER it handed in
%%     TRUNCATE         {engine_log, {set_last_index, Idx}}              ra_log:set_last_index/2 (:1301)
%%     COMMIT           {aux, eval}                                      (:3611-3614)
%%     APPLY            {engine_apply, From, To}                         ra_machine:apply/3 over the entries, then the
%%                                                                       machine's own effects / {notify,_} / {reply,_,_,_}
%%     SEND_SNAPSHOT    {send_snapshot, Peer, {Module, Ref, Self, Term}} (:2395)
%%     CANCEL_SNAPSHOT_RETRY  {cancel_snapshot_retry_timer, Peer}        (:2342)
%%     NOT_LEADER       {reply, {error, not_leader}} / redirect to the leader the note names
%%     QUERY_INDEX / QUERY_AGREED / QUERY_APPLY   the consistent-query bookkeeping of :3722-3795
%%     STATUS           the next gen_statem state; TERM_VOTE_CHANGED -> ra_log_meta:store_sync of term and
%%                      voted_for BEFORE any record of that member is released (:3024-3025);
%%                      LEADER_MSG -> {record_leader_msg, Leader} (:1280); START_ELECTION_TMO ->
%%                      start_election_timeout (:2934,2942); PIPELINE_PENDING -> {next_event, info, pipeline_rpcs};
%%                      FATAL -> exit(Reason) as the reference would; NOTE_OVERFLOW -> the unconsumed
%%                      events are put back at the head of the member's queue.
%%
%% SOURCE ONLY in the build image (no OTP toolchain): never compiled there.  The C ABI it drives is
%% exercised end to end by ra_b200/csrc/host_flood.cu and the Python tests.
-module(ra_engine_batcher).
-behaviour(gen_server).

-export([start_link/1, register_member/4, handle/3, flush/1]).
-export([init/1, handle_call/3, handle_cast/2, handle_info/2, terminate/2]).

-include_lib("ra/src/ra.hrl").

-define(LOCAL_CAP, 4).                      %% RA_LOCAL_CAP
-define(BATCH_INTERVAL_MS, 1).
%% RA_NOTE_STATUS aux flags
-define(ST_TERM_VOTE_CHANGED, 16#0001).
-define(ST_ROLE_CHANGED, 16#0002).
-define(ST_LEADER_MSG, 16#0004).
-define(ST_START_ELECTION_TMO, 16#0008).
-define(ST_PIPELINE_PENDING, 16#0020).
-define(ST_FATAL, 16#0040).
-define(ST_CMD_POSTPONED, 16#0080).
-define(ST_NOTE_OVERFLOW, 16#0200).

-record(member, {id :: ra_server_id(),
                 row :: non_neg_integer(),
                 group :: non_neg_integer(),
                 slot :: 0..7,
                 role = follower :: atom(),
                 queue = queue:new() :: queue:queue({gen_server:from(), term()})}).

-record(st, {eng :: reference(),
             n_groups :: pos_integer(),
             members = #{} :: #{ra_server_id() => #member{}},
             rows = #{} :: #{non_neg_integer() => ra_server_id()},
             tokens = #{} :: #{reference() | integer() => integer() | reference()},
             next_token = 1 :: pos_integer(),
             timer :: undefined | reference()}).

%% ------------------------------------------------------------------------------------------------
start_link(#{groups := _, members := _, device := _} = Cfg) ->
    gen_server:start_link(?MODULE, Cfg, []).

%% A member (ra_server_proc) announces itself: its id, its group number on this engine and its slot in
%% the group (position in the sorted member list, as the placement of DESIGN.md section 7 assumes).
%% InitRow is the map of ra_engine_codec:encode_row/1 built from ra_server:init/1's values
%% (src/ra_server.erl:434-457) and the recovered log tail.
register_member(Batcher, ServerId, {Group, Slot}, InitRow) ->
    gen_server:call(Batcher, {register, ServerId, Group, Slot, InitRow}, infinity).

%% What ra_server_proc calls in place of ra_server:handle_<state>(Msg, ServerState).
%% Blocks the calling gen_statem until the batch that carries Msg has been evaluated, exactly as the
%% synchronous call into ra_server does today; returns what that call returns, minus the state map.
-spec handle(pid(), ra_server_id(), term()) -> {ra_state(), [term()]}.
handle(Batcher, ServerId, Msg) ->
    gen_server:call(Batcher, {handle, ServerId, Msg}, infinity).

flush(Batcher) ->
    gen_server:call(Batcher, flush, infinity).

%% ------------------------------------------------------------------------------------------------
init(#{groups := G, members := M, device := D}) ->
    {ok, Eng} = ra_engine_nif:new(G, M, D),
    ok = ra_engine_nif:reset_empty(Eng),
    {ok, #st{eng = Eng, n_groups = G}}.

handle_call({register, Id, Group, Slot, InitRow}, _From, #st{eng = Eng, n_groups = G} = S) ->
    Row = Slot * G + Group,                                   %% slot-major rows (include/ra_engine.h)
    ok = ra_engine_nif:load_rows(Eng, ra_engine_codec:encode_row(InitRow#{row => Row, self_slot => Slot})),
    Mem = #member{id = Id, row = Row, group = Group, slot = Slot,
                  role = maps:get(role, InitRow, follower)},
    {reply, {ok, Row}, S#st{members = (S#st.members)#{Id => Mem}, rows = (S#st.rows)#{Row => Id}}};
handle_call({handle, Id, Msg}, From, #st{members = Ms} = S0) ->
    #member{queue = Q} = Mem = maps:get(Id, Ms),
    S = S0#st{members = Ms#{Id => Mem#member{queue = queue:in({From, Msg}, Q)}}},
    {noreply, arm(S)};
handle_call(flush, _From, S) ->
    {reply, ok, run_batch(S)}.

handle_cast(_, S) -> {noreply, S}.

handle_info(batch, S) ->
    {noreply, run_batch(S#st{timer = undefined})};
handle_info(_, S) ->
    {noreply, S}.

terminate(_, _) -> ok.

arm(#st{timer = undefined} = S) ->
    S#st{timer = erlang:send_after(?BATCH_INTERVAL_MS, self(), batch)};
arm(S) -> S.

%% ------------------------------------------------------------------------------------------------
%% one batch
%% ------------------------------------------------------------------------------------------------
run_batch(#st{eng = Eng, members = Ms0} = S0) ->
    %% up to LOCAL_CAP queued messages per member, adjacent, in arrival order
    {Taken, Ms1} =
        maps:fold(fun(Id, #member{queue = Q0} = M, {Acc, MsAcc}) ->
                          {Items, Q} = take(?LOCAL_CAP, Q0, []),
                          case Items of
                              [] -> {Acc, MsAcc};
                              _ -> {[{M#member.row, Id, Items} | Acc], MsAcc#{Id => M#member{queue = Q}}}
                          end
                  end, {[], Ms0}, Ms0),
    case Taken of
        [] -> S0;
        _ ->
            {S1, Events} = encode_batch(lists:keysort(1, Taken), S0#st{members = Ms1}),
            case ra_engine_nif:step(Eng, iolist_to_binary(Events)) of
                {MsgsBin, NotesBin} ->
                    S2 = dispatch(Taken, MsgsBin, NotesBin, S1),
                    case lists:any(fun(#member{queue = Q}) -> not queue:is_empty(Q) end,
                                   maps:values(S2#st.members)) of
                        true -> arm(S2);
                        false -> S2
                    end;
                {error, Code} ->
                    %% -4 ungrouped / -5 capacity / -1 bad row: a bug of this module, not of a member
                    exit({ra_engine_step, Code})
            end
    end.

take(0, Q, Acc) -> {lists:reverse(Acc), Q};
take(N, Q0, Acc) ->
    case queue:out(Q0) of
        {{value, It}, Q} -> take(N - 1, Q, [It | Acc]);
        {empty, Q} -> {lists:reverse(Acc), Q}
    end.

encode_batch(Taken, S0) ->
    {Events, S1} =
        lists:mapfoldl(
          fun({Row, Id, Items}, S) ->
                  #member{group = G} = maps:get(Id, S#st.members),
                  SlotOf = slot_fun(G, S),
                  lists:mapfoldl(fun({_From, Msg0}, SA) ->
                                         {Msg, SB} = intern_token(Msg0, SA),
                                         {ra_engine_codec:encode_event(Row, host_form(Msg), SlotOf), SB}
                                 end, S, Items)
          end, S0, Taken),
    {S1, Events}.

%% ra_msg() shapes of ra_server -> what the codec takes for host-origin events
host_form({command, {noop, _, _}}) -> {noop_command, 1};
host_form({command, _}) -> {command, 1};
host_form({commands, Cmds}) -> {command, length(Cmds)};
host_form({ra_log_event, {written, Term, [{From, To}]}}) -> {ra_log_event, {written, Term, {From, To}}};
host_form({ra_log_event, {written, Term, [Idx]}}) when is_integer(Idx) -> {ra_log_event, {written, Term, {Idx, Idx}}};
host_form({tick, _}) -> tick;
host_form(Other) -> Other.                 %% RPC records, election_timeout, pipeline_rpcs, ...

%% pre-vote tokens are references; the engine compares 64-bit integers
intern_token(#pre_vote_rpc{token = T} = R, S0) when is_reference(T) ->
    {I, S} = token_id(T, S0), {R#pre_vote_rpc{token = I}, S};
intern_token({P, #pre_vote_result{token = T} = R}, S0) when is_reference(T) ->
    {I, S} = token_id(T, S0), {{P, R#pre_vote_result{token = I}}, S};
intern_token(Msg, S) -> {Msg, S}.

token_id(T, #st{tokens = Tk, next_token = N} = S) ->
    case Tk of
        #{T := I} -> {I, S};
        _ -> {N, S#st{tokens = Tk#{T => N, N => T}, next_token = N + 1}}
    end.

slot_fun(Group, #st{members = Ms}) ->
    fun(Id) -> case Ms of
                   #{Id := #member{group = Group, slot = Sl}} -> Sl;
                   _ -> 16#FF                               %% not a member of this group
               end
    end.

id_fun(#st{rows = Rows, n_groups = G}) ->
    fun(Row, Slot) -> maps:get(Slot * G + (Row rem G), Rows, undefined) end.

%% ------------------------------------------------------------------------------------------------
%% outputs -> {NextState, Effects} per member, replied to the waiting ra_server_procs
%% ------------------------------------------------------------------------------------------------
dispatch(Taken, MsgsBin, NotesBin, #st{n_groups = G, rows = Rows} = S0) ->
    IdOf = id_fun(S0),
    %% records belong to the effects of their SENDER: row of (group of the destination, from_slot)
    Recs = lists:foldl(
             fun({DstRow, From, _Seq, _Flags, Msg}, Acc) ->
                     Sender = case From of 16#FF -> DstRow; _ -> From * G + (DstRow rem G) end,
                     maps:update_with(Sender, fun(L) -> [{DstRow, Msg} | L] end, [{DstRow, Msg}], Acc)
             end, #{}, ra_engine_codec:decode_records(MsgsBin, IdOf)),
    Notes = lists:foldl(fun({Row, N, Aux}, Acc) ->
                                maps:update_with(Row, fun(L) -> [{N, Aux} | L] end, [{N, Aux}], Acc)
                        end, #{}, ra_engine_codec:decode_notes(NotesBin)),
    lists:foldl(
      fun({Row, Id, Items}, S) ->
              Mem0 = maps:get(Id, S#st.members),
              RowNotes = lists:reverse(maps:get(Row, Notes, [])),
              RowRecs = lists:reverse(maps:get(Row, Recs, [])),
              {Role, Flags, Unconsumed, NoteEffs} = note_effects(RowNotes, Mem0, S),
              RecEffs = [record_effect(DstRow, Msg, Mem0, Rows, S) || {DstRow, Msg} <- RowRecs],
              %% persist term / voted_for before any record of this member leaves (:3024-3025)
              Pre = case Flags band ?ST_TERM_VOTE_CHANGED of
                        0 -> [];
                        _ -> [{engine_persist_term_and_vote, status_of(RowNotes)}]
                    end,
              Effects = Pre ++ NoteEffs ++ RecEffs,
              %% every message of the batch but the last gets {Role, []}: its effects are folded into the
              %% last reply (the engine reports per step, not per message); unconsumed ones go back
              {Done, Back} = lists:split(length(Items) - Unconsumed, Items),
              reply_all(Done, Role, Effects),
              Q = lists:foldr(fun(It, QA) -> queue:in_r(It, QA) end, Mem0#member.queue, Back),
              S#st{members = (S#st.members)#{Id => Mem0#member{role = Role, queue = Q}}}
      end, S0, Taken).

reply_all([], _Role, _Effects) -> ok;
reply_all([{From, _}], Role, Effects) -> gen_server:reply(From, {Role, Effects});
reply_all([{From, _} | T], Role, Effects) -> gen_server:reply(From, {Role, []}), reply_all(T, Role, Effects).

status_of(RowNotes) ->
    case [M || {{status, M}, _} <- RowNotes] of
        [M | _] -> M;
        [] -> #{}
    end.

note_effects(RowNotes, #member{role = Role0, group = Group, slot = Self}, S) ->
    IdOf = id_fun(S),
    SelfRow = Self * S#st.n_groups + Group,
    Folded = lists:foldl(
      fun({{wal_append, From, To, Term}, Aux}, {R, F, U, Acc}) ->
              Op = case R of leader -> append; _ -> write end,
              {R, F bor Aux, U, Acc ++ [{engine_log, {Op, From, To, Term}}]};
         ({{truncate, Idx, _Term}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{engine_log, {set_last_index, Idx}}]};
         ({{commit, _Old, _New}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{aux, eval}]};
         ({{apply, From, To}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{engine_apply, From, To}]};
         ({{send_snapshot, Slot, SnapIdx}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{engine_send_snapshot, IdOf(SelfRow, Slot), SnapIdx}]};
         ({{cancel_snapshot_retry, Slot}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{cancel_snapshot_retry_timer, IdOf(SelfRow, Slot)}]};
         ({{not_leader, _N, Leader}, Aux}, {R, F, U, Acc}) ->
              L = case Leader of undefined -> undefined; _ -> IdOf(SelfRow, Leader) end,
              {R, F bor Aux, U, Acc ++ [{reply, {error, {not_leader, L}}}]};
         ({{query_index, QI, Commit}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{engine_query_waiting, QI, Commit}]};
         ({{query_agreed, QI}, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [{engine_query_agreed, QI}]};
         ({query_apply, Aux}, {R, F, U, Acc}) ->
              {R, F bor Aux, U, Acc ++ [engine_query_apply]};
         ({{status, #{flags := Fl, role := NewRole, leader := Ld, fatal := Fatal, unconsumed := Un}}, _Aux},
          {_R, F, _U, Acc}) ->
              Fatal =:= 0 orelse exit({ra_engine_fatal, Fatal}),          %% the reference would exit too (:840, ?assert)
              E0 = case Fl band ?ST_LEADER_MSG of
                       0 -> [];
                       _ -> [{record_leader_msg, IdOf(SelfRow, Ld)}]
                   end,
              E1 = case Fl band ?ST_START_ELECTION_TMO of 0 -> E0; _ -> E0 ++ [start_election_timeout] end,
              E2 = case Fl band ?ST_PIPELINE_PENDING of 0 -> E1; _ -> E1 ++ [{next_event, info, pipeline_rpcs}] end,
              E3 = case Fl band ?ST_CMD_POSTPONED of 0 -> E2; _ -> E2 ++ [engine_postpone] end,
              {NewRole, F bor Fl, Un, Acc ++ E3}
      end, {Role0, 0, 0, []}, RowNotes),
    leader_msg_in_aux(Folded).

%% record_leader_msg alone rides in the aux of the member's last note (no STATUS note then); its leader is
%% the sender of the AppendEntries just handled, which the proc knows (the rpc's leader_id)
leader_msg_in_aux({Role, Flags, Un, Effs}) ->
    case Flags band ?ST_LEADER_MSG =/= 0 andalso not lists:keymember(record_leader_msg, 1, Effs) of
        true -> {Role, Flags, Un, Effs ++ [{record_leader_msg, from_rpc}]};
        false -> {Role, Flags, Un, Effs}
    end.

record_effect(DstRow, Msg, #member{id = Self}, Rows, S) ->
    Dst = maps:get(DstRow, Rows, undefined),
    case Msg of
        {aer, #append_entries_rpc{} = Rpc, {FromIdx, ToIdx}, _Runs} ->
            %% the proc fills `entries` from its ra_log: ra_log:fold(FromIdx, ToIdx, ...) as
            %% make_append_entries_rpc/6 does (src/ra_server.erl:2401-2418)
            {engine_send_aer, Dst, Rpc, {FromIdx, ToIdx}};
        {_Peer, #append_entries_reply{} = Reply} ->
            {cast, Dst, {Self, Reply}};
        #request_vote_rpc{} = Rpc ->
            {send_vote_requests, [{Dst, Rpc}]};
        #pre_vote_rpc{token = T} = Rpc ->
            {send_vote_requests, [{Dst, Rpc#pre_vote_rpc{token = maps:get(T, S#st.tokens, T)}}]};
        #request_vote_result{} = Res ->
            {reply, Res};
        #pre_vote_result{token = T} = Res ->
            {reply, Res#pre_vote_result{token = maps:get(T, S#st.tokens, T)}};
        #heartbeat_rpc{} = Rpc ->
            {send_rpc, Dst, Rpc};
        {_Peer, #heartbeat_reply{} = Reply} ->
            {cast, Dst, {Self, Reply}}
    end.
