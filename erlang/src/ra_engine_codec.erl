%% ra_engine_codec -- Erlang terms of the Raft hot path <-> the plain-old-data records of
%% include/ra_engine.h (ra_event 64 B, ra_host_event 32 B, ra_note 32 B, ra_row_state 544 B).
%%
%% Everything is little endian, no padding beyond what the structs show.  Members are identified
%% inside a group by a SLOT 0..M-1; the batcher owns the {ra_server_id() <-> {Row, Slot}} maps
%% (ra_engine_batcher).  Record definitions mirrored: rabbitmq/ra src/ra.hrl:122-169 (RPC records),
%% :63-75 (ra_peer_state()), src/ra_log.erl:73 ({written, Term, Seq}), src/ra_server.erl:144-164
%% (ra_msg()).
%%
%% SOURCE ONLY: no OTP toolchain exists in the build image, so this module has never been compiled
%% there.  Byte layouts are pinned on the C side (tests/test_abi_exports.py, tests/test_etf_codec.py
%% check the same offsets through ctypes).
-module(ra_engine_codec).

-export([encode_event/3, encode_host_event/2, decode_record/2, decode_records/2,
         decode_note/1, decode_notes/1,
         encode_row/1, decode_row/1, row_ids/1,
         term_runs/1]).

-include_lib("ra/src/ra.hrl").

%% enum ra_event_type
-define(EV_AER, 1).
-define(EV_AER_REPLY, 2).
-define(EV_REQUEST_VOTE, 3).
-define(EV_REQUEST_VOTE_RES, 4).
-define(EV_PRE_VOTE, 5).
-define(EV_PRE_VOTE_RES, 6).
-define(EV_WRITTEN, 7).
-define(EV_COMMAND, 8).
-define(EV_ELECTION_TIMEOUT, 9).
-define(EV_AWAIT_COND_TIMEOUT, 10).
-define(EV_PIPELINE_RPCS, 11).
-define(EV_TICK, 12).
-define(EV_HEARTBEAT_RPC, 13).
-define(EV_HEARTBEAT_REPLY, 14).
-define(EV_CONSISTENT_QUERY, 15).
-define(NO_SLOT, 16#FF).
-define(EVF_NOOP, 1).

-type row() :: non_neg_integer().
-type slot() :: 0..7.
%% SlotOf :: fun((ra_server_id()) -> slot() | 255) supplied by the batcher for the member's group.
-type slot_fun() :: fun((ra_server_id()) -> slot() | 255).

%% ---------------------------------------------------------------------------------------------
%% ra_msg() -> one 64-byte ra_event (include/ra_engine.h "One 64-byte wire record")
%% ---------------------------------------------------------------------------------------------
-spec encode_event(row(), term(), slot_fun()) -> binary().
encode_event(Row, #append_entries_rpc{term = T, leader_id = L, leader_commit = C,
                                      prev_log_index = PI, prev_log_term = PT,
                                      entries = Es}, SlotOf) ->
    %% an AER record spans <= 2 term runs (DESIGN.md contract item 2): the batcher splits longer
    %% batches with split_aer/1 before calling this
    {N, N1, D, E} = term_runs(Es),
    ev(Row, ?EV_AER, SlotOf(L), 0, N, N1, T, PI, PT, C, D, E);
encode_event(Row, {Peer, #append_entries_reply{term = T, success = S, next_index = NI,
                                                last_index = LI, last_term = LT}}, SlotOf) ->
    ev(Row, ?EV_AER_REPLY, SlotOf(Peer), 0, 0, 0, T, NI, LI, LT, bool(S), 0);
encode_event(Row, #request_vote_rpc{term = T, candidate_id = Cand, last_log_index = LI,
                                    last_log_term = LT}, SlotOf) ->
    ev(Row, ?EV_REQUEST_VOTE, SlotOf(Cand), 0, 0, 0, T, LI, LT, 0, 0, 0);
encode_event(Row, {Voter, #request_vote_result{term = T, vote_granted = G}}, SlotOf) ->
    ev(Row, ?EV_REQUEST_VOTE_RES, SlotOf(Voter), 0, 0, 0, T, 0, 0, 0, bool(G), 0);
encode_event(Row, #request_vote_result{term = T, vote_granted = G}, _SlotOf) ->
    ev(Row, ?EV_REQUEST_VOTE_RES, ?NO_SLOT, 0, 0, 0, T, 0, 0, 0, bool(G), 0);
encode_event(Row, #pre_vote_rpc{version = V, machine_version = MV, term = T, token = Tok,
                                candidate_id = Cand, last_log_index = LI, last_log_term = LT},
             SlotOf) ->
    %% the token is a reference(); the engine only compares it for equality, so the batcher maps
    %% it to a 64-bit integer (ra_engine_batcher:token_id/2) and back
    ev(Row, ?EV_PRE_VOTE, SlotOf(Cand), 0, 0, 0, T, LI, LT, Tok, V bor (MV bsl 32), 0);
encode_event(Row, {Voter, #pre_vote_result{term = T, token = Tok, vote_granted = G}}, SlotOf) ->
    ev(Row, ?EV_PRE_VOTE_RES, SlotOf(Voter), 0, 0, 0, T, 0, 0, Tok, bool(G), 0);
encode_event(Row, #pre_vote_result{term = T, token = Tok, vote_granted = G}, _SlotOf) ->
    ev(Row, ?EV_PRE_VOTE_RES, ?NO_SLOT, 0, 0, 0, T, 0, 0, Tok, bool(G), 0);
encode_event(Row, #heartbeat_rpc{query_index = QI, term = T, leader_id = L}, SlotOf) ->
    ev(Row, ?EV_HEARTBEAT_RPC, SlotOf(L), 0, 0, 0, T, QI, 0, 0, 0, 0);
encode_event(Row, {Peer, #heartbeat_reply{query_index = QI, term = T}}, SlotOf) ->
    ev(Row, ?EV_HEARTBEAT_REPLY, SlotOf(Peer), 0, 0, 0, T, QI, 0, 0, 0, 0);
encode_event(Row, Msg, _SlotOf) ->
    %% host-origin events also have the 64-byte form (mixed batches use one record size)
    <<Row:32/little, Type, Flags, N:16/little, Term:64/little, A:64/little, B:64/little>> =
        encode_host_event(Row, Msg),
    ev(Row, Type, ?NO_SLOT, Flags, N, 0, Term, A, B, 0, 0, 0).

%% events only the host originates: 32-byte ra_host_event (ra_engine_step_host / _submit_host)
-spec encode_host_event(row(), term()) -> binary().
encode_host_event(Row, {ra_log_event, {written, Term, {From, To}}}) ->
    hev(Row, ?EV_WRITTEN, 0, 0, Term, From, To);
encode_host_event(Row, {command, N}) when is_integer(N) ->      %% {command, _} / {commands, [_]}
    hev(Row, ?EV_COMMAND, 0, N, 0, 0, 0);
encode_host_event(Row, {noop_command, N}) ->                    %% {command, {noop, _, _}}
    hev(Row, ?EV_COMMAND, ?EVF_NOOP, N, 0, 0, 0);
encode_host_event(Row, election_timeout) ->
    hev(Row, ?EV_ELECTION_TIMEOUT, 0, 0, 0, 0, 0);
encode_host_event(Row, await_condition_timeout) ->
    hev(Row, ?EV_AWAIT_COND_TIMEOUT, 0, 0, 0, 0, 0);
encode_host_event(Row, pipeline_rpcs) ->
    hev(Row, ?EV_PIPELINE_RPCS, 0, 0, 0, 0, 0);
encode_host_event(Row, tick) ->
    hev(Row, ?EV_TICK, 0, 0, 0, 0, 0);
encode_host_event(Row, consistent_query) ->
    hev(Row, ?EV_CONSISTENT_QUERY, 0, 0, 0, 0, 0).

ev(Row, Type, From, Flags, N, N1, Term, A, B, C, D, E) ->
    <<Row:32/little, Type, From, Flags, 0, N:16/little, N1:16/little, 0:32/little,
      Term:64/little, A:64/little, B:64/little, C:64/little, D:64/little, E:64/little>>.

hev(Row, Type, Flags, N, Term, A, B) ->
    <<Row:32/little, Type, Flags, N:16/little, Term:64/little, A:64/little, B:64/little>>.

bool(true) -> 1;
bool(false) -> 0.

%% [log_entry()] -> {N, N1, D, E}: N entries, the first N1 of term D and the rest of term E
%% (N1 = 0: all of term D).  The caller guarantees at most two runs.
-spec term_runs([log_entry()]) -> {non_neg_integer(), non_neg_integer(), ra_term(), ra_term()}.
term_runs([]) ->
    {0, 0, 0, 0};
term_runs([{_, T0, _} | _] = Es) ->
    {Same, Rest} = lists:splitwith(fun({_, T, _}) -> T =:= T0 end, Es),
    case Rest of
        [] -> {length(Es), 0, T0, 0};
        [{_, T1, _} | _] ->
            true = lists:all(fun({_, T, _}) -> T =:= T1 end, Rest),
            {length(Es), length(Same), T0, T1}
    end.

%% ---------------------------------------------------------------------------------------------
%% engine output: RPC records (already addressed: `row` = destination member) -> ra_msg()
%% ---------------------------------------------------------------------------------------------
%% IdOf :: fun((Row, Slot) -> ra_server_id()) maps a slot of the destination's group back to the
%% member id.  AER records carry index ranges, not payloads: the caller reads entries
%% PrevIdx+1 .. PrevIdx+N from its ra_log (ra_log:fold, as make_append_entries_rpc/6 does at
%% src/ra_server.erl:2401-2418) and fills `entries`.
-spec decode_record(binary(), fun((row(), slot()) -> ra_server_id())) ->
    {row(), FromSlot :: slot() | 255, Seq :: non_neg_integer(), Flags :: non_neg_integer(), term()}.
decode_record(<<Row:32/little, Type, From, Flags, _Pad, N:16/little, N1:16/little, Seq:32/little,
                Term:64/little, A:64/little, B:64/little, C:64/little, D:64/little,
                E:64/little>>, IdOf) ->
    Msg = case Type of
              ?EV_AER ->
                  {aer, #append_entries_rpc{term = Term, leader_id = IdOf(Row, From),
                                            leader_commit = C, prev_log_index = A,
                                            prev_log_term = B, entries = []},
                   {A + 1, A + N}, {N1, D, E}};
              ?EV_AER_REPLY ->
                  {IdOf(Row, From),
                   #append_entries_reply{term = Term, success = D =/= 0, next_index = A,
                                         last_index = B, last_term = C}};
              ?EV_REQUEST_VOTE ->
                  #request_vote_rpc{term = Term, candidate_id = IdOf(Row, From),
                                    last_log_index = A, last_log_term = B};
              ?EV_REQUEST_VOTE_RES ->
                  #request_vote_result{term = Term, vote_granted = D =/= 0};
              ?EV_PRE_VOTE ->
                  #pre_vote_rpc{version = D band 16#FFFFFFFF, machine_version = D bsr 32,
                                term = Term, token = C, candidate_id = IdOf(Row, From),
                                last_log_index = A, last_log_term = B};
              ?EV_PRE_VOTE_RES ->
                  #pre_vote_result{term = Term, token = C, vote_granted = D =/= 0};
              ?EV_HEARTBEAT_RPC ->
                  #heartbeat_rpc{query_index = A, term = Term, leader_id = IdOf(Row, From)};
              ?EV_HEARTBEAT_REPLY ->
                  {IdOf(Row, From), #heartbeat_reply{query_index = A, term = Term}};
              ?EV_WRITTEN -> {ra_log_event, {written, Term, {A, B}}};
              ?EV_ELECTION_TIMEOUT -> election_timeout;
              ?EV_AWAIT_COND_TIMEOUT -> await_condition_timeout;
              ?EV_CONSISTENT_QUERY -> consistent_query;
              ?EV_PIPELINE_RPCS -> pipeline_rpcs;          %% pure mode: {next_event, info, pipeline_rpcs}
              ?EV_COMMAND -> {command, N, Flags band ?EVF_NOOP =/= 0};
              ?EV_TICK -> tick
          end,
    {Row, From, Seq, Flags, Msg}.

-spec decode_records(binary(), fun((row(), slot()) -> ra_server_id())) -> [tuple()].
decode_records(Bin, IdOf) ->
    [decode_record(R, IdOf) || <<R:64/binary>> <= Bin].

%% ---------------------------------------------------------------------------------------------
%% host notes (enum ra_note_type)
%% ---------------------------------------------------------------------------------------------
-spec decode_note(binary()) -> {row(), tuple(), Aux :: non_neg_integer()}.
decode_note(<<Row:32/little, Type, Slot, Aux:16/little, A:64/little, B:64/little, C:64/little>>) ->
    N = case Type of
            1 -> {wal_append, A, B, C};                 %% entries A..B (term C) now in the log view
            2 -> {truncate, A, B};                      %% ra_log:set_last_index(A), its term B
            3 -> {commit, A, B};                        %% commit_index A -> B ({aux, eval})
            4 -> {apply, A, B};                         %% run A..B through ra_machine:apply/3
            5 -> {status, #{flags => Aux, term => A,
                            voted_for => slot(B band 16#FF), leader => slot((B bsr 8) band 16#FF),
                            role_old => role((B bsr 16) band 16#FF), role => role((B bsr 24) band 16#FF),
                            fatal => C band 16#FF, unconsumed => (C bsr 8) band 16#FF}};
            6 -> {send_snapshot, Slot, B};
            7 -> {not_leader, A, slot(B)};
            8 -> {query_index, A, B};
            9 -> {query_agreed, A};
            10 -> query_apply;
            11 -> {cancel_snapshot_retry, Slot}
        end,
    {Row, N, Aux}.

-spec decode_notes(binary()) -> [{row(), tuple(), non_neg_integer()}].
decode_notes(Bin) ->
    [decode_note(N) || <<N:32/binary>> <= Bin].

slot(16#FF) -> undefined;
slot(S) -> S.

role(0) -> follower;
role(1) -> candidate;
role(2) -> pre_vote;
role(3) -> leader;
role(4) -> await_condition.

role_code(follower) -> 0;
role_code(candidate) -> 1;
role_code(pre_vote) -> 2;
role_code(leader) -> 3;
role_code(await_condition) -> 4.

%% ---------------------------------------------------------------------------------------------
%% ra_row_state (544 bytes): what ra_server:init/1 (src/ra_server.erl:434-457) + the log facade hold
%% ---------------------------------------------------------------------------------------------
%% Row :: #{row, role, self_slot, n_members, leader_slot, voted_for, membership, condition,
%%          votes, machine_version, effective_machine_version, flags, current_term, commit_index,
%%          last_applied, pre_vote_token, token_counter, first_index, last_index, last_term,
%%          last_written :: {Idx, Term}, snapshot :: undefined | {Idx, Term},
%%          runs :: [{StartIdx, Term}], cond_reply :: {Term, Next, LastIdx, LastTerm},
%%          peers :: [#{next_index, match_index, commit_index_sent, status, voter}]}  (by slot)
-spec encode_row(map()) -> binary().
encode_row(#{row := Row, role := Role, self_slot := Self, n_members := M} = R) ->
    Leader = slot_code(maps:get(leader_slot, R, undefined)),
    Voted = slot_code(maps:get(voted_for, R, undefined)),
    {SnapIdx, SnapTerm, HasSnap} = case maps:get(snapshot, R, undefined) of
                                       undefined -> {0, 0, 0};
                                       {SI, ST} -> {SI, ST, 1}
                                   end,
    Runs = maps:get(runs, R, []),
    NRuns = length(Runs),
    true = NRuns =< 8,
    Pad = lists:duplicate(8 - NRuns, 0),
    Starts = [S || {S, _} <- Runs] ++ Pad,
    Terms = [T || {_, T} <- Runs] ++ Pad,
    {LwI, LwT} = maps:get(last_written, R, {0, 0}),
    {CT, CN, CL, CLT} = maps:get(cond_reply, R, {0, 0, 0, 0}),
    Peers0 = maps:get(peers, R, []),
    Peers = Peers0 ++ lists:duplicate(8 - length(Peers0), #{}),
    iolist_to_binary(
      [<<Row:32/little, (role_code(Role)), Self, M, Leader, Voted,
         (membership_code(maps:get(membership, R, voter))), (maps:get(condition, R, 0)), HasSnap,
         (maps:get(votes, R, 0)):32/little, (maps:get(machine_version, R, 0)):32/little,
         (maps:get(effective_machine_version, R, 0)):32/little, NRuns:32/little,
         (maps:get(flags, R, 0)):32/little,
         (maps:get(current_term, R, 0)):64/little, (maps:get(commit_index, R, 0)):64/little,
         (maps:get(last_applied, R, 0)):64/little, (maps:get(pre_vote_token, R, 0)):64/little,
         (maps:get(token_counter, R, 0)):64/little, (maps:get(first_index, R, 0)):64/little,
         (maps:get(last_index, R, 0)):64/little, (maps:get(last_term, R, 0)):64/little,
         LwI:64/little, LwT:64/little, SnapIdx:64/little, SnapTerm:64/little>>,
       [<<S:64/little>> || S <- Starts], [<<T:64/little>> || T <- Terms],
       <<CT:64/little, CN:64/little, CL:64/little, CLT:64/little>>,
       [encode_peer(P) || P <- Peers]]).

encode_peer(P) ->
    <<(maps:get(next_index, P, 0)):64/little, (maps:get(match_index, P, 0)):64/little,
      (maps:get(commit_index_sent, P, 0)):64/little,
      (peer_status_code(maps:get(status, P, normal))), (bool(maps:get(voter, P, false))), 0:48>>.

-spec decode_row(binary()) -> map().
decode_row(<<Row:32/little, Role, Self, M, Leader, Voted, Membership, Cond, HasSnap,
             Votes:32/little, MacVer:32/little, EffMacVer:32/little, NRuns:32/little,
             Flags:32/little,
             CurTerm:64/little, Commit:64/little, Applied:64/little, Tok:64/little,
             TokCtr:64/little, First:64/little, Last:64/little, LastTerm:64/little,
             LwI:64/little, LwT:64/little, SnapIdx:64/little, SnapTerm:64/little,
             StartsBin:64/binary, TermsBin:64/binary,
             CT:64/little, CN:64/little, CL:64/little, CLT:64/little, PeersBin:256/binary>>) ->
    Starts = [S || <<S:64/little>> <= StartsBin],
    Terms = [T || <<T:64/little>> <= TermsBin],
    Runs = lists:sublist(lists:zip(Starts, Terms), NRuns),
    Peers = [#{next_index => NI, match_index => MI, commit_index_sent => CS,
               status => peer_status(St), voter => V =/= 0}
             || <<NI:64/little, MI:64/little, CS:64/little, St, V, _:48>> <= PeersBin],
    #{row => Row, role => role(Role), self_slot => Self, n_members => M,
      leader_slot => slot(Leader), voted_for => slot(Voted), membership => membership(Membership),
      condition => Cond, votes => Votes, machine_version => MacVer,
      effective_machine_version => EffMacVer, flags => Flags,
      current_term => CurTerm, commit_index => Commit, last_applied => Applied,
      pre_vote_token => Tok, token_counter => TokCtr,
      first_index => First, last_index => Last, last_term => LastTerm,
      last_written => {LwI, LwT},
      snapshot => case HasSnap of 0 -> undefined; _ -> {SnapIdx, SnapTerm} end,
      runs => Runs, cond_reply => {CT, CN, CL, CLT},
      peers => lists:sublist(Peers, M)}.

%% argument of ra_engine_nif:read_rows/2
-spec row_ids([row()]) -> binary().
row_ids(Rows) ->
    << <<R:32/little>> || R <- Rows >>.

slot_code(undefined) -> ?NO_SLOT;
slot_code(S) when is_integer(S) -> S.

membership_code(voter) -> 0;
membership_code(promotable) -> 1;
membership_code(non_voter) -> 2;
membership_code(unknown) -> 3.

membership(0) -> voter;
membership(1) -> promotable;
membership(2) -> non_voter;
membership(3) -> unknown.

peer_status_code(normal) -> 0;
peer_status_code({sending_snapshot, _}) -> 1;
peer_status_code(sending_snapshot) -> 1;
peer_status_code({snapshot_backoff, _}) -> 2;
peer_status_code(snapshot_backoff) -> 2;
peer_status_code(suspended) -> 3;
peer_status_code(disconnected) -> 4.

peer_status(0) -> normal;
peer_status(1) -> sending_snapshot;
peer_status(2) -> snapshot_backoff;
peer_status(3) -> suspended;
peer_status(4) -> disconnected.
