%% ra_log_meta -- B1 harness: term / voted_for / last_applied in the process dictionary, exactly what
%% ra_server_SUITE:setup_log/0 mocks (test/ra_server_SUITE.erl:176-187).  SOURCE ONLY in the build image.
-module(ra_log_meta).
-export([store/4, store_sync/4, fetch/3, fetch/4, delete/2, delete_sync/2]).
store(_, U, K, V) -> put({U, K}, V), ok.
store_sync(_, U, K, V) -> put({U, K}, V), ok.
fetch(_, U, K) -> get({U, K}).
fetch(_, U, K, D) -> case get({U, K}) of undefined -> D; V -> V end.
delete(_, _) -> ok.
delete_sync(_, _) -> ok.
