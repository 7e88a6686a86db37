#!/bin/bash
# compute-sanitizer over small engine runs (GPU box): memcheck on the host-facing calls and a sharded peer-store
# flood, racecheck (shared-memory hazards: the TMA ring, the gather table) on a flood.  Logs -> gpurun_out/.
cat > /tmp/san_flood.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from ra_b200 import abi
from ra_b200.engine import Engine, HostFlood
from ra_b200.sharded import LocalPeerTransport, Shard, ShardedFlood
e = Engine(256, 5, route_on_device=True); e.reset_empty()
e.step([abi.ev_simple(e.row_of(g, 0), abi.EV_ELECTION_TIMEOUT) for g in range(256)])
e.flood(30, 1, 20, seed=3); e.flood(20, 2, 20, seed=3, faults=(20, 50, 30, 8))
print("flood", e.counters()["commits"])
b = Engine(128, 5, route_on_device=True); b.reset_empty()
hf = HostFlood([b]); st = hf.run(20, 1, 10, seed=5, bootstrap=True); print("hostsim", b.counters()["commits"], st["d2h_bytes"])
shards = [Shard(64, 5, 4, k, buckets=False) for k in range(4)]
fl = ShardedFlood(LocalPeerTransport(shards)); fl.bootstrap(); fl.run(20, 1, 10, seed=7); fl.sync(); print("peer", fl.counters()["commits"])
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_flood.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
