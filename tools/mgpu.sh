#!/bin/bash
# multi-GPU checks in one call: tools/mgpu.sh <N>
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests/test_sharded_nccl.py -x -q -m gpu -k "$N-" 2>&1 | tail -6
for bar in device nccl; do
  echo "== bench --gpus $N barrier=$bar"
  RA_PEER_BARRIER=$bar timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-e2e --no-extra-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ms_per_step %.4f value %.1fM parity %s dropped %s' % (d['ms_per_step'], d['value']/1e6, d.get('parity'), d['run']['msgs_dropped']))" || echo FAILED
done
echo "== bench --gpus $N a2a"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu --no-e2e --no-extra-configs --transport a2a 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ms_per_step %.4f value %.1fM parity %s dropped %s' % (d['ms_per_step'], d['value']/1e6, d.get('parity'), d['run']['msgs_dropped']))" || echo FAILED
echo "== full bench --gpus $N (default barrier)"
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 3000 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
