#!/bin/bash
# final multi-GPU evidence on one 8-GPU box: the scaling points N = 1, 2, 4, 8 of the shipped build (short), the
# all-to-all transport at N = 8 (bucket capacity), and the full bench line at N = 8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P='import sys,json; d=json.loads(sys.stdin.readline()); print("N %d ms_per_step %.4f value %.1fM parity %s/%s dropped %s" % (d["n_gpus"], d["ms_per_step"], d["value"]/1e6, d["parity"]["rows_bad"], d["parity"]["rows_checked"], d["run"]["msgs_dropped"]))'
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --no-extra-configs --no-latency 2>&1 | tail -1 | python -c "$P"
for N in 2 4 8; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29520+N)) bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-e2e --no-extra-configs 2>&1 | tail -1 | python -c "$P" || echo "N $N FAILED"
done
echo "== a2a N=8"
timeout 600 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu --no-e2e --no-extra-configs --transport a2a 2>&1 | tail -1 | python -c "$P" || echo FAILED
echo "== full bench N=8"
timeout 900 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; tail -1 gpurun_out/r2_bench_n8_final.json | python -c "$P"
echo "== reference arm N=8"
timeout 300 $TR --nproc-per-node 8 --master-port 29542 bench.py --impl reference --gpus 8 --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
