run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-engines $K 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('e2e %.1fM (%.3f ms/step, call %.3f model %.3f)'%(d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['engine_call_ms_per_step'], d['e2e']['host_model_ms_per_step']))"; }
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
K=1 run RA_HOSTSIM_THREADS=16
K=1 run RA_HOSTSIM_THREADS=32
K=2 run RA_HOSTSIM_THREADS=8 OMP_WAIT_POLICY=passive
K=2 run RA_HOSTSIM_THREADS=16 OMP_WAIT_POLICY=passive
K=2 run RA_HOSTSIM_THREADS=16 OMP_PROC_BIND=false GOMP_SPINCOUNT=0
K=3 run RA_HOSTSIM_THREADS=8 OMP_WAIT_POLICY=passive
