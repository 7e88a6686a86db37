#!/bin/bash
# bench several builds / flag sets of the engine in one GPU call:
#   tools/bench_variants.sh base "base --permille 0" D ...
for spec in "$@"; do
  set -- $spec
  v=$1; shift
  so=ra_b200/csrc/libra_engine_$v.so
  [ "$v" = "base" ] && so=ra_b200/csrc/libra_engine.so
  echo "== $v $*"
  RA_ENGINE_SO=$PWD/$so timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --no-parity --no-extra-configs --no-latency "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ms_per_step %.4f value %.1fM frac %.3f events/step %.0f'%(d['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['run']['events_per_step']))"
done
