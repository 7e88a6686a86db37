#!/bin/bash
# A/B of engine builds / switches in one GPU call.  Each argument: "<label> <so-variant|base> [ENV=val ...] -- [bench flags]"
for spec in "$@"; do
  label=${spec%% *}; rest=${spec#* }
  v=${rest%% *}; rest=${rest#* }
  envs=${rest%%--*}; flags=${rest#*--}
  so=ra_b200/csrc/libra_engine_$v.so; [ "$v" = "base" ] && so=ra_b200/csrc/libra_engine.so
  echo -n "== $label [$v $envs --$flags] "
  env $envs RA_ENGINE_SO=$PWD/$so timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --no-parity --no-extra-configs --no-latency $flags 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ms_per_step %.4f value %.1fM frac %.3f events/step %.0f'%(d['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['run']['events_per_step']))" || echo FAILED
done
