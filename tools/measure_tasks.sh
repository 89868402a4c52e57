#!/bin/bash
# 1-GPU bench lines of the other BASELINE configs (round-end measurement): tools/measure_tasks.sh <tag> [tasks...]
T=${1:-tasks}; shift; O=gpurun_out/$T; mkdir -p $O
for t in ${@:-orient search insert tool_grasp tool_orient}; do
  timeout 240 python bench.py --task $t --steps 64 --warmup 8 --no-cpu-baseline --no-sleep-off > $O/bench_$t.json 2> $O/bench_$t.err
  python -c "
import json,sys
try:
    d=json.load(open('$O/bench_$t.json')); print('$t', round(d['value']), 'env-steps/s, rollout', round(d['rollout_only']['value']), 'k_simulate ms', round(d['roofline']['ms_per_launch'],3), 'contacts', round(d['contacts_per_env']['mean'],1))
except Exception as e: print('$t FAILED', e)"
done
