#!/bin/bash
# A/B of several builds of the library on one box: rollout-only bench (k_simulate is 95 % of it) with each build_ab/lib_*.so
# AB_ARGS: extra bench.py arguments (e.g. "--edge-contacts 0")
for so in "$@"; do
  n=$(basename $so .so)
  SEQDEX_B200_LIB=$PWD/$so python bench.py --mode rollout --no-cpu-baseline --no-sleep-off --e2e-steps 2 --steps 96 --warmup 8 $AB_ARGS 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', '$AB_ARGS', round(d['value']), 'env-steps/s', round(d['roofline']['ms_per_launch'], 4), 'ms per k_simulate launch, contacts', round(d['contacts_per_env']['mean'], 1), 'asleep', round(d['bricks_asleep_frac'], 3))"
done
