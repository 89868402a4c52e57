#!/bin/bash
# Round-end measurement pass on ONE B200 (run through gpurun): GPU test suite, both bench arms, the ncu launch list of the bench
# command and one ncu --set full capture of k_simulate at 16 384 envs.  Outputs under gpurun_out/.
O=gpurun_out; T=${1:-final3}
timeout 600 python -m pytest tests -m gpu -q > $O/t_$T.log 2>&1; tail -3 $O/t_$T.log
timeout 600 python bench.py > $O/bench_${T}_1gpu.json 2> $O/bench_${T}_1gpu.err; tail -c 400 $O/bench_${T}_1gpu.json
timeout 600 python bench.py --impl reference > $O/bench_${T}_reference_arm.json 2> $O/bench_${T}_reference_arm.err; tail -c 300 $O/bench_${T}_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$T.csv \
    python bench.py --steps 8 --warmup 8 --no-cpu-baseline --e2e-steps 1 > $O/launches_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simulate --launch-skip 335 --launch-count 1 -f -o $O/prof_$T \
    python bench.py --mode rollout --steps 8 --warmup 8 --no-cpu-baseline --e2e-steps 1 > $O/prof_$T.log 2>&1
ls -la $O/prof_$T.ncu-rep
