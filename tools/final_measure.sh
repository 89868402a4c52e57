#!/bin/bash
# Round-end measurement pass on ONE B200 (run through gpurun): GPU test suite, smoke, both bench arms, the ncu launch list of the bench
# command and one ncu --set full capture of k_simulate at 16 384 envs.  Outputs under gpurun_out/<tag>/.
T=${1:-r2final}; O=gpurun_out/$T; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.json
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 300 $O/bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 16 --warmup 8 --no-cpu-baseline --no-sleep-off --e2e-steps 1 > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simulate -s 300 -c 2 -f -o $O/sim \
    python bench.py --mode rollout --steps 4 --warmup 3 --no-cpu-baseline --no-sleep-off --e2e-steps 1 > $O/ncu_sim.log 2>&1
ls -la $O
