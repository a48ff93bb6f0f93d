#!/bin/bash
# Single-GPU measurement set for profiles/ (run under gpurun):  bash scripts/profile_round.sh <tag>
# bench lines are taken OUTSIDE any profiler; ncu passes follow the recipe in /opt/skills/guides/B200_PROFILING.md.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_qft32.json 2> $out/${tag}_bench_qft32.err
python bench.py --steps 5 --warmup 3 --nqubits 30 --no-cpu > $out/${tag}_bench_qft30.json 2> $out/${tag}_bench_qft30.err
python bench.py --impl reference --steps 1 --warmup 0 --cpu-budget 10 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
python scripts/configs_bench.py > $out/${tag}_configs.log 2>&1
cp $out/configs_bench.json $out/${tag}_configs3_4.json
python scripts/k8_probe.py > $out/${tag}_k8_probe.log 2>&1
# launch list (device time per launch, cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $out/${tag}_ncu_launches_qft30.csv python bench.py --steps 2 --warmup 1 --nqubits 30 --no-cpu > $out/${tag}_ncu_launches.log 2>&1
# DRAM traffic per launch at the headline size
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep_kernel|k8_permute' -c 5 --csv \
    --log-file $out/${tag}_ncu_dram_traffic_qft32.csv python scripts/prof_case.py qftswap 32 > $out/${tag}_ncu_traffic.log 2>&1
# the sweep kernel in full (first and last sweep of QFT(30))
REPS=1 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 4 -o /tmp/prof_full python scripts/prof_case.py qft 30 > $out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > $out/${tag}_ncu_full_qft30_sweep_kernel_raw.csv
ncu -i /tmp/prof_full.ncu-rep --page details --csv > $out/${tag}_ncu_full_qft30_sweep_kernel_details.csv
ncu -i /tmp/prof_full.ncu-rep --page source --csv --kernel-id :::1 > $out/${tag}_ncu_full_qft30_sweep0_source.csv 2>/dev/null
tail -1 $out/${tag}_bench_qft32.json
tail -1 $out/${tag}_bench_qft30.json
tail -1 $out/${tag}_bench_reference.json
tail -3 $out/${tag}_configs.log
cat $out/${tag}_k8_probe.log
