#!/bin/bash
# Evidence run for profiles/: the bench, then (separately, never as a bench value) the ncu launch list of the same
# command and one full capture per hot kernel at the bench's size.
set -x
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || tail -5 gpurun_out/bench_final.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err || tail -5 gpurun_out/bench_final_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'encode_single|decode_fast_sync|decode_fast_write|hist_warp' -s 5 -c 5 \
    -o gpurun_out/prof_final python tools/profile_target.py --size-mib 1024 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
