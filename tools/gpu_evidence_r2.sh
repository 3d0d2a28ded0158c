#!/bin/bash
# Round-2 evidence run for profiles/: the GPU suite, the bench (both arms), per-kernel timings for every radix, then
# (separately, never as a bench value) the ncu launch list of the bench command and one full capture per hot kernel at
# the bench's size.
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || tail -5 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_n1_reference.json 2> gpurun_out/r2_bench_n1_reference.err || tail -5 gpurun_out/r2_bench_n1_reference.err
python tools/bench_kernels.py --size-mib 1024 --radices 2,3,4,16 --hist-variants 0,1 > gpurun_out/r2_kernel_timings_1GiB.jsonl 2> gpurun_out/r2_kernel_timings.err || tail -5 gpurun_out/r2_kernel_timings.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fsm_sync|fsm_write|encode_fast|hist_runs|encode_plan|table_kernel' -s 6 -c 6 \
    -o gpurun_out/prof_r2_final python tools/profile_target.py --size-mib 1024 > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
