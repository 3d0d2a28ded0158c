set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_planned_encode.py tests/test_gpu_hostapi.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -6
python tools/bench_kernels.py --size-mib 1024 --radices 3,4 --hist-variants 0 2>&1 | grep -i "decode\|encode" | cut -c1-600
ncu --set full --clock-control none --import-source on -k regex:'fsm_sync|fsm_write|encode_fast|hist_runs' -s 5 -c 5 \
    -o gpurun_out/prof_r2e python tools/profile_target.py --size-mib 256 > gpurun_out/ncu_r2e.log 2>&1
tail -3 gpurun_out/ncu_r2e.log
