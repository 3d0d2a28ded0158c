set -x
DC_DECODE_TMA=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostapi.py tests/test_gpu_fuzz.py tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -3
for t in 0 1; do DC_DECODE_TMA=$t python tools/bench_kernels.py --size-mib 1024 --radices 4,16 --hist-variants 0 2>&1 | grep '"decode\[' | grep -o '"kernels_ms.*'; done
