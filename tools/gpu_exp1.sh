set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostapi.py tests/test_gpu_shard.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -4
python tools/bench_kernels.py --size-mib 1024 --radices 3,4,16 --hist-variants 0 2>&1 | grep '"decode\[' | cut -c1-520
