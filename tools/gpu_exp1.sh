set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostapi.py tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -5
timeout 200 python tools/fuzz.py 60 2>&1 | tail -3
python tools/bench_kernels.py --size-mib 1024 --radices 2 --hist-variants 0 2>&1 | grep '"decode\[' | cut -c1-520
