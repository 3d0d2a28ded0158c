set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_hostapi.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python tools/fuzz.py 40 2>&1 | tail -2
python tools/bench_kernels.py --size-mib 256 --radices 2,3,4,16 --hist-variants 0 2>&1 | grep '"table\[' | cut -c1-200
