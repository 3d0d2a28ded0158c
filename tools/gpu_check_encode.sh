set -e
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encode or degenerate or wide or shard or roundtrip" 2>&1 | tail -5
python tools/bench_kernels.py --size-mib 1024 --radices 2,4,16 --hist-variants 0 2>&1 | grep -v "histogram\|nybble\|torch" | cut -c1-400
