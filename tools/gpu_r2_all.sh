set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/bench_kernels.py --size-mib 1024 --radices 2,3,4,16 --hist-variants 0,1 2>&1 | grep -v "nybble_text\|adaptive\|copy" | cut -c1-500
