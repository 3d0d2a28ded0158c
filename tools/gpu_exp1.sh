set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/bench_kernels.py --size-mib 1024 --radices 2,3,4 --hist-variants 0 2>&1 | grep '"decode\[\|encode' | cut -c1-520
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r2b.err | tee gpurun_out/bench_r2b.json | cut -c1-1500
