set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostapi.py tests/test_gpu_shard.py tests/test_gpu_shard_nccl.py tests/test_gpu_fuzz.py tests/test_gpu_planned_encode.py -x -q -m gpu 2>&1 | tail -4
timeout 100 python tools/fuzz.py 40 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --e2e-steps 0 > gpurun_out/bench_r2h_n2.json 2> gpurun_out/bench_r2h_n2.err; tail -1 gpurun_out/bench_r2h_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2h_n2.json"))
b=d["breakdown"]
print(d["config"].get("shard_exchange"), d["ms_per_step"], b["encode_ms"], b["decode_ms"], {k:round(v["ms_avg"],4) for k,v in b["kernels"].items()})
PY
