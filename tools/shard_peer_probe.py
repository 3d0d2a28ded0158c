#!/usr/bin/env python
"""Two (or more) ranks, one sharded encode: does the shard layer's peer-memory exchange come up on this box?
    DC_SHARD_DEBUG=1 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_peer_probe.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import data_compression_b200 as dc
from data_compression_b200 import shard
S = shard.NcclShards(torch.device("cuda", rank))
x = torch.randint(1, 200, (1 << 20,), dtype=torch.uint8, device="cuda")
for _ in range(3):
    buf = S.encode(x, 4)
torch.cuda.synchronize()
print("rank", rank, "peer exchange active:", dc.lib().dc_debug_shard_peer_active(S.comm), "info", S.encode_info(buf, x.numel()), flush=True)
S.close()
dist.destroy_process_group()
