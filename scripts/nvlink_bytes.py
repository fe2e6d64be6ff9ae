"""NVLink bytes per decoded token of the tensor-parallel llama2-7B step (evidence for the fused
GEMV + all-reduce data plane): `nvidia-smi nvlink -gt d` counters around N device loops.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/nvlink_bytes.py

Expected from the design, per rank and token, sent to EACH peer: 64 reduce points x 4096 rows
x 8 B (one LL unit per row) = 2.10 MB, plus 16 B of argmax keys in the device loop.
"""
import os
import re
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama2_zig_b200 as l2b
from llama2_zig_b200.checkpoint import shape_checkpoint


def counters(idx):
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(idx)], capture_output=True, text=True).stdout
    tx = sum(int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return tx * 1024, rx * 1024, out


rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(l2b.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
ck = shape_checkpoint("llama2-7B")
t = l2b.Transformer(ck, synthetic_seed=7, rank=rank, world_size=world, device=local, comm_id=bytes(idt.cpu().numpy().tobytes()))
forced = np.array([(1 + 7919 * p) % ck.vocab_size for p in range(1, 257)], dtype=np.int32)
t.generate_argmax(1, 0, 256, forced=forced, stop_on_bos=False)     # warm-up
torch.cuda.synchronize(); dist.barrier()
tx0, rx0, raw0 = counters(local)
runs = 4
for _ in range(runs):
    t.reset()
    t.generate_argmax(1, 0, 256, forced=forced, stop_on_bos=False)
torch.cuda.synchronize(); dist.barrier()
tx1, rx1, raw1 = counters(local)
tokens = runs * 256
if rank == 0:
    expect = 64 * 4096 * 8 * (world - 1)
    print(f"world {world}: rank 0 NVLink Tx {(tx1 - tx0) / tokens / 1e6:.3f} MB/token, Rx {(rx1 - rx0) / tokens / 1e6:.3f} MB/token "
          f"(design: {expect / 1e6:.3f} MB/token of LL units per direction; counters are per-link data KiB summed over links)")
    if tx1 == tx0:
        print("counters did not move; raw output follows"); print(raw1[:1500])
t.close()
dist.barrier()
dist.destroy_process_group()
