"""CPU, world_size 2 (gloo): the N>1 host logic — the tensor-parallel shard plan
(llama2.zig_b200/tp_plan.py, the same windows csrc/llama2_b200.cu uploads) with one all-reduce
of the hidden vector after wo and after w2 reproduces the unsharded oracle, and the rendezvous
token broadcast bench.py performs works over torch.distributed."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPE = (64, 176, 2, 8, 4, -96, 24)   # GQA, unshared classifier
STEPS = 12


def _worker(rank, world, port, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import llama2_zig_b200 as l2b
    import oracle_lib as O
    from llama2_zig_b200.checkpoint import shape_checkpoint
    from llama2_zig_b200 import tp_plan

    # rendezvous token: rank 0 makes it, everyone receives it (bench.py does this with NCCL's id)
    token = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        token.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(token, 0)
    assert token.tolist() == list(range(128))

    ck = shape_checkpoint(SHAPE)
    ck.data = l2b.synth_checkpoint_host(ck, 3)
    full = tp_plan.payload_views(ck)
    plan = tp_plan.shard_plan(ck, rank, world)
    dim, hid, L, H, KV, V, S = ck.shape_tuple
    hs = dim // H
    kv_mul = H // KV
    loc = {k: (full[k][..., s.row0:s.row1, s.col0:s.col1] if full[k].ndim == 3 else full[k])
           for k, s in plan.items()}
    wcls_loc = full["wcls"][plan["wcls"].row0:plan["wcls"].row1]
    h_loc, kv_loc = H // world, KV // world
    kc = np.zeros((L, S, kv_loc * hs), np.float32)
    vc = np.zeros((L, S, kv_loc * hs), np.float32)
    lib = O.load("strict")
    import ctypes as C
    oracle = O.OracleModel(O.make_config(*ck.shape_tuple), ck.data, ck.shared_weights, W=8, kind="strict") if rank == 0 else None

    def rms(x, w):
        return (x * np.float32(1.0 / np.sqrt(np.mean(x.astype(np.float64) ** 2) + 1e-5))).astype(np.float32) * w

    def allreduce(v):
        t = torch.from_numpy(v.copy())
        dist.all_reduce(t)
        return t.numpy()

    worst = 0.0
    for pos in range(STEPS):
        tok = (1 + 7919 * pos) % V
        x = full["token_embedding_table"][tok].copy()
        for l in range(L):
            xb = rms(x, full["rms_att_weight"][l])
            q = loc["wq"][l] @ xb
            k = loc["wk"][l] @ xb
            v = loc["wv"][l] @ xb
            a, b = C.c_float(), C.c_float()
            for i in range(0, q.size, 2):
                lib.orc_rope_angle(i, hs, pos, C.byref(a), C.byref(b))
                q[i], q[i + 1] = q[i] * a.value - q[i + 1] * b.value, q[i] * b.value + q[i + 1] * a.value
                if i < k.size:
                    k[i], k[i + 1] = k[i] * a.value - k[i + 1] * b.value, k[i] * b.value + k[i + 1] * a.value
            kc[l, pos], vc[l, pos] = k, v
            att_out = np.zeros(h_loc * hs, np.float32)
            for h in range(h_loc):
                kh = (h // kv_mul) * hs
                sc = kc[l, :pos + 1, kh:kh + hs] @ q[h * hs:(h + 1) * hs] / np.float32(np.sqrt(hs))
                p = np.exp(sc - sc.max())
                p /= p.sum()
                att_out[h * hs:(h + 1) * hs] = p @ vc[l, :pos + 1, kh:kh + hs]
            x = x + allreduce(loc["wo"][l] @ att_out)            # all-reduce #1 (after wo)
            xb = rms(x, full["rms_ffn_weight"][l])
            h1 = loc["w1"][l] @ xb
            h3 = loc["w3"][l] @ xb
            hb = h1 * (1.0 / (1.0 + np.exp(-h1))) * h3
            x = x + allreduce(loc["w2"][l] @ hb.astype(np.float32))   # all-reduce #2 (after w2)
        xn = rms(x, full["rms_final_weight"])
        logits_loc = (wcls_loc @ xn).astype(np.float32)
        gathered = [torch.zeros(V // world) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(logits_loc))
        logits = torch.cat(gathered).numpy()
        if rank == 0:
            want = oracle.forward(tok, pos)
            worst = max(worst, float(np.max(np.abs(logits - want)) / np.max(np.abs(want))))
    if rank == 0:
        result_q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


def test_tp_shard_plan_matches_oracle_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    worst = q.get(timeout=5)
    assert worst <= 1e-4, worst


def test_plan_partitions_every_sharded_tensor_exactly():
    sys.path.insert(0, ROOT)
    from llama2_zig_b200.checkpoint import shape_checkpoint
    from llama2_zig_b200 import tp_plan
    ck = shape_checkpoint("llama2-7B")
    assert tp_plan.weight_bytes_per_token(ck, 1) == 26_429_374_464
    for world in (2, 4, 8):
        plans = [tp_plan.shard_plan(ck, r, world) for r in range(world)]
        for name, full_rows, full_cols in (("wq", 4096, 4096), ("wk", 4096, 4096), ("wo", 4096, 4096),
                                           ("w1", 11008, 4096), ("w2", 4096, 11008), ("wcls", 32000, 4096)):
            covered = np.zeros((full_rows, full_cols), np.int8) if full_rows * full_cols < 5e7 else None
            rows = sorted((p[name].row0, p[name].row1) for p in plans)
            cols = sorted((p[name].col0, p[name].col1) for p in plans)
            if plans[0][name].col1 - plans[0][name].col0 == full_cols:      # row-sharded
                assert rows[0][0] == 0 and rows[-1][1] == full_rows
                assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            else:                                                           # column-sharded
                assert cols[0][0] == 0 and cols[-1][1] == full_cols
                assert all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
        total = sum(tp_plan.weight_bytes_per_token(ck, world) for _ in range(world))
        replicated = 4 * (32 * 2 * 4096 + 4096) * (world - 1)               # norms are replicated
        assert total == 26_429_374_464 + replicated
    with pytest.raises(ValueError):
        tp_plan.shard_plan(shape_checkpoint("stories15M"), 0, 4)            # 6 kv heads % 4 != 0


def _exchange_worker(rank, world, port, result_q):
    """The data flow of the fused GEMV + all-reduce with gloo as the wire: every rank's
    partial rows land in every rank's landing area [slot][source rank][dim] tagged with the step's
    epoch; slice owners fold the ranks' rows into x in RANK ORDER.  Checks what the CUDA path relies on:
    (1) every rank forms bit-identical x, (2) it equals the all-reduce within fp32 reordering noise,
    (3) a step that skips a reduce point (prefill skips the classifier's) leaves stale epochs that are
    never mistaken for this step's data."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    dim, slots = 256, 4
    landing = np.zeros((slots, world, dim), np.float32)
    flags = np.zeros((slots, world), np.int64)           # epoch carried by the units of (slot, source rank)
    x = np.ones(dim, np.float32)
    ok = True
    for epoch in range(1, 6):
        skip_last = (epoch == 3)                         # a prefill-style step: the last reduce point is not consumed
        for slot in range(slots):
            partial = rng.standard_normal(dim).astype(np.float32)
            gathered = [torch.zeros(dim) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(partial))        # "posted stores into every rank's landing area"
            for r in range(world):
                landing[slot, r] = gathered[r].numpy()
                flags[slot, r] = epoch
            if skip_last and slot == slots - 1:
                continue                                                # producer ran, nobody consumes this slot this step
            assert np.all(flags[slot] == epoch)                         # the consumer only accepts this step's units
            ref = torch.from_numpy(partial.copy())
            dist.all_reduce(ref)
            xn = x.copy()
            for r in range(world):                                      # fixed rank order (tp_reduce_tail)
                xn = (xn + landing[slot, r]).astype(np.float32)
            ok = ok and np.allclose(xn, x + ref.numpy(), rtol=0, atol=1e-5)
            x = xn
        # stale epochs of a skipped reduce point are older than the next step's epoch
        assert np.all(flags <= epoch)
    digest = torch.tensor(np.frombuffer(x.tobytes(), dtype=np.uint8).astype(np.int64).sum())
    both = [torch.zeros((), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(both, digest)
    bits = [torch.from_numpy(np.zeros(dim, np.float32)) for _ in range(world)]
    dist.all_gather(bits, torch.from_numpy(x))
    identical = all(np.array_equal(bits[0].numpy().view(np.uint32), b.numpy().view(np.uint32)) for b in bits)
    if rank == 0:
        result_q.put(bool(ok and identical))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_dataflow_is_rank_order_deterministic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29750 + (os.getpid() % 200)
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
