"""Tensor-parallel parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/tp_check.py [--shape dim,hidden,L,H,KV,V,S] [--steps K]

Every rank builds its shard (l2b_create_sharded on the host payload, and l2b_create_synthetic
for the device-generated variant); rank 0 also builds the single-GPU context and the CPU oracle.
Checks per position: TP logits vs 1-GPU logits and vs the oracle <= 1e-4 relative
(BASELINE.json config 5), argmax path equal, synthetic-device weights == host-payload weights.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,1376,3,8,8,-1024,96")
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--dead-peer", action="store_true",
                    help="rank 1 stops stepping: rank 0 must get L2B_ERR_COMM (-6), not hang")
    args = ap.parse_args()
    import llama2_zig_b200 as l2b
    from llama2_zig_b200.checkpoint import shape_checkpoint

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    shape = tuple(int(v) for v in args.shape.split(","))
    ck = shape_checkpoint(shape)
    ck.data = l2b.synth_checkpoint_host(ck, args.seed)

    def fresh_id():
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(l2b.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    tp = l2b.Transformer(ck, rank=rank, world_size=world, device=local, comm_id=fresh_id())
    if args.dead_peer:
        tp.forward(1, 0)                       # one good step on every rank
        dist.barrier()
        status = 0
        if rank == 0:
            try:
                tp.forward(2, 1)               # rank 1 never makes this call
            except l2b.L2BError as e:
                status = e.status
            print("DEAD_PEER_OK" if status == -6 else f"DEAD_PEER_FAIL status={status}", flush=True)
        dist.barrier()
        tp.close()
        dist.destroy_process_group()
        sys.exit(0 if (rank != 0 or status == -6) else 1)
    tp_syn = l2b.Transformer(shape_checkpoint(shape), synthetic_seed=args.seed, rank=rank, world_size=world,
                             device=local, comm_id=fresh_id())
    single = oracle = None
    if rank == 0:
        import oracle_lib as O
        single = l2b.Transformer(ck, device=local)
        oracle = O.OracleModel(O.make_config(*ck.shape_tuple), ck.data, ck.shared_weights, W=8, kind="strict")
    toks = [(1 + 7919 * p) % ck.vocab_size for p in range(args.steps)]
    worst_single = worst_oracle = 0.0
    ok = True
    for pos, tok in enumerate(toks):
        got = tp.forward(tok, pos)
        got_syn = tp_syn.forward(tok, pos)
        nxt = tp.forward_argmax(tok, pos)
        if not np.array_equal(got, got_syn):
            ok = False
            print(f"rank {rank}: device-synthetic shard != host-payload shard at pos {pos}", flush=True)
        if rank == 0:
            ref = single.forward(tok, pos)
            want = oracle.forward(tok, pos)
            scale = float(np.max(np.abs(want)))
            worst_single = max(worst_single, float(np.max(np.abs(got - ref))) / scale)
            worst_oracle = max(worst_oracle, float(np.max(np.abs(got - want))) / scale)
            if nxt != int(np.argmax(got)):
                ok = False
                print(f"device argmax {nxt} != argmax of gathered logits {int(np.argmax(got))} at pos {pos}", flush=True)
    # the on-device generation loop (argmax exchanged between ranks inside advance_kernel)
    tp.reset()
    gen = tp.generate_argmax(1, 0, min(16, ck.seq_len), stop_on_bos=False)
    if rank == 0:
        single.reset()
        gen1 = single.generate_argmax(1, 0, min(16, ck.seq_len), stop_on_bos=False)
        if gen.tolist() != gen1.tolist():
            ok = False
            print(f"generate_argmax differs: tp {gen.tolist()} vs 1 GPU {gen1.tolist()}", flush=True)
    # temperature sampling preparation on the device under TP (logits gathered from all ranks)
    tp.reset()
    probs, cand = tp.forward_sample(1, 0, 0.8, 0.9)
    if rank == 0:
        single.reset()
        p1, c1 = single.forward_sample(1, 0, 0.8, 0.9)
        if float(np.max(np.abs(probs - p1))) > 1e-4 * float(np.max(p1)) or (cand is None) != (c1 is None):
            ok = False
            print("forward_sample differs between tp and 1 GPU", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        good = flag.item() == 0 and worst_single <= 1e-4 and worst_oracle <= 1e-4
        print(f"TP_CHECK world={world} shape={shape} steps={args.steps} max_rel_vs_1gpu={worst_single:.3e} "
              f"max_rel_vs_oracle={worst_oracle:.3e} {'PASS' if good else 'FAIL'}", flush=True)
    for t in (tp, tp_syn, single):
        if t is not None:
            t.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 and worst_single <= 1e-4 and worst_oracle <= 1e-4 else 1)


if __name__ == "__main__":
    main()
