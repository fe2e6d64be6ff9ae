"""GPU: transformer() parity through the C ABI (l2b_forward & friends) against the oracle.

Tolerances are the north star's: tokens identical at temperature 0, per-step logits within
1e-4 relative (measured as max|diff| / max|logit| per step)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-4


def rel_err(got, want):
    return float(np.max(np.abs(got.astype(np.float64) - want)) / (np.max(np.abs(want)) + 1e-30))


def teacher_tokens(n, vocab):
    return [(1 + 7919 * p) % vocab for p in range(n)]


def make_pair(l2b, oracle, shape, seed):
    from llama2_zig_b200.checkpoint import shape_checkpoint
    ck = shape_checkpoint(shape)
    ck.data = l2b.synth_checkpoint_host(ck, seed)
    om = oracle.OracleModel(oracle.make_config(*ck.shape_tuple), ck.data, ck.shared_weights, W=8, kind="strict")
    return ck, om


def test_stories15m_logits_and_tokens_match_oracle(l2b, oracle, stories15m):
    """configs[1]: stories15M.bin fp32 on 1xB200, -t 0, 256-token decode, bit-exact tokens."""
    with open(os.path.join(GOLDEN, "stories15M_t0_tokens.json")) as f:
        gold = json.load(f)
    ck = l2b.read_checkpoint(stories15m, mmap=False)
    cfg, shared, data = oracle.read_checkpoint(stories15m)
    om = oracle.OracleModel(cfg, data, shared, W=8, kind="strict")
    worst = 0.0
    with l2b.Transformer(ck) as t:
        token, toks = 1, []
        for pos in range(ck.seq_len):
            got = t.forward(token, pos)
            want = om.forward(token, pos)
            worst = max(worst, rel_err(got, want))
            nxt = int(np.argmax(got))
            assert nxt == int(np.argmax(want)), f"argmax differs at pos {pos}"
            if nxt == 1:
                break
            toks.append(nxt)
            token = nxt
    assert worst <= REL_TOL, worst
    assert toks == gold["tokens"]
    assert hashlib.sha256(np.array(toks, "<u4").tobytes()).hexdigest() == gold["sha256_le_u32"]


def test_stories15m_golden_logits_fixture(l2b, stories15m):
    """Same check against the committed fixture only (no oracle in the loop)."""
    gold = np.load(os.path.join(GOLDEN, "stories15M_logits.npz"))
    with open(os.path.join(GOLDEN, "stories15M_t0_tokens.json")) as f:
        toks = json.load(f)["tokens"] + [1]
    ck = l2b.read_checkpoint(stories15m, mmap=False)
    with l2b.Transformer(ck) as t:
        token = 1
        for pos in range(int(gold["positions"].max()) + 1):
            lg = t.forward(token, pos)
            if pos in gold["positions"]:
                scale = np.max(np.abs(gold[f"p{pos}_top_val"]))
                assert np.max(np.abs(lg[::16] - gold[f"p{pos}_strided"])) / scale <= REL_TOL
                assert np.max(np.abs(lg[gold[f"p{pos}_top_idx"]] - gold[f"p{pos}_top_val"])) / scale <= REL_TOL
                assert int(np.argmax(lg)) == int(gold[f"p{pos}_top_idx"][0])
            token = toks[pos]


def test_stories15m_argmax_and_generate_paths(l2b, stories15m):
    """l2b_forward_argmax and the on-device loop l2b_generate_argmax reproduce the stream."""
    with open(os.path.join(GOLDEN, "stories15M_t0_tokens.json")) as f:
        gold = json.load(f)
    ck = l2b.read_checkpoint(stories15m, mmap=False)
    with l2b.Transformer(ck) as t:
        token, toks = 1, []
        for pos in range(256):
            nxt = t.forward_argmax(token, pos)
            if nxt == 1:
                break
            toks.append(nxt)
            token = nxt
        assert toks == gold["tokens"]
        t.reset()
        out = t.generate_argmax(1, 0, 256, stop_on_bos=True)
        assert out[-1] == 1 and out[:-1].tolist() == gold["tokens"]
        assert len(out) == gold["forward_calls"]
        # without BOS-stop the loop runs all 256 positions; prefix is unchanged
        t.reset()
        out2 = t.generate_argmax(1, 0, 256, stop_on_bos=False)
        assert len(out2) == 256 and out2[:222].tolist() == out.tolist()
        # prompt forcing (src/main.zig:999-1000): forced tokens are fed, the rest free-runs
        t.reset()
        forced = np.full(256, -1, np.int32)
        forced[:5] = gold["tokens"][:5]
        out3 = t.generate_argmax(1, 0, 40, forced=forced, stop_on_bos=True)
        assert out3.tolist() == gold["tokens"][:40]


def test_rope_table_from_host_is_used(l2b, oracle, stories15m):
    """rope_cos/rope_sin passed through the ABI (host libm) give the same result as the
    library's own table here (same libm); a deliberately wrong table must change the logits."""
    ck = l2b.read_checkpoint(stories15m, mmap=False)
    hs = ck.dim // ck.n_heads
    import ctypes as C
    lib = oracle.load("strict")
    cos = np.zeros((ck.seq_len, hs // 2), np.float32)
    sin = np.zeros_like(cos)
    a, b = C.c_float(), C.c_float()
    for p in range(ck.seq_len):
        for j in range(hs // 2):
            lib.orc_rope_angle(2 * j, hs, p, C.byref(a), C.byref(b))
            cos[p, j], sin[p, j] = a.value, b.value
    with l2b.Transformer(ck) as t0, l2b.Transformer(ck, rope_cos=cos, rope_sin=sin) as t1, \
            l2b.Transformer(ck, rope_cos=np.ones_like(cos), rope_sin=np.zeros_like(sin)) as t2:
        for pos, tok in enumerate([1, 9038, 2501, 263]):
            l0, l1, l2_ = t0.forward(tok, pos), t1.forward(tok, pos), t2.forward(tok, pos)
            assert np.array_equal(l0, l1)
        assert not np.array_equal(l0, l2_)


@pytest.mark.parametrize("shape,seed,steps", [
    ((64, 172, 2, 4, 2, 96, 32), 3, 32),            # GQA branch (:314-320), odd-ish hidden, full context
    ((128, 344, 3, 4, 1, -200, 48), 4, 48),         # MQA + unshared classifier (:112, :942-944)
    ((288, 768, 6, 6, 6, 32000, 256), 15, 40),      # stories15M shape, synthetic weights
    ((768, 2048, 2, 12, 12, 32000, 1024), 110, 12), # stories110M shape, 2 layers
    ((4096, 11008, 1, 32, 32, -32000, 2048), 7, 6), # llama2-7B shape, 1 layer, unshared classifier
])
def test_synthetic_models_match_oracle_teacher_forced(l2b, oracle, shape, seed, steps):
    """Synthetic weights have no argmax margin, so compare logits under teacher forcing
    (SURVEY.md 7 'hard parts') and never free-run."""
    ck, om = make_pair(l2b, oracle, shape, seed)
    toks = teacher_tokens(steps, ck.vocab_size)
    with l2b.Transformer(ck) as t:
        for pos, tok in enumerate(toks):
            got = t.forward(tok, pos)
            want = om.forward(tok, pos)
            assert rel_err(got, want) <= REL_TOL, (pos, rel_err(got, want))
        # RunState parity of the last step (src/main.zig:119-135)
        kc = t.state("key_cache")
        np.testing.assert_allclose(kc, om.state("key_cache"), rtol=0, atol=1e-4 * np.max(np.abs(om.state("key_cache"))))
        vc = t.state("value_cache")
        np.testing.assert_allclose(vc, om.state("value_cache"), rtol=0, atol=1e-4 * np.max(np.abs(om.state("value_cache"))))


def test_device_synthetic_weights_equal_host_mirror(l2b, oracle):
    """l2b_create_synthetic (weights generated in HBM) == l2b_create on the host mirror of the
    same generator: the two contexts must produce bit-identical logits."""
    from llama2_zig_b200.checkpoint import shape_checkpoint
    ck = shape_checkpoint((128, 344, 3, 4, 2, -200, 48))
    ck.data = l2b.synth_checkpoint_host(ck, 21)
    with l2b.Transformer(ck) as a, l2b.Transformer(shape_checkpoint(ck.shape_tuple[:5] + (-200, 48)), synthetic_seed=21) as b:
        for pos, tok in enumerate(teacher_tokens(10, 200)):
            assert np.array_equal(a.forward(tok, pos), b.forward(tok, pos))


def test_long_context_attention_splits(l2b, oracle):
    """Full 1024-position context on a 1-layer 110M-shaped model: exercises the split-KV
    attention path at every length and the last position (maximum size edge case)."""
    ck, om = make_pair(l2b, oracle, (768, 2048, 1, 12, 12, 512, 1024), 5)
    toks = teacher_tokens(1024, 512)
    with l2b.Transformer(ck) as t:
        for pos, tok in enumerate(toks):
            got = t.forward(tok, pos)
            if pos % 97 == 0 or pos >= 1020 or pos in (63, 64, 65, 127, 128, 129):
                want = om.forward(tok, pos)
                assert rel_err(got, want) <= REL_TOL, (pos, rel_err(got, want))
            else:
                om.forward(tok, pos)


def test_call_order_and_argument_errors(l2b, stories15m):
    ck = l2b.read_checkpoint(stories15m, mmap=False)
    with l2b.Transformer(ck) as t:
        with pytest.raises(l2b.L2BError) as e:
            t.forward(1, 5)                      # skips ahead of the KV cache
        assert e.value.status == -7
        with pytest.raises(l2b.L2BError) as e:
            t.forward(32000, 0)                  # token out of range
        assert e.value.status == -1
        with pytest.raises(l2b.L2BError) as e:
            t.forward(1, 256)                    # pos == seq_len
        assert e.value.status == -1
        a = t.forward(1, 0)
        b = t.forward(1, 0)                      # re-running a position is allowed and idempotent
        assert np.array_equal(a, b)


def test_determinism_and_reset(l2b, stories15m):
    ck = l2b.read_checkpoint(stories15m, mmap=False)
    with l2b.Transformer(ck) as t:
        run1 = [t.forward(tok, pos) for pos, tok in enumerate([1, 9038, 2501, 263, 931])]
        t.reset()
        run2 = [t.forward(tok, pos) for pos, tok in enumerate([1, 9038, 2501, 263, 931])]
        for a, b in zip(run1, run2):
            assert np.array_equal(a, b)
        w, kv = t.step_bytes(0)
        assert w == 60_766_848                  # SURVEY.md 8d weight bytes per token
        assert kv == 4 * 6 * 2 * 288


def test_megakernel_path_matches_oracle(stories15m):
    """The opt-in persistent megakernel (L2B_MEGA=1, csrc/l2b_mega.cuh) must produce the same
    stream and logits as the default CUDA-graph + PDL chain.  Runs in a subprocess because the
    switch is read from the environment when a context is created."""
    import subprocess
    import sys
    code = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import llama2_zig_b200 as l2b, oracle_lib as O
from llama2_zig_b200.checkpoint import shape_checkpoint
gold = json.load(open("tests/golden/stories15M_t0_tokens.json"))
ck = l2b.read_checkpoint(sys.argv[1], mmap=False)
with l2b.Transformer(ck) as t:
    out = t.generate_argmax(1, 0, 256, stop_on_bos=True)
    assert out[-1] == 1 and out[:-1].tolist() == gold["tokens"], "stream differs"
    ms, launches = t.last_timing()
    assert launches <= 2 * len(out) + 1, launches          # one kernel per step (+ set_ctl)
    t.reset()
    cfg, shared, data = O.read_checkpoint(sys.argv[1])
    om = O.OracleModel(cfg, data, shared, W=8, kind="strict")
    tok = 1
    for pos in range(24):
        got, want = t.forward(tok, pos), om.forward(tok, pos)
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) <= 1e-4
        tok = int(np.argmax(want))
ck2 = shape_checkpoint((768, 2048, 2, 12, 4, -512, 300)); ck2.data = l2b.synth_checkpoint_host(ck2, 9)
om2 = O.OracleModel(O.make_config(*ck2.shape_tuple), ck2.data, ck2.shared_weights, W=8, kind="strict")
with l2b.Transformer(ck2) as t2:
    for pos in range(300):
        tok = (1 + 7919 * pos) % 512
        got = t2.forward(tok, pos); want = om2.forward(tok, pos)
        if pos % 37 == 0 or pos > 295:
            assert np.max(np.abs(got - want)) / np.max(np.abs(want)) <= 1e-4, pos
print("MEGA_OK")
'''
    env = dict(os.environ, L2B_MEGA="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code, stories15m], capture_output=True, text=True, env=env, cwd=root, timeout=300)
    assert r.returncode == 0 and "MEGA_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
