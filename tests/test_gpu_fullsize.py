"""Size-independent properties at BASELINE.json's full shapes and batch sizes.

The direct comparison with the CPU oracle at these shapes lives in tests/test_gpu_parity_oracle.py (the oracle runs a
12-layer / 768-wide / 1024-token sequence in about two seconds); the checks here are the properties that hold at
any size and batch, on the full batch the bench uses:
  * causality        -- logits at positions < t do not change (bit-exact) when token t changes;
  * batch independence -- permuting the sequences of a batch permutes the logits (bit-exact);
  * key-pad mask     -- tail pads leave the prefix logits unchanged;
  * two implementations -- the tcgen05 bf16 path and the exact-order fp32 SIMT path agree to the bf16
                        tolerance the golden tests establish, gradients included;
  * decode           -- the KV-cache step equals the last position of the full forward pass;
  * conditioning sweep (configs[4]) and the 24L/1024d/16h, L = 2048 = max_seq shape (configs[2], fewer layers).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import rel_err
    from midi_emotion_b200 import KVCacheDecoder, build_model

CFG2 = dict(vocab_size=1007, n_layer=12, n_head=12, d_model=768, d_inner=3072, dropout=0.0, d_condition=192,
            conditioning="continuous_concat")


def _build(cfg, seed=1234, e_scale=0.2):
    torch.manual_seed(seed)
    model, _ = build_model(dict(cfg))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("rga.E"):
                p.mul_(e_scale)   # N(0,1) E saturates the logits of a random-init deep model
    return model.cuda().eval()


def _batch(cfg, B, L, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(1, cfg["vocab_size"] if cfg["conditioning"] != "discrete_token" else 1007, (B, L), generator=g)
    tok[:, 0] = 1
    if cfg["conditioning"] == "discrete_token":
        tok[:, 0] = torch.randint(1007, 1017, (B,), generator=g)   # <V*>/<A*> emotion tokens lead the sequence
    cond = torch.rand(B, 2, generator=g) * 2 - 1
    return tok.cuda(), cond.cuda()


def _fwd(model, tok, cond, bf16=True):
    with torch.no_grad():
        if bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return model(tok, cond).float()
        model.precision = "fp32"
        try:
            return model(tok, cond).float()
        finally:
            model.precision = "auto"


@pytest.mark.parametrize("bf16", [True, False])
def test_cfg2_causality_and_batch_independence(bf16):
    L = 1024 if bf16 else 384       # the exact fp32 path runs on CUDA cores: keep it short
    model = _build(CFG2)
    tok, cond = _batch(CFG2, 3, L, 5)
    base = _fwd(model, tok, cond, bf16)
    assert torch.isfinite(base).all()
    t = L // 2 + 7
    tok2 = tok.clone()
    tok2[:, t] = (tok[:, t] % 1000) + 3
    pert = _fwd(model, tok2, cond, bf16)
    assert torch.equal(base[:, :t], pert[:, :t]), "a later token changed earlier logits"
    assert not torch.equal(base[:, t:], pert[:, t:])
    perm = torch.tensor([2, 0, 1], device="cuda")
    assert torch.equal(_fwd(model, tok[perm], cond[perm], bf16), base[perm]), "batch rows are not independent"


def test_cfg2_tail_padding_leaves_prefix_unchanged():
    model = _build(CFG2)
    tok, cond = _batch(CFG2, 2, 1024, 6)
    base = _fwd(model, tok, cond)
    padded = tok.clone()
    padded[0, 900:] = 0
    padded[1, 517:] = 0
    out = _fwd(model, padded, cond)
    assert torch.equal(out[0, :900], base[0, :900]) and torch.equal(out[1, :517], base[1, :517])
    assert torch.isfinite(out).all()


def test_cfg2_tensor_core_path_agrees_with_exact_fp32_path_forward_and_backward():
    cfg = dict(CFG2, n_layer=4)
    model = _build(cfg).train()
    tok, cond = _batch(cfg, 2, 512, 7)
    tgt = torch.roll(tok, -1, 1)
    grads = {}
    for prec in ("fp32", "auto"):
        model.zero_grad()
        model.precision = prec
        if prec == "auto":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = model(tok, cond)
        else:
            logits = model(tok, cond)
        loss = torch.nn.functional.cross_entropy(logits.float().reshape(-1, logits.size(-1)), tgt.reshape(-1),
                                                 ignore_index=0)
        loss.backward()
        grads[prec] = (logits.detach().float(), loss.item(),
                       {n: p.grad.detach().clone() for n, p in model.named_parameters()})
    model.precision = "auto"
    l32, loss32, g32 = grads["fp32"]
    l16, loss16, g16 = grads["auto"]
    assert rel_err(l16, l32) < 2e-2                      # bf16 tolerance of the golden tests (BF16_VS_REF + margin)
    assert abs(loss16 - loss32) < 2e-2 * max(1.0, abs(loss32))
    top = max(v.abs().max().item() for v in g32.values())
    for n in g32:
        if g32[n].abs().max().item() < 1e-3 * top:
            continue                                     # (e.g. Wk.bias: identically zero up to rounding)
        assert rel_err(g16[n], g32[n]) < 6e-2, n


def test_cfg2_decode_step_equals_full_forward_last_position():
    model = _build(CFG2)
    B, L = 4, 600
    tok, cond = _batch(CFG2, B, L, 8)
    full = _fwd(model, tok, cond)                     # [B, L, V]
    dec = KVCacheDecoder(model, B, max_len=1024, precision="bf16")
    last = dec.prefill(tok[:, :L - 3], cond)
    outs = [last.clone()]
    for t in range(L - 3, L):
        outs.append(dec.step(tok[:, t]).clone())   # step() returns a view of a static buffer
    for k, o in enumerate(outs):
        ref = full[:, L - 4 + k]
        assert rel_err(o.float(), ref) < 2e-2, k


@pytest.mark.parametrize("mode", ["discrete_token", "continuous_token", "continuous_concat", "none"])
def test_conditioning_sweep_at_cfg2_width(mode):
    cfg = dict(CFG2, n_layer=3, conditioning=mode)
    if mode == "discrete_token":
        cfg.update(vocab_size=1017, d_condition=-1)
    elif mode in ("continuous_token", "none"):
        cfg.update(d_condition=-1)
    L = 1022 if mode == "continuous_token" else 1024   # ctoken prepends two positions: Ls = 1024
    model = _build(cfg)
    tok, cond = _batch(cfg, 2, L, 9)
    if mode in ("none", "discrete_token"):
        cond = torch.full_like(cond, float("nan"))      # the reference passes NaN conditions there
    out = _fwd(model, tok, cond)
    Ls = L + 2 if mode == "continuous_token" else L
    assert out.shape == (2, Ls, cfg["vocab_size"]) and torch.isfinite(out).all()
    tok2 = tok.clone()
    tok2[:, 700] = (tok[:, 700] % 1000) + 3
    cut = 700 + (2 if mode == "continuous_token" else 0)
    assert torch.equal(out[:, :cut], _fwd(model, tok2, cond)[:, :cut])
    if mode.startswith("continuous"):
        other = _fwd(model, tok, -cond)
        assert not torch.equal(out, other), "the (valence, arousal) pair must reach the logits"


def test_cfg3_width_at_max_sequence_length_trains():
    """24L/1024d/16h of configs[2] at L = 2048 = max_seq (the longest sequence the relative table covers),
    two layers: forward is causal, backward produces finite gradients for every parameter."""
    cfg = dict(vocab_size=1007, n_layer=2, n_head=16, d_model=1024, d_inner=4096, dropout=0.0, d_condition=192,
               conditioning="continuous_concat")
    model = _build(cfg).train()
    tok, cond = _batch(cfg, 2, 2048, 10)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(tok, cond)
    loss = torch.nn.functional.cross_entropy(logits.float().reshape(-1, 1007), torch.roll(tok, -1, 1).reshape(-1),
                                             ignore_index=0)
    loss.backward()
    assert torch.isfinite(loss)
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # the relative table is touched from its first row (distance 2047) to its last (distance 0)
    gE = model.enc_layers[0].rga.E.grad
    assert gE[0].abs().sum() > 0 and gE[-1].abs().sum() > 0
    model.eval()
    base = _fwd(model, tok, cond)
    tok2 = tok.clone()
    tok2[:, 2047] = (tok[:, 2047] % 1000) + 3
    assert torch.equal(base[:, :2047], _fwd(model, tok2, cond)[:, :2047])
