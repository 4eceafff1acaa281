"""The regression side model (models/music_regression.py) on the CUDA path, through
build_model({"regression": True}) and the C-ABI, against golden vectors recorded from the unmodified reference
(scripts/make_golden_regression.py) and against the CPU oracle at a larger shape."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import rel_err
    from midi_emotion_b200 import MusicRegression, build_model
    from oracle import midi_oracle as O


def _model(g, precision):
    model, _ = build_model(dict(g["cfg"]))
    assert isinstance(model, MusicRegression)
    missing = model.load_state_dict(g["params"])
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.cuda()
    model.precision = precision
    return model


def test_state_dict_keys_match_reference(regression_golden):
    g = regression_golden
    model, _ = build_model(dict(g["cfg"]))
    sd = model.state_dict()
    assert set(sd) == set(g["params"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["params"][k].shape), k


def test_forward_fp32_matches_reference_golden(regression_golden):
    g = regression_golden
    model = _model(g, "fp32").eval()
    with torch.no_grad():
        out = model(g["tokens"].cuda()).cpu()
    assert out.shape == g["out_fp32"].shape and out.dtype == torch.float32
    assert (out - g["out_fp32"]).abs().max().item() < 2e-5


def test_backward_fp32_matches_reference_golden(regression_golden):
    g = regression_golden
    model = _model(g, "fp32").train()
    out = model(g["tokens"].cuda())
    loss = ((out - torch.tensor([[0.8, -0.8]], device="cuda")) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - g["loss_fp32"]) < 2e-5
    floor = 1e-4 * max(v.abs().max().item() for v in g["grads"].values())
    for name, p in model.named_parameters():
        ref = g["grads"][name]
        denom = max(ref.abs().max().item(), floor)
        assert (p.grad.cpu() - ref).abs().max().item() / denom < 1e-3, name


def test_forward_bf16_close_to_reference_autocast(regression_golden):
    g = regression_golden
    model = _model(g, "auto").eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(g["tokens"].cuda())
    assert out.dtype == torch.bfloat16
    out = out.float().cpu()
    ours = (out - g["out_fp32"]).abs().max().item()
    theirs = (g["out_bf16"] - g["out_fp32"]).abs().max().item()
    assert ours <= 2.0 * theirs + 1e-2, (ours, theirs)


@pytest.mark.parametrize("attn", ["simt", "tensor"])
def test_bf16_tensor_core_noncausal_attention_matches_oracle(attn):
    """dh = 64, L = 300 (three query tiles, partial last key tile): the non-causal template of the tcgen05 attention
    kernels, forward and backward, against the fp32 oracle."""
    cfg = dict(vocab_size=300, n_layer=2, n_head=2, d_model=128, d_inner=256, dropout=0.0, d_condition=-1,
               conditioning="none", regression=True)
    shapes = O.regression_param_shapes(cfg)
    g = torch.Generator().manual_seed(5)
    params = {}
    for k, sh in shapes.items():
        if k.endswith("rga.E"):
            params[k] = torch.randn(sh, generator=g) * 0.3
        elif "layernorm" in k:
            params[k] = torch.ones(sh) if k.endswith("weight") else torch.zeros(sh)
        else:
            params[k] = (torch.rand(sh, generator=g) * 2 - 1) * (0.1 if len(sh) == 2 else 0.02)
    tokens = torch.randint(1, 300, (3, 300), generator=g)
    tokens[1, 250:] = 0
    model, _ = build_model(dict(cfg))
    model.load_state_dict(params)
    model = model.cuda().train()
    model.precision, model.attn_impl = "bf16", attn
    out = model(tokens.cuda())
    tgt = torch.tensor([[0.5, -0.5]], device="cuda")
    (out.float() - tgt).abs().mean().backward()            # train.py:284 L1 loss
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    want = O.regression_forward(leaves, cfg, tokens)
    (want - tgt.cpu()).abs().mean().backward()
    assert (out.float().cpu() - want.detach()).abs().max().item() < 3e-2
    for name, p in model.named_parameters():
        ref = leaves[name].grad
        if ref is None or ref.abs().max() < 1e-7:
            continue
        assert rel_err(p.grad.cpu(), ref) < 8e-2, (name, rel_err(p.grad.cpu(), ref))


def test_regression_is_unmasked():
    """Pads are keys like any other token for this model (no_mask=True): changing a pad position's token changes
    the output, and position order matters only through the relative term."""
    cfg = dict(vocab_size=67, n_layer=1, n_head=2, d_model=64, d_inner=128, dropout=0.0, d_condition=-1,
               conditioning="none", regression=True)
    torch.manual_seed(3)
    model, _ = build_model(dict(cfg))
    model = model.cuda().eval()
    tok = torch.randint(1, 67, (2, 40)).cuda()
    with torch.no_grad():
        a = model(tok)
        tok2 = tok.clone()
        tok2[:, 30] = 0
        b = model(tok2)
    assert not torch.equal(a, b)
