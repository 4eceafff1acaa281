"""Kernel-level parity on the B200 (through the C-ABI): tcgen05 GEMM, fp32 GEMM, LayerNorm, input
stage and relative attention, each against a straightforward fp32 evaluation of the same maths."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import dev, gemm_bf16, rel_err, stream
    from midi_emotion_b200 import _lib
    from midi_emotion_b200._lib import ME_BF16, ME_F32, ptr
    from oracle import midi_oracle as O


def _operands(M, N, K, a_mn, b_mn, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    def store(X, mn):
        if not mn:
            return X
        rows = X.shape[0]
        pitch = (rows + 7) // 8 * 8          # [K, rows] storage with a 16-byte aligned row pitch
        buf = torch.zeros(X.shape[1], pitch, device="cuda", dtype=X.dtype)
        buf[:, :rows] = X.t()
        return buf[:, :rows]

    return A, B, store(A, a_mn), store(B, b_mn)


GEMM_SHAPES = [
    # M, N, K, a_mn, b_mn, tile_n
    (128, 256, 64, 0, 0, 256),
    (128, 128, 128, 0, 0, 128),
    (256, 64, 192, 0, 0, 64),
    (384, 32, 64, 0, 0, 32),
    (200, 1007, 96, 0, 0, 0),        # ragged M, N, K (head GEMM, V = 1007, d = 96)
    (520, 2304, 768, 0, 0, 0),       # QKV projection shape
    (300, 3072, 768, 0, 0, 256),
    (256, 768, 3072, 0, 0, 128),
    (256, 256, 128, 0, 1, 256),      # dgrad: B stored [K, N]
    (333, 768, 3072, 0, 1, 128),
    (130, 64, 128, 0, 1, 128),       # N smaller than the tile
    (768, 3072, 512, 1, 1, 256),     # wgrad: both stored [K, rows]
    (1007, 96, 300, 1, 1, 128),      # head wgrad, ragged everything
    (96, 192, 74, 1, 1, 128),
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,tile_n", GEMM_SHAPES)
def test_gemm_tcgen05_matches_fp32(M, N, K, a_mn, b_mn, tile_n):
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn)
    want = A.float() @ B.float().t()
    got = gemm_bf16(As, Bs, M, N, K, a_mn, b_mn, ME_F32, tile_n=tile_n)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert rel_err(got, want) < 1e-5, rel_err(got, want)
    assert (got - want).abs().max() <= 1e-3 * max(1.0, math.sqrt(K))


def test_gemm_tcgen05_against_device_reference_kernel():
    M, N, K = 257, 515, 200
    A, B, As, Bs = _operands(M, N, K, 0, 0, seed=3)
    ref = torch.empty(M, N, device="cuda")
    _lib.call("me_gemm_bf16_reference", ptr(As), ptr(Bs), ptr(ref), M, N, K, K, K, N, 0, 0, stream())
    got = gemm_bf16(As, Bs, M, N, K, 0, 0, ME_F32)
    torch.cuda.synchronize()
    assert rel_err(got, ref) < 2e-6


@pytest.mark.parametrize("flags", ["bias", "bias_relu", "bias_add", "mask"])
@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
def test_gemm_epilogues(flags, out_dtype):
    M, N, K = 300, 520, 256
    A, B, As, Bs = _operands(M, N, K, 0, 0, seed=5)
    bias = torch.randn(N, device="cuda")
    addend = torch.randn(M, N, device="cuda")
    mask = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    od = ME_F32 if out_dtype == "f32" else ME_BF16
    want = A.float() @ B.float().t()
    f = 0
    kw = {}
    if "bias" in flags:
        f |= _lib.EPI_BIAS
        want = want + bias
        kw["bias"] = bias
    if "add" in flags:
        f |= _lib.EPI_ADD_F32
        want = want + addend
        kw["addend"] = addend
    if "relu" in flags:
        f |= _lib.EPI_RELU
        want = want.relu()
    if flags == "mask":
        f |= _lib.EPI_RELU_MASK
        want = torch.where(mask.float() > 0, want, torch.zeros_like(want))
        kw["mask"] = mask
    got = gemm_bf16(As, Bs, M, N, K, 0, 0, od, flags=f, **kw)
    torch.cuda.synchronize()
    if od == ME_BF16:
        assert torch.equal(got, want.to(torch.bfloat16)) or rel_err(got.float(), want) < 3e-3
    else:
        assert rel_err(got, want) < 1e-5


def test_gemm_split_k_matches_single_pass():
    M, N, K = 256, 512, 4096
    A, B, As, Bs = _operands(M, N, K, 1, 1, seed=7)
    one = gemm_bf16(As, Bs, M, N, K, 1, 1, ME_F32, tile_n=256, splits=1)
    four = gemm_bf16(As, Bs, M, N, K, 1, 1, ME_F32, tile_n=256, splits=4)
    torch.cuda.synchronize()
    assert rel_err(four, one) < 1e-5     # fp32 atomics: summation order differs
    assert rel_err(one, A.float() @ B.float().t()) < 1e-5


def test_gemm_rejects_misaligned_operands():
    A = torch.zeros(64, 70, device="cuda", dtype=torch.bfloat16)  # pitch 70 elements: not 16-byte aligned
    B = torch.zeros(64, 70, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        gemm_bf16(A, B, 64, 64, 70, 0, 0, ME_F32)


def test_gemm_output_alignment_is_an_error_not_a_fault():
    """A row pitch that allows 16-byte stores with a base pointer that does not is refused up front; an odd
    pitch (scalar stores) works from any 4-byte aligned base."""
    M, N, K = 128, 64, 64
    A, B, As, Bs = _operands(M, N, K, 0, 0, seed=11)
    want = A.float() @ B.float().t()
    buf = torch.zeros(M * 68 + 8, device="cuda")
    off = buf[1:]                                               # 4-byte offset from a 16-byte aligned allocation
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        _lib.call("me_gemm_bf16_ex", ptr(As), ptr(Bs), ptr(off), M, N, K, K, K, 68, 0, 0, ME_F32, 0, None, None, None, 0,
                  0, 0, stream())
    _lib.call("me_gemm_bf16_ex", ptr(As), ptr(Bs), ptr(off), M, N, K, K, K, 67, 0, 0, ME_F32, 0, None, None, None, 0,
              0, 0, stream())
    torch.cuda.synchronize()
    got = off[:M * 67].view(M, 67)[:, :N]
    assert rel_err(got, want) < 1e-5


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_f32_simt(a_mn, b_mn):
    M, N, K = 150, 203, 77
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    bias = torch.randn(N, device="cuda")
    D = torch.empty(M, N, device="cuda")
    _lib.call("me_gemm_f32", ptr(As), ptr(Bs), ptr(D), M, N, K, As.stride(0), Bs.stride(0), N, a_mn, b_mn,
              _lib.EPI_BIAS | _lib.EPI_RELU, ptr(bias), None, None, 0, stream())
    want = (A.double() @ B.double().t() + bias.double()).relu()
    assert rel_err(D, want) < 1e-6


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("d", [64, 96, 768, 1024])
def test_add_layernorm_forward_backward(dtype, d):
    M = 77
    dt = ME_F32 if dtype == "f32" else ME_BF16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(d)
    x = torch.randn(M, d, device="cuda", generator=g)
    y = torch.randn(M, d, device="cuda", generator=g).to(tdt)
    gamma = torch.randn(d, device="cuda", generator=g)
    beta = torch.randn(d, device="cuda", generator=g)
    out = torch.empty(M, d, device="cuda")
    out_T = out if dtype == "f32" else torch.empty(M, d, device="cuda", dtype=tdt)
    z = torch.empty(M, d, device="cuda")
    mean = torch.empty(M, device="cuda")
    rstd = torch.empty(M, device="cuda")
    _lib.call("me_add_layernorm_forward", ptr(x), ptr(y), dt, ptr(gamma), ptr(beta), 1e-6, M, d, 0.0, 0, ptr(out),
              ptr(out_T), ptr(z), ptr(mean), ptr(rstd), stream())
    xr = x.double().requires_grad_(True)
    yr = y.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    want = torch.nn.functional.layer_norm(xr + yr, (d,), gr, br, 1e-6)
    assert rel_err(out, want.detach()) < 2e-6
    if dtype == "bf16":
        assert torch.equal(out_T, out.to(torch.bfloat16))
    dout = torch.randn(M, d, device="cuda", generator=g)
    want.backward(dout.double())
    dz = torch.empty(M, d, device="cuda")
    dy_T = torch.empty(M, d, device="cuda", dtype=tdt)
    dgam = torch.zeros(d, device="cuda")
    dbet = torch.zeros(d, device="cuda")
    _lib.call("me_add_layernorm_backward", ptr(dout), None, ptr(z), ptr(mean), ptr(rstd), ptr(gamma), M, d, 0.0, 0, dt,
              ptr(dz), ptr(dy_T), ptr(dgam), ptr(dbet), stream())
    assert rel_err(dz, xr.grad) < 1e-5
    assert rel_err(dy_T.float(), yr.grad) < (1e-5 if dtype == "f32" else 5e-3)
    assert rel_err(dgam, gr.grad) < 1e-5
    assert rel_err(dbet, br.grad) < 1e-5


def test_add_layernorm_dropout_is_consistent_between_forward_and_backward():
    M, d, p = 64, 256, 0.25
    x = torch.zeros(M, d, device="cuda")
    y = torch.ones(M, d, device="cuda")
    gamma = torch.ones(d, device="cuda")
    beta = torch.zeros(d, device="cuda")
    out = torch.empty(M, d, device="cuda")
    z = torch.empty(M, d, device="cuda")
    mean = torch.empty(M, device="cuda")
    rstd = torch.empty(M, device="cuda")
    _lib.call("me_add_layernorm_forward", ptr(x), ptr(y), ME_F32, ptr(gamma), ptr(beta), 1e-6, M, d, p, 1234, ptr(out),
              ptr(out), ptr(z), ptr(mean), ptr(rstd), stream())
    keep = z != 0
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02
    assert torch.allclose(z[keep], torch.full_like(z[keep], 1 / (1 - p)))
    # backward: dy = mask/(1-p) * dz, same mask
    dout = torch.randn(M, d, device="cuda")
    dz = torch.empty(M, d, device="cuda")
    dy = torch.empty(M, d, device="cuda")
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    _lib.call("me_add_layernorm_backward", ptr(dout), None, ptr(z), ptr(mean), ptr(rstd), ptr(gamma), M, d, p, 1234,
              ME_F32, ptr(dz), ptr(dy), ptr(dg), ptr(db), stream())
    assert torch.equal(dy != 0, keep & (dz != 0))
    assert torch.allclose(dy[keep], dz[keep] / (1 - p))


def _attn_reference(q, k, v, E, keypad, max_seq):
    """fp64 evaluation of S = (QK^T + Srel)/sqrt(dh), causal + key-pad mask, softmax, PV."""
    B, H, L, dh = q.shape
    q, k, v, E = q.double(), k.double(), v.double(), E.double()
    i = torch.arange(L, device=q.device)[:, None]
    j = torch.arange(L, device=q.device)[None, :]
    idx = (max_seq - 1 - (i - j)).clamp(0, max_seq - 1)
    Eg = E[idx]                                            # [L, L, dh]
    srel = torch.einsum("bhid,ijd->bhij", q, Eg)
    s = (q @ k.transpose(-1, -2) + srel) / math.sqrt(dh)
    mask = (j > i)[None, None] | keypad.bool()[:, None, None, :]
    s = s.masked_fill(mask, float("-inf"))
    return torch.softmax(s, -1) @ v


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("B,H,L,dh", [(2, 2, 37, 48), (1, 3, 130, 64), (2, 1, 64, 32)])
def test_attention_simt_forward_backward(dtype, B, H, L, dh):
    dt = ME_F32 if dtype == "f32" else ME_BF16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    MS = 2048
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = (torch.randn(B, L, 3, H, dh, device="cuda", generator=g) * 0.7).to(tdt)
    E = (torch.randn(MS, dh, device="cuda", generator=g) * 0.3).to(tdt)
    keypad = torch.zeros(B, L, device="cuda", dtype=torch.uint8)
    keypad[0, L - 5:] = 1
    out = torch.empty(B, L, d, device="cuda", dtype=tdt)
    lse = torch.empty(B, H, L, device="cuda")
    a = _lib.AttnArgs()
    a.dtype, a.impl = dt, _lib.ATTN_SIMT
    a.B, a.H, a.Lq, a.Lk, a.dh, a.max_seq, a.q_pos0 = B, H, L, L, dh, MS, 0
    es = qkv.element_size()
    a.q, a.k, a.v, a.E = qkv.data_ptr(), qkv.data_ptr() + d * es, qkv.data_ptr() + 2 * d * es, E.data_ptr()
    for n in "qkv":
        setattr(a, f"{n}_sb", L * 3 * d)
        setattr(a, f"{n}_sh", dh)
    a.q_si = a.k_sj = a.v_sj = 3 * d
    a.keypad, a.keypad_ld = keypad.data_ptr(), L
    a.out, a.o_sb, a.o_si = out.data_ptr(), L * d, d
    a.lse, a.pos_dev, a.stream = lse.data_ptr(), None, stream()
    _lib.call("me_attention_forward", C.byref(a))

    q = qkv[:, :, 0].permute(0, 2, 1, 3).float().requires_grad_(True)
    k = qkv[:, :, 1].permute(0, 2, 1, 3).float().requires_grad_(True)
    v = qkv[:, :, 2].permute(0, 2, 1, 3).float().requires_grad_(True)
    Er = E.float().requires_grad_(True)
    want = _attn_reference(q, k, v, Er, keypad, MS)         # [B,H,L,dh]
    want_m = want.permute(0, 2, 1, 3).reshape(B, L, d)
    tol = 1e-5 if dtype == "f32" else 6e-3
    assert rel_err(out.float(), want_m.detach()) < tol

    dout = (torch.randn(B, L, d, device="cuda", generator=g)).to(tdt)
    want_m.backward(dout.double())
    g_qkv = torch.zeros_like(qkv)
    dE = torch.zeros(MS, dh, device="cuda")
    dsum = torch.empty(B, H, L, device="cuda")
    ba = _lib.AttnBwdArgs()
    ba.f = a
    ba.dout = dout.data_ptr()
    ba.dq, ba.dk, ba.dv = g_qkv.data_ptr(), g_qkv.data_ptr() + d * es, g_qkv.data_ptr() + 2 * d * es
    ba.dE, ba.dsum = dE.data_ptr(), dsum.data_ptr()
    _lib.call("me_attention_backward", C.byref(ba))
    tol = 2e-5 if dtype == "f32" else 1e-2
    assert rel_err(g_qkv[:, :, 0].permute(0, 2, 1, 3).float(), q.grad) < tol
    assert rel_err(g_qkv[:, :, 1].permute(0, 2, 1, 3).float(), k.grad) < tol
    assert rel_err(g_qkv[:, :, 2].permute(0, 2, 1, 3).float(), v.grad) < tol
    assert rel_err(dE, Er.grad) < tol


def test_embed_forward_matches_oracle(golden):
    g = golden
    cfg = g["cfg"]
    from midi_emotion_b200 import build_model
    model, _ = build_model(dict(cfg))
    model.load_state_dict(g["params"])
    model = model.cuda().eval()
    want, mask = O.embed(g["params"], cfg, g["tokens"], g["cond"])
    tokens = g["tokens"].cuda()
    B, L = tokens.shape
    Ls, d = want.shape[1], want.shape[2]
    x = torch.empty(B * Ls, d, device="cuda")
    keypad = torch.empty(B, Ls, device="cuda", dtype=torch.uint8)
    cw0, cb0, cw1, cb1 = model._cond_params()
    cond = g["cond"].cuda() if model.mode != 0 else None
    _lib.call("me_embed_forward", ptr(tokens), ptr(cond), ptr(model.embedding.weight), ptr(cw0), ptr(cb0), ptr(cw1),
              ptr(cb1), ptr(model._pe(tokens.device)), B, L, d, model.d_condition, cfg["vocab_size"], model.mode, 0,
              0.0, 0, ME_F32, ptr(x), ptr(x), ptr(keypad), stream())
    got = x.view(B, Ls, d).cpu()
    assert torch.allclose(got, want, rtol=0, atol=1e-6), (got - want).abs().max()
    # key-pad bitmap == last query row of the reference mask (row Ls-1 has no causal masking)
    assert torch.equal(keypad.cpu().bool(), mask[:, -1, :])


def _run_attention_forward(impl, qkv, E, keypad, B, H, L, dh, save_probs=False):
    MS = E.shape[0]
    d = H * dh
    out = torch.full((B, L, d), float("nan"), device="cuda", dtype=qkv.dtype)
    lse = torch.empty(B, H, L, device="cuda")
    a = _lib.AttnArgs()
    a.dtype, a.impl = (ME_F32 if qkv.dtype == torch.float32 else ME_BF16), impl
    a.B, a.H, a.Lq, a.Lk, a.dh, a.max_seq, a.q_pos0 = B, H, L, L, dh, MS, 0
    es = qkv.element_size()
    a.q, a.k, a.v, a.E = qkv.data_ptr(), qkv.data_ptr() + d * es, qkv.data_ptr() + 2 * d * es, E.data_ptr()
    for n in "qkv":
        setattr(a, f"{n}_sb", L * 3 * d)
        setattr(a, f"{n}_sh", dh)
    a.q_si = a.k_sj = a.v_sj = 3 * d
    a.keypad, a.keypad_ld = (keypad.data_ptr() if keypad is not None else None), L
    a.out, a.o_sb, a.o_si = out.data_ptr(), L * d, d
    a.lse, a.pos_dev, a.stream = lse.data_ptr(), None, stream()
    if save_probs:      # the forward pass leaves its probability tiles for backward
        tiles = B * H * _lib.load().me_attention_saved_tiles(L, 0)
        a._keep = (torch.empty(tiles * 128 * 64, device="cuda", dtype=torch.bfloat16), torch.empty(tiles * 128, device="cuda"))
        a.p_tiles, a.m_tiles = a._keep[0].data_ptr(), a._keep[1].data_ptr()
    _lib.call("me_attention_forward", C.byref(a))
    return out, lse, a


TC_ATTN_SHAPES = [(2, 2, 37, 48), (1, 3, 130, 64), (2, 1, 64, 32), (2, 2, 300, 64), (1, 2, 1024, 64),
                  (1, 1, 2048, 64), (3, 2, 129, 32), (2, 4, 1026, 48)]


@pytest.mark.parametrize("B,H,L,dh", TC_ATTN_SHAPES)
@pytest.mark.parametrize("pad", [False, True])
def test_attention_tensor_core_forward(B, H, L, dh, pad):
    MS = 2048
    g = torch.Generator(device="cuda").manual_seed(L * 7 + dh)
    qkv = (torch.randn(B, L, 3, H, dh, device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    E = (torch.randn(MS, dh, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    keypad = torch.zeros(B, L, device="cuda", dtype=torch.uint8)
    if pad:
        keypad[0, L - min(L // 3, 70):] = 1
        keypad[B - 1, 1::5] = 1
    out, lse, _ = _run_attention_forward(_lib.ATTN_TENSOR, qkv, E, keypad if pad else None, B, H, L, dh)
    q = qkv[:, :, 0].permute(0, 2, 1, 3).float()
    k = qkv[:, :, 1].permute(0, 2, 1, 3).float()
    v = qkv[:, :, 2].permute(0, 2, 1, 3).float()
    want = _attn_reference(q, k, v, E.float(), keypad, MS).permute(0, 2, 1, 3).reshape(B, L, H * dh)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float(), want) < 8e-3, rel_err(out.float(), want)
    assert (out.float() - want).abs().max() < 0.05
    # the two implementations agree on the log-sum-exp they hand to backward
    out2, lse2, _ = _run_attention_forward(_lib.ATTN_SIMT, qkv, E, keypad if pad else None, B, H, L, dh)
    assert torch.allclose(lse, lse2, rtol=1e-3, atol=2e-3), (lse - lse2).abs().max()
    assert rel_err(out.float(), out2.float()) < 8e-3


def _run_attention_backward(impl, a, qkv, out, lse, dout, B, H, L, dh):
    d = H * dh
    es = qkv.element_size()
    g_qkv = torch.full_like(qkv, float("nan"))
    dE = torch.zeros(2048, dh, device="cuda")
    dsum = torch.empty(B, H, L, device="cuda")
    dq_acc = torch.empty(_lib.load().me_attention_backward_workspace_floats(B, H, L, dh, 2048), device="cuda")
    ba = _lib.AttnBwdArgs()
    ba.f = a
    ba.f.impl = impl
    ba.f.out, ba.f.lse = out.data_ptr(), lse.data_ptr()
    ba.dout = dout.data_ptr()
    ba.dq, ba.dk, ba.dv = g_qkv.data_ptr(), g_qkv.data_ptr() + d * es, g_qkv.data_ptr() + 2 * d * es
    ba.dE, ba.dsum, ba.dq_acc = dE.data_ptr(), dsum.data_ptr(), dq_acc.data_ptr()
    _lib.call("me_attention_backward", C.byref(ba))
    return g_qkv, dE


@pytest.mark.parametrize("L,dh", [(700, 64), (1024, 32)])
def test_attention_tensor_core_forward_running_maximum_moves(L, dh):
    """Keys whose logits grow along the sequence: the row maximum rises by far more than the 2^8 the forward
    kernel's lazy running maximum tolerates, so its accumulator-rescale path (in TMEM) runs on most tiles."""
    B, H, MS = 2, 2, 2048
    g = torch.Generator(device="cuda").manual_seed(L + dh)
    qkv = torch.randn(B, L, 3, H, dh, device="cuda", generator=g)
    ramp = (1.0 + torch.arange(L, device="cuda").float() / 48.0).view(1, L, 1, 1)
    qkv[:, :, 0] = qkv[:, :, 0].abs() * 0.9                   # q >= 0, k >= 0 and growing: logit ~ position
    qkv[:, :, 1] = qkv[:, :, 1].abs() * ramp * 0.9
    qkv = qkv.to(torch.bfloat16)
    E = (torch.randn(MS, dh, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    out, lse, a = _run_attention_forward(_lib.ATTN_TENSOR, qkv, E, None, B, H, L, dh, save_probs=True)
    q = qkv[:, :, 0].permute(0, 2, 1, 3).float()
    k = qkv[:, :, 1].permute(0, 2, 1, 3).float()
    v = qkv[:, :, 2].permute(0, 2, 1, 3).float()
    logits = (q @ k.transpose(-1, -2)) / dh ** 0.5
    spread = (logits[..., -1, :].max(-1).values - logits[..., -1, :64].max(-1).values).min().item()
    assert spread * 1.4427 > 3 * 8, spread                     # the case is what it claims to be
    keypad = torch.zeros(B, L, device="cuda", dtype=torch.uint8)
    want = _attn_reference(q, k, v, E.float(), keypad, MS).permute(0, 2, 1, 3).reshape(B, L, H * dh)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert rel_err(out.float(), want) < 8e-3, rel_err(out.float(), want)
    out2, lse2, _ = _run_attention_forward(_lib.ATTN_SIMT, qkv, E, None, B, H, L, dh)
    assert torch.allclose(lse, lse2, rtol=1e-3, atol=5e-3), (lse - lse2).abs().max()
    # ... and the saved probability tiles (exponent offsets per tile) still drive a correct backward
    dout = (torch.randn(B, L, H * dh, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    g_tc, dE_tc = _run_attention_backward(_lib.ATTN_TENSOR, a, qkv, out, lse, dout, B, H, L, dh)
    g_si, dE_si = _run_attention_backward(_lib.ATTN_SIMT, a, qkv, out, lse, dout, B, H, L, dh)
    assert rel_err(g_tc.float(), g_si.float()) < 2e-2, rel_err(g_tc.float(), g_si.float())
    assert rel_err(dE_tc, dE_si) < 2e-2, rel_err(dE_tc, dE_si)


@pytest.mark.parametrize("B,H,L,dh", [(2, 2, 37, 48), (1, 3, 130, 64), (2, 1, 64, 32), (2, 2, 300, 64),
                                      (1, 2, 1024, 64), (1, 1, 2048, 64), (2, 4, 1026, 48)])
@pytest.mark.parametrize("pad", [False, True])
@pytest.mark.parametrize("saved", [False, True], ids=["recompute", "saved_probs"])
def test_attention_tensor_core_backward(B, H, L, dh, pad, saved):
    MS = 2048
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(L * 3 + dh)
    qkv = (torch.randn(B, L, 3, H, dh, device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    E = (torch.randn(MS, dh, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    keypad = torch.zeros(B, L, device="cuda", dtype=torch.uint8)
    if pad:
        keypad[0, L - min(L // 3, 70):] = 1
        keypad[B - 1, 1::5] = 1
    kp = keypad if pad else None
    out, lse, a = _run_attention_forward(_lib.ATTN_TENSOR, qkv, E, kp, B, H, L, dh, save_probs=saved)
    dout = torch.randn(B, L, d, device="cuda", generator=g).to(torch.bfloat16)
    g_tc, dE_tc = _run_attention_backward(_lib.ATTN_TENSOR, a, qkv, out, lse, dout, B, H, L, dh)
    torch.cuda.synchronize()
    assert torch.isfinite(g_tc.float()).all() and torch.isfinite(dE_tc).all()

    q = qkv[:, :, 0].permute(0, 2, 1, 3).float().requires_grad_(True)
    k = qkv[:, :, 1].permute(0, 2, 1, 3).float().requires_grad_(True)
    v = qkv[:, :, 2].permute(0, 2, 1, 3).float().requires_grad_(True)
    Er = E.float().requires_grad_(True)
    want = _attn_reference(q, k, v, Er, keypad, MS).permute(0, 2, 1, 3).reshape(B, L, d)
    want.backward(dout.double())
    tol = 1.5e-2
    assert rel_err(g_tc[:, :, 0].permute(0, 2, 1, 3).float(), q.grad) < tol
    assert rel_err(g_tc[:, :, 1].permute(0, 2, 1, 3).float(), k.grad) < tol
    assert rel_err(g_tc[:, :, 2].permute(0, 2, 1, 3).float(), v.grad) < tol
    assert rel_err(dE_tc, Er.grad) < tol
    # and against the SIMT backward on the same saved forward
    g_si, dE_si = _run_attention_backward(_lib.ATTN_SIMT, a, qkv, out, lse, dout, B, H, L, dh)
    assert rel_err(g_tc.float(), g_si.float()) < tol
    assert rel_err(dE_tc, dE_si) < tol


PAIR_SHAPES = [
    # M, N, K, a_mn, b_mn, splits
    (512, 512, 128, 0, 0, 0),
    (256, 256, 64, 0, 0, 0),
    (1000, 1007, 96, 0, 0, 0),       # ragged M, N, K
    (2048, 2304, 768, 0, 0, 0),
    (777, 768, 3072, 0, 1, 0),       # dgrad
    (768, 3072, 1024, 1, 1, 2),      # wgrad, split-K
    (1007, 96, 300, 1, 1, 1),        # narrow N forced through the pair kernel
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,splits", PAIR_SHAPES)
def test_gemm_tcgen05_cta_pair_matches_fp32(M, N, K, a_mn, b_mn, splits):
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn, seed=11)
    want = A.float() @ B.float().t()
    got = gemm_bf16(As, Bs, M, N, K, a_mn, b_mn, ME_F32, tile_n=512, splits=splits)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert rel_err(got, want) < 1e-5, rel_err(got, want)


@pytest.mark.parametrize("flags", ["bias_relu", "bias_add", "mask"])
def test_gemm_tcgen05_cta_pair_epilogues(flags):
    M, N, K = 520, 768, 256
    A, B, As, Bs = _operands(M, N, K, 0, 0, seed=13)
    bias = torch.randn(N, device="cuda")
    addend = torch.randn(M, N, device="cuda")
    mask = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    want = A.float() @ B.float().t()
    f, kw = 0, {}
    if "bias" in flags:
        f |= _lib.EPI_BIAS
        want = want + bias
        kw["bias"] = bias
    if "add" in flags:
        f |= _lib.EPI_ADD_F32
        want = want + addend
        kw["addend"] = addend
    if "relu" in flags:
        f |= _lib.EPI_RELU
        want = want.relu()
    if flags == "mask":
        f |= _lib.EPI_RELU_MASK
        want = torch.where(mask.float() > 0, want, torch.zeros_like(want))
        kw["mask"] = mask
    got = gemm_bf16(As, Bs, M, N, K, 0, 0, ME_BF16, flags=f, tile_n=512, **kw)
    torch.cuda.synchronize()
    assert rel_err(got.float(), want) < 3e-3


@pytest.mark.parametrize("M,N,K,b_mn", [(300, 1007, 128, 0), (1000, 520, 192, 1), (256, 256, 64, 0)])
def test_gemm_tcgen05_cta_pair_staged_store_ragged(M, N, K, b_mn):
    """bf16 output through the shared-memory staging tile + TMA store, ragged M / N, padded row pitch."""
    A, B, As, Bs = _operands(M, N, K, 0, b_mn, seed=17)
    bias = torch.randn(N, device="cuda")
    ldd = (N + 7) // 8 * 8
    want = (A.float() @ B.float().t() + bias)
    got = gemm_bf16(As, Bs, M, N, K, 0, b_mn, ME_BF16, flags=_lib.EPI_BIAS, bias=bias, tile_n=512, ldd=ldd)
    torch.cuda.synchronize()
    assert rel_err(got[:, :N].float(), want) < 3e-3
    if ldd != N:
        pad = got[:, N:].float()                          # pad columns: untouched (NaN fill) or written as zero
        assert (torch.isnan(pad) | (pad == 0)).all()


# ---------------------------------------------------------------------------------------------
# column sums with scratch (no same-address atomics) and the batched weight-copy refresh
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,ld", [(4096, 256, 256), (2048, 1007, 1008), (5000, 2304, 2304), (3000, 3072, 3080),
                                    (100, 512, 512)])
def test_colsum_with_scratch_matches_torch(M, N, ld):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    X = torch.randn(M, ld, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.full((N,), 0.5, device="cuda")
    ws = torch.empty(160 * N, device="cuda")
    _lib.call("me_colsum_ws", ptr(X), ME_BF16, M, N, ld, ptr(out), ptr(ws), ws.numel(), stream())
    want = X[:, :N].float().sum(0) + 0.5          # accumulates into `out`
    assert torch.allclose(out, want, rtol=1e-4, atol=1e-2)


def test_convert_batched_matches_convert_2d():
    import ctypes as C
    srcs = [torch.randn(33, 40, device="cuda"), torch.randn(7, 64, device="cuda"), torch.randn(1, 100, device="cuda"),
            torch.randn(16, 24, device="cuda").to(torch.bfloat16)]
    dsts = [torch.empty(33, 40, device="cuda", dtype=torch.bfloat16), torch.empty(7, 72, device="cuda", dtype=torch.bfloat16),
            torch.empty(1, 100, device="cuda"), torch.empty(16, 24, device="cuda")]
    code = {torch.float32: ME_F32, torch.bfloat16: ME_BF16}
    table = (_lib.ConvertDesc * len(srcs))(*[
        _lib.ConvertDesc(s.data_ptr(), d.data_ptr(), s.shape[0], s.shape[1], s.stride(0), d.stride(0), code[s.dtype], code[d.dtype])
        for s, d in zip(srcs, dsts)])
    dev_table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).cuda()
    _lib.call("me_convert_batched", ptr(dev_table), len(srcs), stream())
    for s, d in zip(srcs, dsts):
        cols = s.shape[1]
        assert torch.equal(d[:, :cols], s.to(d.dtype))
        assert (d[:, cols:] == 0).all()               # pad columns are zeroed
