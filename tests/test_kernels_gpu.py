"""Per-kernel parity tests on the B200, through the C ABI (crossscore_b200._lib).

Each CUDA entry point is compared with a plain PyTorch fp32/fp64 restatement of the same operator on the
same seeded inputs (for the fp32 parity mode: tight tolerances; for the bf16 tensor-core path: bf16
round-off of operands/outputs).  Ragged sizes (M, L not multiples of the tile) are the default here
because the real shapes are ragged (T = 1370, P = 1369, M = 6845).
"""
import math

import pytest
import torch

from crossscore_b200 import _lib
from crossscore_b200._lib import ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, DT_BF16, DT_F16, DT_F32, DT_TF32, call

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def st():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def act_ref(x, act):
    if act == ACT_GELU:
        return 0.5 * x * (1 + torch.erf(x / math.sqrt(2)))
    if act == ACT_RELU:
        return torch.relu(x)
    if act == ACT_LEAKY:
        return torch.where(x >= 0, x, 0.01 * x)
    return x


def test_library_and_device():
    lib = _lib.load()
    assert lib.xs_version() == 1
    call("xs_device_check")


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,act", [(300, 384, 384, ACT_NONE), (1000, 1152, 384, ACT_GELU), (77, 196, 588, ACT_RELU),
                                       (513, 384, 1536, ACT_LEAKY)])
def test_gemm_f32(M, N, K, act):
    A, W, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    out = torch.empty(M, N, device=DEV)
    call("xs_gemm_bias_act", P(A), K, P(W), K, P(b), P(out), N, M, N, K, act, DT_F32, DT_F32, st())
    ref = act_ref(A.double() @ W.double().T + b.double(), act)
    assert (out.double() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("M,N,K,act", [
    (128, 192, 64, ACT_NONE),       # one tile, one k-block
    (128, 384, 384, ACT_NONE),      # BN=192 x2
    (300, 384, 384, ACT_NONE),      # M tail
    (1000, 1152, 384, ACT_GELU),    # qkv-like / gelu
    (1370 * 3, 1536, 384, ACT_GELU),
    (1369, 384, 592, ACT_NONE),     # patch-embed K (zero-filled K tail block)
    (777, 384, 1536, ACT_NONE),     # fc2 K
    (1369, 2048, 384, ACT_NONE),    # BN=256
    (1369, 512, 384, ACT_RELU),
    (37800, 384, 384, ACT_LEAKY),   # A-stationary CTA pairs (148 row blocks = two full waves), BN=192, ragged M
    (18900, 1152, 384, ACT_NONE),   # A-stationary CTA pairs, 6 n-tiles per row block, ragged M (peer CTA rows out of range)
    (37801, 1536, 384, ACT_GELU),   # A-stationary CTA pairs, BN=256
    (40000, 384, 384, ACT_LEAKY),   # 157 row blocks = 2.1 waves: streaming CTA pairs at K = 384 (many tiles per CTA, phase wrap)
    (19999, 1152, 384, ACT_NONE),   # streaming CTA pairs, short K, ragged M
    (20000, 1536, 384, ACT_GELU),   # streaming CTA pairs, short K, BN=256
    (19999, 384, 1536, ACT_NONE),   # streaming CTA pairs (cta_group::2), long K
    (18945, 512, 1024, ACT_RELU),   # streaming CTA pairs, BN=256
])
def test_gemm_bf16_tc(M, N, K, act):
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    W = rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    b = rnd(N, seed=3)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    call("xs_gemm_bias_act", P(A), K, P(W), K, P(b), P(out), N, M, N, K, act, DT_BF16, DT_BF16, st())
    torch.cuda.synchronize()
    ref = act_ref(A.float() @ W.float().T + b, act)
    err = (out.float() - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    assert torch.isfinite(out.float()).all()
    assert (err <= tol).all(), f"max err {err.max().item()} at {err.argmax().item()}"
    assert err.mean() < 5e-3


@pytest.mark.parametrize("M,N,K,act,out_bf16", [
    (128, 192, 32, ACT_NONE, False),     # one tile, one k-block
    (300, 384, 384, ACT_NONE, False),    # decoder out-proj / linear2
    (1369, 384, 384, ACT_RELU, False),   # linear1
    (1369, 384, 384, ACT_LEAKY, False),  # head.0
    (2738, 384, 588, ACT_NONE, False),   # patch embedding: K tail block zero-filled by TMA
    (1369, 1536, 384, ACT_NONE, True),   # decoder self-attention in-proj -> bf16 q|k|v
    (1369, 512, 384, ACT_NONE, True),    # cross-attention q-proj
    (30000, 384, 384, ACT_NONE, False),  # persistent loop; CTA pairs (cta_group::2, kind::tf32)
    (43808, 384, 588, ACT_RELU, False),  # CTA pairs, patch-embedding K (zero-filled tail block), ragged M
    (19001, 1152, 384, ACT_LEAKY, False),  # CTA pairs, six n-tiles, peer CTA rows out of range
])
def test_gemm_tf32_tc(M, N, K, act, out_bf16):
    A, W, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    odt = torch.bfloat16 if out_bf16 else torch.float32
    out = torch.full((M, N), float("nan"), device=DEV, dtype=odt)
    call("xs_gemm_bias_act", P(A), K, P(W), K, P(b), P(out), N, M, N, K, act, DT_TF32,
         DT_BF16 if out_bf16 else DT_F32, st())
    torch.cuda.synchronize()
    ref = act_ref(A.double() @ W.double().T + b.double(), act)
    err = (out.double() - ref).abs()
    assert torch.isfinite(out.float()).all()
    tol = (0.02 + 0.01 * ref.abs()) if out_bf16 else (8e-3 + 2e-3 * ref.abs())  # TF32 truncates both operands
    assert (err <= tol).all(), f"max err {err.max().item()}"
    assert err.mean() < (5e-3 if out_bf16 else 1.5e-3)


def test_gemm_bf16_in_f32_out():
    M, N, K = 1000, 384, 1536
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    W = rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    b = rnd(N, seed=3)
    out = torch.full((M, N), float("nan"), device=DEV)
    call("xs_gemm_bias_act", P(A), K, P(W), K, P(b), P(out), N, M, N, K, ACT_NONE, DT_BF16, DT_F32, st())
    ref = A.double() @ W.double().T + b.double()
    assert (out.double() - ref).abs().max() < 2e-3


@pytest.mark.parametrize("M,N,K", [
    (128, 384, 384),      # single CTA per tile
    (1000, 384, 1536),    # single CTA, long K, ragged M
    (37801, 384, 384),    # A-stationary CTA pairs (proj), two full waves, ragged M: the store clips the tail rows
    (40001, 384, 384),    # 2.1 waves of row blocks: streaming CTA pairs at K = 384
    (39999, 384, 1536),   # streaming CTA pairs (fc2)
])
def test_gemm_bias_residual(M, N, K):
    """h += A W^T + b in place (TMA reduce-store epilogue): every element gets exactly one fp32 add."""
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    W = rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    b = rnd(N, seed=3)
    h0 = rnd(M + 3, N, seed=4, scale=3.0)  # 3 guard rows behind the tail
    h = h0.clone()
    call("xs_gemm_bias_residual", P(A), K, P(W), K, P(b), P(h), N, M, N, K, DT_BF16, st())
    torch.cuda.synchronize()
    ref = h0[:M].double() + A.double() @ W.double().T + b.double()
    assert (h[:M].double() - ref).abs().max() < 2e-3
    assert torch.equal(h[M:], h0[M:])
    # twice more: accumulates, deterministic
    h2 = h0.clone()
    call("xs_gemm_bias_residual", P(A), K, P(W), K, P(b), P(h2), N, M, N, K, DT_BF16, st())
    torch.cuda.synchronize()
    assert torch.equal(h2, h)
    with pytest.raises(_lib.XsError):
        call("xs_gemm_bias_residual", P(A), K, P(W), K, P(b), P(h), N, M, N, K, DT_F32, st())


@pytest.mark.parametrize("M,K", [(20000, 384), (19001, 384), (20000, 1536), (37931, 1536), (1000, 384), (9000, 1536)])
def test_gemm_bias_residual_ln(M, K):
    """h += A W^T + b and y = LayerNorm(h) in one kernel (fused epilogue for M large enough to fill the GPU: CTA pairs,
    A-stationary for K = 384 and streaming for K = 1536; ragged row tails; the small shapes take the two-launch path).
    Rows carry a large common offset and a few outlier channels: the statistics are merged from two half rows."""
    N = 384
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    W = rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    b = rnd(N, seed=3)
    h0 = rnd(M, N, seed=4, scale=2.0) + 7.0
    h0[:, [5, 133, 301]] *= 40.0
    g, be = rnd(N, seed=5) * 0.2 + 1, rnd(N, seed=6, scale=0.1)
    h = h0.clone()
    y = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    call("xs_gemm_bias_residual_ln", P(A), K, P(W), K, P(b), P(h), N, P(g), P(be), 1e-6, P(y), N, M, N, K, DT_BF16, st())
    torch.cuda.synchronize()
    href = h0.double() + A.double() @ W.double().T + b.double()
    assert ((h.double() - href).abs() <= 2e-3 + 1e-5 * href.abs()).all()
    yref = torch.nn.functional.layer_norm(href, (N,), g.double(), be.double(), 1e-6)
    err = (y.double() - yref).abs()
    assert torch.isfinite(y).all()
    assert (err <= 0.02 + 0.01 * yref.abs()).all(), err.max().item()


@pytest.mark.parametrize("M,K", [(20000, 384), (19001, 384), (20000, 1536), (37931, 1536), (1000, 384), (9000, 1536)])
def test_gemm_bias_residual_stats(M, K):
    """Producer half of the folded LayerNorm: h += A W^T + b, hb = bf16(h), stats = (mean, rstd) per row -- fused epilogue
    for large M (both CTA-pair modes, ragged row tails), two launches otherwise.  Rows carry a large common offset and
    outlier channels."""
    N = 384
    A = rnd(M, K, seed=1, dtype=torch.bfloat16)
    W = rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    b = rnd(N, seed=3)
    h0 = rnd(M, N, seed=4, scale=2.0) + 7.0
    h0[:, [5, 133, 301]] *= 40.0
    h = h0.clone()
    hb = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.full((M, 2), float("nan"), device=DEV)
    call("xs_gemm_bias_residual_stats", P(A), K, P(W), K, P(b), P(h), N, P(hb), N, P(stats), 1e-6, M, N, K, DT_BF16, st())
    torch.cuda.synchronize()
    href = h0.double() + A.double() @ W.double().T + b.double()
    assert ((h.double() - href).abs() <= 2e-3 + 1e-5 * href.abs()).all()
    assert torch.equal(hb, h.bfloat16())                       # the copy is the rounding of what was stored
    mean = h.double().mean(-1)
    rstd = torch.rsqrt(h.double().var(-1, unbiased=False) + 1e-6)
    assert torch.isfinite(stats).all()
    assert ((stats[:, 0].double() - mean).abs() <= 1e-4 * (1 + mean.abs())).all()
    assert ((stats[:, 1].double() - rstd).abs() <= 1e-4 * rstd).all()


@pytest.mark.parametrize("M,N,act", [(20000, 1152, ACT_NONE), (19001, 1536, ACT_GELU), (37931, 1152, ACT_NONE),
                                     (1000, 1536, ACT_GELU), (777, 1152, ACT_NONE)])
def test_gemm_ln_folded(M, N, act):
    """Consumer half: act(LayerNorm(h) W^T + b) from hb = bf16(h), Wf = bf16(gamma * W), c1 = row sums of Wf,
    c0 = W beta + b and the per-row (mean, rstd).  Checked tightly against the folded formula in fp64 on the same bf16
    operands, and at bf16 level against the LayerNorm-then-Linear it replaces."""
    K = 384
    h = rnd(M, K, seed=4, scale=2.0) + 3.0
    h[:, [5, 133, 301]] *= 20.0
    g, be = rnd(K, seed=5) * 0.2 + 1, rnd(K, seed=6, scale=0.1)
    W = rnd(N, K, seed=2, scale=0.05)
    b = rnd(N, seed=3)
    Wf = (W * g[None, :]).bfloat16().contiguous()
    c1 = Wf.float().sum(-1).contiguous()
    c0 = (W @ be + b).contiguous()
    hb = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    stats = torch.empty(M, 2, device=DEV)
    call("xs_row_stats", P(h), P(hb), P(stats), 1e-6, M, st())
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    call("xs_gemm_ln_folded", P(hb), K, P(Wf), K, P(c0), P(c1), P(stats), P(out), N, M, N, K, act, DT_BF16, st())
    torch.cuda.synchronize()
    assert torch.equal(hb, h.bfloat16())
    f = (lambda x: torch.nn.functional.gelu(x)) if act == ACT_GELU else (lambda x: x)
    mean, rstd = stats[:, :1].double(), stats[:, 1:].double()
    folded = f(rstd * (hb.double() @ Wf.double().T - mean * c1.double()) + c0.double())
    err = (out.double() - folded).abs()
    assert torch.isfinite(out).all()
    assert (err <= 2e-3 + 6e-3 * folded.abs()).all(), err.max().item()   # bf16 rounding of the output
    ref = f(torch.nn.functional.layer_norm(h.double(), (K,), g.double(), be.double(), 1e-6) @ W.double().T + b.double())
    err = (out.double() - ref).abs()
    assert err.mean() < 6e-3 and (err <= 0.06 + 0.02 * ref.abs()).all(), (err.mean().item(), err.max().item())
    with pytest.raises(_lib.XsError):
        call("xs_gemm_ln_folded", P(hb), K, P(Wf), K, P(c0), P(c1), P(stats), P(out), N, M, N, K, ACT_RELU, DT_BF16, st())


def test_gemm_bf16_rejects_bad_shapes():
    A = rnd(128, 384, dtype=torch.bfloat16)
    W = rnd(200, 384, dtype=torch.bfloat16)
    b = rnd(200)
    out = torch.empty(128, 200, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.XsError):
        call("xs_gemm_bias_act", P(A), 384, P(W), 384, P(b), P(out), 200, 128, 200, 384, ACT_NONE, DT_BF16, DT_BF16, st())


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
@pytest.fixture(params=[0, 1], ids=["layout64", "layout_pair"])
def layout(request):
    """Both bf16 flash-attention layouts: 64-key blocks with two CTAs per SM (xs_attn_tc.cu) and two query tiles per
    CTA sharing 128-key blocks (xs_attn_tc2.cu)."""
    _lib.load().xs_attn_set_layout(request.param)
    yield request.param
    _lib.load().xs_attn_set_layout(1)  # the library default


def attn_ref(q, k, v, scale):
    # q (B,H,Lq,d), k/v (B,H,Lk,d) double
    s = (q @ k.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, -1)
    return torch.softmax(s, -1) @ v, lse


def run_attn(dtype_flag, B, H, Lq, Lk, d, slot, nsplit=1, kv_shared=False, seed=0, qscale=1.0, force_f32_out=False,
             scale=None):
    adt = {DT_BF16: torch.bfloat16, DT_F16: torch.float16, DT_F32: torch.float32}[dtype_flag]
    Bkv = 1 if kv_shared else B
    q = rnd(B, Lq, H * slot, seed=seed, scale=qscale, dtype=adt)
    k = rnd(Bkv, Lk, H * slot, seed=seed + 1, dtype=adt)
    v = rnd(Bkv, Lk, H * slot, seed=seed + 2, dtype=adt)
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    o_f32 = 1 if (dtype_flag == DT_F32 or nsplit > 1 or force_f32_out) else 0
    odt = torch.bfloat16 if dtype_flag == DT_F16 else adt  # the fp16 kernel still emits bf16 (GEMM A operand) or fp32
    o = torch.full((nsplit, B * Lq, H * d), float("nan"), device=DEV, dtype=torch.float32 if o_f32 else odt)
    lse = torch.full((nsplit, B, H, Lq), float("nan"), device=DEV)
    call("xs_flash_attn", P(q), P(k), P(v), P(o), P(lse), B, H, Lq, Lk, d, slot, H * slot, Lq * H * slot, H * slot,
         Lk * H * slot, int(kv_shared), nsplit, o_f32, scale, dtype_flag, st())
    if nsplit > 1:
        mdt = torch.float32 if force_f32_out else odt
        merged = torch.empty(B * Lq, H * d, device=DEV, dtype=mdt)
        lse_m = torch.empty(B, H, Lq, device=DEV)
        call("xs_lse_merge", P(o), P(lse), P(merged), P(lse_m), nsplit, B, Lq, H, d, 0, 0,
             DT_F32 if mdt == torch.float32 else DT_BF16, st())
        o, lse = merged, lse_m
    else:
        o, lse = o[0], lse[0]
    torch.cuda.synchronize()
    qd = q.double().view(B, Lq, H, slot)[..., :d].transpose(1, 2)
    kd = k.double().view(Bkv, Lk, H, slot)[..., :d].transpose(1, 2).expand(B, -1, -1, -1)
    vd = v.double().view(Bkv, Lk, H, slot)[..., :d].transpose(1, 2).expand(B, -1, -1, -1)
    ref, lse_ref = attn_ref(qd, kd, vd, scale)
    ref = ref.transpose(1, 2).reshape(B * Lq, H * d)
    return o.double(), ref, lse.double(), lse_ref


@pytest.mark.parametrize("B,H,Lq,Lk,d,slot,nsplit,shared", [
    (2, 6, 150, 150, 64, 64, 1, False),
    (1, 8, 130, 333, 48, 48, 1, False),
    (2, 8, 70, 1000, 48, 48, 3, False),
    (3, 8, 65, 200, 48, 48, 1, True),
])
def test_flash_attn_f32(B, H, Lq, Lk, d, slot, nsplit, shared):
    o, ref, lse, lse_ref = run_attn(DT_F32, B, H, Lq, Lk, d, slot, nsplit, shared)
    assert (o - ref).abs().max() < 2e-5
    assert (lse - lse_ref).abs().max() < 2e-5


@pytest.mark.parametrize("B,H,Lq,Lk,d,nsplit,shared,qscale", [
    (1, 1, 128, 128, 64, 1, False, 1.0),     # single tile
    (1, 2, 128, 256, 64, 1, False, 1.0),     # two kv blocks (stage ring, accumulate)
    (2, 6, 1370, 1370, 64, 1, False, 1.0),   # DINOv2 shape, ragged tails
    (1, 8, 128, 128, 48, 1, False, 1.0),     # d=48 single tile
    (2, 8, 1369, 1369, 48, 1, False, 2.0),   # decoder self-attention
    (1, 8, 300, 6845, 48, 1, False, 3.0),    # cross-attention, long kv, peaky softmax (lazy rescale path)
    (1, 8, 300, 6845, 48, 4, False, 1.0),    # split-KV + merge
    (3, 8, 200, 700, 48, 1, True, 1.0),      # shared reference K/V
    (2, 3, 1370, 700, 64, 1, False, 1.0),    # odd tile count AND odd head count (pair layout: mixed units, a lone tile)
    (2, 5, 300, 900, 48, 2, False, 1.0),     # the same with split-KV
    (1, 1, 1, 1, 64, 1, False, 1.0),         # one query, one key
    (2, 3, 5, 77, 48, 1, False, 1.0),        # a few rows, less than one key block
    (1, 8, 129, 127, 48, 1, False, 1.0),     # one row into the second tile, one key short of a block
    (3, 1, 257, 385, 64, 1, True, 1.0),      # single head: every odd tile is a lone unit
])
def test_flash_attn_bf16_tc(B, H, Lq, Lk, d, nsplit, shared, qscale, layout):
    o, ref, lse, lse_ref = run_attn(DT_BF16, B, H, Lq, Lk, d, 64, nsplit, shared, qscale=qscale)
    assert torch.isfinite(o).all()
    err = (o - ref).abs()
    assert err.max() < 0.03, f"max err {err.max().item()}"
    assert err.mean() < 3e-3
    assert (lse - lse_ref).abs().max() < 0.02


@pytest.mark.parametrize("B,H,Lq,Lk,d,nsplit,shared,qscale", [
    (1, 2, 128, 256, 64, 1, False, 1.0),
    (2, 6, 1370, 1370, 64, 1, False, 1.0),
    (1, 8, 300, 6845, 48, 1, False, 3.0),
    (1, 8, 300, 6845, 48, 4, False, 1.0),
    (3, 8, 200, 700, 48, 1, True, 1.0),
    (2, 3, 1370, 700, 64, 1, False, 1.0),
])
def test_flash_attn_bf16_tc_online_pass(B, H, Lq, Lk, d, nsplit, shared, qscale, layout):
    """xs_attn_set_optimistic(0): every tile goes through the second (online softmax, per-block row max, O rescale)
    pass -- the path that otherwise only runs for tiles whose row sums left the safe range."""
    _lib.load().xs_attn_set_optimistic(0)
    try:
        o, ref, lse, lse_ref = run_attn(DT_BF16, B, H, Lq, Lk, d, 64, nsplit, shared, qscale=qscale)
    finally:
        _lib.load().xs_attn_set_optimistic(1)
    assert torch.isfinite(o).all()
    err = (o - ref).abs()
    assert err.max() < 0.03 and err.mean() < 3e-3
    assert (lse - lse_ref).abs().max() < 0.02


def _attn_case(q, k, v, scale, d, nsplit=1):
    """q (B,Lq,H*64), k/v (B,Lk,H*64) bf16 -> (o, ref, lse, lse_ref) in float64; the reference uses the rounded inputs."""
    B, Lq, W = q.shape
    Lk, H = k.shape[1], W // 64
    o_f32 = 1 if nsplit > 1 else 0
    o = torch.full((nsplit, B * Lq, H * d), float("nan"), device=DEV, dtype=torch.float32 if o_f32 else torch.bfloat16)
    lse = torch.full((nsplit, B, H, Lq), float("nan"), device=DEV)
    call("xs_flash_attn", P(q), P(k), P(v), P(o), P(lse), B, H, Lq, Lk, d, 64, W, Lq * W, W, Lk * W, 0, nsplit, o_f32,
         scale, DT_BF16, st())
    if nsplit > 1:
        merged = torch.empty(B * Lq, H * d, device=DEV, dtype=torch.bfloat16)
        lse_m = torch.empty(B, H, Lq, device=DEV)
        call("xs_lse_merge", P(o), P(lse), P(merged), P(lse_m), nsplit, B, Lq, H, d, 0, 0, DT_BF16, st())
        o, lse = merged, lse_m
    else:
        o, lse = o[0], lse[0]
    torch.cuda.synchronize()
    qd = q.double().view(B, Lq, H, 64)[..., :d].transpose(1, 2)
    kd = k.double().view(B, Lk, H, 64)[..., :d].transpose(1, 2)
    vd = v.double().view(B, Lk, H, 64)[..., :d].transpose(1, 2)
    ref, lse_ref = attn_ref(qd, kd, vd, scale)
    return o.double(), ref.transpose(1, 2).reshape(B * Lq, H * d), lse.double(), lse_ref


def _ramp_inputs(B, H, Lq, Lk, d, key_logit, seed=0):
    """Inputs whose logits are (small noise) + key_logit[j] for every query row: feature 0 of q is 1, feature 0 of
    k carries the wanted per-key offset (values exactly representable in bf16 when multiples of 0.5 below 256)."""
    q = rnd(B, Lq, H * 64, seed=seed, scale=0.3, dtype=torch.bfloat16)
    k = rnd(B, Lk, H * 64, seed=seed + 1, scale=0.3, dtype=torch.bfloat16)
    v = rnd(B, Lk, H * 64, seed=seed + 2, dtype=torch.bfloat16)
    q.view(B, Lq, H, 64)[..., 0] = 1.0
    k.view(B, Lk, H, 64)[..., 0] = key_logit.to(DEV).to(torch.bfloat16)[None, :, None]
    return q, k, v


@pytest.mark.parametrize("d", [64, 48])
@pytest.mark.parametrize("case", ["rising_small", "rising_large", "late_spike", "early_spike", "all_large",
                                  "all_very_negative", "falling_large"])
def test_flash_attn_bf16_tc_adversarial_max(case, d, layout):
    """The softmax reference point.  The first pass takes 2^logit with no maximum at all; rows whose sums leave
    [2^-80, 2^100] must be caught and redone with the online softmax: a row maximum that rises in every one of >= 100
    key blocks, a +30 spike in the last block, logits that are all huge or all hugely negative."""
    Lq, Lk = 200, 100 * 64 + 1  # 101 blocks, tail of ONE key
    j = torch.arange(Lk, dtype=torch.float32)
    blk = torch.div(j, 64, rounding_mode="floor")
    ramp = {
        "rising_small": blk * 0.5,              # +0.5 per block: max rises 50 over the row (stays optimistic)
        "rising_large": blk * 2.0,              # +2 per block: 200 over the row -> 2^(288) overflows -> redo
        "late_spike": torch.where(j == Lk - 1, 30.0, 0.0),
        "early_spike": torch.where(j == 0, 150.0, 0.0),
        "all_large": torch.full((Lk,), 120.0),   # every logit ~ +120: P overflows without a reference
        "all_very_negative": torch.full((Lk,), -120.0),
        "falling_large": -blk * 2.0,
    }[case]
    q, k, v = _ramp_inputs(1, 2, Lq, Lk, d, ramp)
    o, ref, lse, lse_ref = _attn_case(q, k, v, 1.0, d)
    assert torch.isfinite(o).all(), case
    err = (o - ref).abs()
    assert err.max() < 0.03 and err.mean() < 3e-3, f"{case}: {err.max().item()} {err.mean().item()}"
    assert (lse - lse_ref).abs().max() < 0.05, case


def test_flash_attn_bf16_tc_long_kv_rows_mixed(layout):
    """Lk = 87 616 (cfg 4 / cfg 5 key count), split and unsplit; half of the query rows see a rising maximum that
    overflows the optimistic pass, the other half stay benign, so redone and first-pass tiles mix in one launch."""
    Lq, Lk, d, H = 256, 87616, 48, 8
    j = torch.arange(Lk, dtype=torch.float32)
    q, k, v = _ramp_inputs(1, H, Lq, Lk, d, j * (300.0 / Lk))
    q.view(1, Lq, H, 64)[:, :128, :, 0] = 0.0  # first query tile: no ramp
    for nsplit in (1, 8):
        o, ref, lse, lse_ref = _attn_case(q, k, v, 1.0 / math.sqrt(d), d, nsplit=nsplit)
        assert torch.isfinite(o).all()
        err = (o - ref).abs()
        assert err.max() < 0.03 and err.mean() < 3e-3, f"nsplit {nsplit}: {err.max().item()}"
        assert (lse - lse_ref).abs().max() < 0.05


def test_gemm_f16_out():
    """fp16 GEMM outputs (q|k|v operands of the fp16 attention kernel): bf16 x bf16 and tf32 x tf32 inputs."""
    for M, N, K in [(1000, 1152, 384), (40000, 1152, 384), (5000, 2048, 384)]:
        A = rnd(M, K, seed=1, dtype=torch.bfloat16)
        W = rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
        b = rnd(N, seed=3)
        out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float16)
        call("xs_gemm_bias_act", P(A), K, P(W), K, P(b), P(out), N, M, N, K, ACT_NONE, DT_BF16, DT_F16, st())
        torch.cuda.synchronize()
        ref = A.float() @ W.float().T + b
        assert ((out.float() - ref).abs() <= 2e-3 + 1e-3 * ref.abs()).all()
    A, W, b = rnd(1369, 384, seed=1), rnd(1536, 384, seed=2, scale=0.05), rnd(1536, seed=3)
    out = torch.full((1369, 1536), float("nan"), device=DEV, dtype=torch.float16)
    call("xs_gemm_bias_act", P(A), 384, P(W), 384, P(b), P(out), 1536, 1369, 1536, 384, ACT_NONE, DT_TF32, DT_F16, st())
    torch.cuda.synchronize()
    ref = A.double() @ W.double().T + b.double()
    assert ((out.double() - ref).abs() <= 8e-3 + 2e-3 * ref.abs()).all()


@pytest.mark.parametrize("nsplit", [1, 2])
def test_flash_attn_bf16_tc_f32_out(nsplit, layout):
    """decoder configuration: bf16 q/k/v, fp32 attention output (feeds the TF32 out-proj GEMM)"""
    o, ref, lse, lse_ref = run_attn(DT_BF16, 2, 8, 300, 1400, 48, 64, nsplit, False, force_f32_out=True)
    err = (o - ref).abs()
    assert err.max() < 0.03 and err.mean() < 3e-3


# ------------------------------------------------------------------------------------------------
# row kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [DT_F32, DT_BF16])
def test_layernorm_residual(dt):
    adt = torch.float32 if dt == DT_F32 else torch.bfloat16
    rows = 1001
    h, d = rnd(rows, 384, seed=1, scale=3.0), rnd(rows, 384, seed=2, dtype=adt)
    g, b = rnd(384, seed=3) * 0.2 + 1, rnd(384, seed=4, scale=0.1)
    h_out, y, y32 = torch.empty_like(h), torch.empty(rows, 384, device=DEV, dtype=adt), torch.empty_like(h)
    call("xs_layernorm", P(h), P(d), P(h_out), P(g), P(b), 1e-6, P(y), P(y32), rows, dt, st())
    x = h.double() + d.double()
    ref = torch.nn.functional.layer_norm(x, (384,), g.double(), b.double(), 1e-6)
    assert (h_out.double() - x).abs().max() < 1e-6
    assert (y32.double() - ref).abs().max() < 1e-5
    assert (y.double() - ref).abs().max() < (1e-5 if dt == DT_F32 else 0.04)
    # no residual (do_short_cut=False), in-place y32 over res_in allowed
    call("xs_layernorm", None, P(d), None, P(g), P(b), 1e-5, P(y), None, rows, dt, st())
    ref2 = torch.nn.functional.layer_norm(d.double(), (384,), g.double(), b.double(), 1e-5)
    assert (y.double() - ref2).abs().max() < (1e-5 if dt == DT_F32 else 0.04)


@pytest.mark.parametrize("W", [117, 126])  # odd width: scalar staging loads; even: two pixels per load
@pytest.mark.parametrize("pe_dt", [DT_F32, DT_TF32, DT_BF16])
def test_patch_embed(pe_dt, W):
    I, H = 3, 84
    ph, pw = H // 14, W // 14
    Pn = ph * pw
    img = rnd(I, 3, H, W, seed=5)
    wconv = rnd(384, 3, 14, 14, seed=6, scale=0.05)
    bias = rnd(384, seed=7, scale=0.1)
    tdt = torch.bfloat16 if pe_dt == DT_BF16 else torch.float32
    K = 592 if pe_dt == DT_BF16 else 588
    w2 = wconv.reshape(384, 588)
    if pe_dt == DT_BF16:
        w2 = torch.nn.functional.pad(w2, (0, 4))
    w2 = w2.to(tdt).contiguous()
    nbytes = _lib.load().xs_workspace_bytes(_lib.OP_PATCH_EMBED, I, H, W, pe_dt)
    assert nbytes == I * Pn * K * (2 if pe_dt == DT_BF16 else 4)
    ws = torch.empty(nbytes, device=DEV, dtype=torch.uint8)
    tok = torch.full((I * Pn, 384), float("nan"), device=DEV, dtype=tdt)
    call("xs_patch_embed", P(img), P(w2), P(bias), P(tok), P(ws), nbytes, I, H, W, pe_dt, st())
    ref = torch.nn.functional.conv2d(img.double(), wconv.double(), bias.double(), stride=14).flatten(2).transpose(1, 2)
    tol = {DT_F32: 1e-4, DT_TF32: 6e-3, DT_BF16: 0.06}[pe_dt]
    assert (tok.double().view(I, Pn, 384) - ref).abs().max() < tol


@pytest.mark.parametrize("dt", [DT_F32, DT_BF16])
def test_embed_and_final_ln(dt):
    adt = torch.float32 if dt == DT_F32 else torch.bfloat16
    I, Pn = 3, 48
    tok = rnd(I * Pn, 384, seed=10)

    # embeddings + first LN
    cls, pos = rnd(384, seed=8), rnd(Pn + 1, 384, seed=9, scale=0.3)
    g, b = rnd(384, seed=3) * 0.2 + 1, rnd(384, seed=4, scale=0.1)
    h = torch.empty(I * (Pn + 1), 384, device=DEV)
    y = torch.empty(I * (Pn + 1), 384, device=DEV, dtype=adt)
    call("xs_embed_cls_pos_ln", P(tok), DT_F32, P(cls), P(pos), P(h), P(g), P(b), 1e-6, P(y), I, Pn, dt, st())
    href = torch.cat([cls.double().expand(I, 1, 384), tok.double().view(I, Pn, 384)], 1) + pos.double()[None]
    assert (h.double().view(I, Pn + 1, 384) - href).abs().max() < 1e-5
    yref = torch.nn.functional.layer_norm(href, (384,), g.double(), b.double(), 1e-6)
    assert (y.double().view(I, Pn + 1, 384) - yref).abs().max() < (1e-4 if dt == DT_F32 else 0.05)

    # final LN + CLS drop + split + PE  (image 0 = query, images 1..2 = references)
    d = rnd(I * (Pn + 1), 384, seed=11, dtype=adt)
    pe = rnd(Pn, 384, seed=12)
    xq32 = torch.empty(Pn, 384, device=DEV)
    xq = torch.empty(Pn, 384, device=DEV, dtype=adt)
    mem = torch.empty(2 * Pn, 384, device=DEV, dtype=adt)
    call("xs_final_ln_drop_cls_add_pe", P(h), P(d), P(g), P(b), 1e-6, P(pe), P(xq32), P(xq), P(mem), I, 1, Pn, dt, st())
    f = torch.nn.functional.layer_norm(h.double() + d.double(), (384,), g.double(), b.double(), 1e-6)
    f = f.view(I, Pn + 1, 384)[:, 1:] + pe.double()[None]
    assert (xq32.double() - f[0]).abs().max() < 1e-4
    tol = 1e-4 if dt == DT_F32 else 0.05
    assert (xq.double() - f[0]).abs().max() < tol
    assert (mem.double().view(2, Pn, 384) - f[1:]).abs().max() < tol


def test_table_resamplers():
    from oracle import crossscore_oracle as O
    t = rnd(40, 40, 384, seed=1)
    for oh, ow in [(37, 37), (5, 5), (6, 8), (74, 74), (37, 49)]:
        out = torch.empty(oh, ow, 384, device=DEV)
        call("xs_pe_resample_bilinear_ac", P(t), P(out), 40, 40, oh, ow, 384, st())
        ref = O.bilinear_resize_ac_true(t.cpu().double(), oh, ow)
        assert (out.cpu().double() - ref).abs().max() < 2e-5
    t = rnd(37, 37, 384, seed=2, scale=0.3)
    for oh, ow in [(5, 5), (6, 8), (74, 74), (37, 49), (12, 11)]:
        out = torch.empty(oh, ow, 384, device=DEV)
        call("xs_pos_embed_resample_bicubic", P(t), P(out), 37, 37, oh, ow, 384, st())
        ref = O.bicubic_resize_ac_false(t.cpu().double(), oh, ow)
        assert (out.cpu().double() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("dt", [DT_F32, DT_BF16])
@pytest.mark.parametrize("use_tanh,power", [(0, 1.0), (1, 1.0), (0, 2.0), (0, 0.5)])
def test_head_score_jigsaw(dt, use_tanh, power):
    adt = torch.float32 if dt == DT_F32 else torch.bfloat16
    B, ph, pw = 2, 9, 17  # 153 tokens per map: straddles the 128-row tile, crosses the batch boundary
    R = B * ph * pw
    A = rnd(R, 384, seed=1, dtype=adt)
    W = rnd(196, 384, seed=2, scale=0.1)
    bias = rnd(196, seed=3, scale=0.05)
    Wk, bk = W, bias
    if dt == DT_BF16:
        Wk = torch.nn.functional.pad(W, (0, 0, 0, 28))
        bk = torch.nn.functional.pad(bias, (0, 28)).contiguous()
    Wk = Wk.to(adt).contiguous()
    score = torch.full((B, 14 * ph, 14 * pw), float("nan"), device=DEV)
    call("xs_head_score_jigsaw", P(A), 384, P(Wk), 384, P(bk), P(score), B, ph, pw, 384, use_tanh, power, dt, st())
    z = A.double() @ Wk.double()[:196].T + bias.double()
    s = torch.tanh(z) if use_tanh else torch.sigmoid(z)
    if power != 1.0:
        s = s ** power
    ref = s.view(B, ph, pw, 14, 14).permute(0, 1, 3, 2, 4).reshape(B, 14 * ph, 14 * pw)
    tol = 1e-5 if dt == DT_F32 else 0.02
    assert torch.isfinite(score).all()
    assert (score.double() - ref).abs().max() < tol


@pytest.mark.parametrize("dt", [DT_F32, DT_BF16, DT_F16])
def test_attn_probs_one_head(dt):
    adt = {DT_F32: torch.float32, DT_BF16: torch.bfloat16, DT_F16: torch.float16}[dt]
    B, H, Lq, Lk, d = 2, 8, 50, 120, 48
    slot = 48 if dt == DT_F32 else 64
    q, k = rnd(B, Lq, H * slot, seed=1, dtype=adt), rnd(B, Lk, H * slot, seed=2, dtype=adt)
    scale = 1 / math.sqrt(d)
    head = 5
    qd = q.double().view(B, Lq, H, slot)[:, :, head, :d]
    kd = k.double().view(B, Lk, H, slot)[:, :, head, :d]
    s = qd @ kd.transpose(-1, -2) * scale
    lse = torch.zeros(B, H, Lq, device=DEV)
    lse[:, head] = torch.logsumexp(s, -1).float()
    probs = torch.empty(B, Lq, Lk, device=DEV)
    call("xs_attn_probs_one_head", P(q), P(k), P(lse), P(probs), B, H, head, Lq, Lk, d, slot, H * slot, Lq * H * slot,
         H * slot, Lk * H * slot, scale, dt, st())
    assert (probs.double() - torch.softmax(s, -1)).abs().max() < 1e-5


def test_gemm_bias_residual_rejects_bad_shapes():
    A = rnd(128, 384, dtype=torch.bfloat16)
    W = rnd(200, 384, dtype=torch.bfloat16)
    b, h = rnd(200), rnd(128, 200)
    with pytest.raises(_lib.XsError):      # N must be a multiple of 192
        call("xs_gemm_bias_residual", P(A), 384, P(W), 384, P(b), P(h), 200, 128, 200, 384, DT_BF16, st())


def test_flash_attn_bf16_tc_redo_in_mixed_units(layout):
    """Odd tile count and odd head count: in the pair layout the last tiles of heads 0 and 1 share a unit (each with its
    own K/V stream) and head 2's last tile is alone in one.  Only head 1 carries a ramp that overflows the max-free pass,
    so redone and first-pass tiles meet inside one mixed unit, and a lone-tile unit runs next to them."""
    B, H, Lq, Lk, d = 2, 3, 384, 1000, 64
    j = torch.arange(Lk, dtype=torch.float32)
    q, k, v = _ramp_inputs(B, H, Lq, Lk, d, j * 0.0)
    k.view(B, Lk, H, 64)[:, :, 1, 0] = (j * (300.0 / Lk)).to(DEV).to(torch.bfloat16)[None, :]   # head 1 only
    o, ref, lse, lse_ref = _attn_case(q, k, v, 1.0, d)
    assert torch.isfinite(o).all()
    err = (o - ref).abs()
    assert err.max() < 0.03 and err.mean() < 3e-3, (err.max().item(), err.mean().item())
    assert (lse - lse_ref).abs().max() < 0.05
