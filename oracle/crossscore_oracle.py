"""CPU oracle for the CrossScore inference hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-math restatement (torch CPU tensors, fp64 by default) of the
reference's forward path.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs, always as the checker, never as the product: nothing under ``crossscore_b200/``
imports it, and the product raises if its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the oracle
is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports
``/root/reference`` (task/core.py::CrossScoreNet, unmodified) in the build container, runs
it on seeded weights/inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.

Each function cites the reference code it restates (paths relative to /root/reference,
``$SP`` = site-packages of the pinned third-party libraries: transformers Dinov2Model and
torch.nn.MultiheadAttention).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

# FAST = True switches the three heaviest primitives to torch's fused CPU ops (F.scaled_dot_product_attention,
# F.layer_norm, F.gelu) -- the very library calls the reference makes on CPU -- so that the timed CPU
# baseline reflects the reference's real CPU speed rather than the spelled-out restatement's.  The
# spelled-out path (FAST = False, the default) is the parity checker; tests pin both against the goldens.
FAST = False

PATCH = 14
HIDDEN = 384
DINO_HEADS = 6
DINO_LAYERS = 12
DEC_HEADS = 8
DEC_LAYERS = 2
DINO_EPS = 1e-6
DEC_EPS = 1e-5
DINO_GRID = 37  # 518 // 14, the pre-trained position-embedding grid


# ----------------------------------------------------------------------------------------
# small math helpers (no torch.nn.functional high-level ops: everything is spelled out)
# ----------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    """LayerNorm over the last dim, biased variance (torch.nn.LayerNorm semantics)."""
    if FAST:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    if FAST:
        return torch.nn.functional.linear(x, w, b)  # fused bias (addmm), as nn.Linear does
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Exact GELU ($SP/transformers/models/dinov2/modeling_dinov2.py:312-328, hidden_act="gelu")."""
    if FAST:
        return torch.nn.functional.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def softmax_rows(s: torch.Tensor) -> torch.Tensor:
    s = s - s.max(dim=-1, keepdim=True).values
    e = torch.exp(s)
    return e / e.sum(dim=-1, keepdim=True)


def mha(q_in, k_in, v_in, w_in, b_in, w_out, b_out, n_heads, return_probs=False):
    """torch.nn.MultiheadAttention forward ($SP/torch/nn/functional.py:5849-5858 packed
    in-proj split, :6630-6697 scaled dot-product + out-proj).  batch_first layout."""
    E = q_in.shape[-1]
    d = E // n_heads
    wq, wk, wv = w_in[:E], w_in[E:2 * E], w_in[2 * E:]
    bq, bk, bv = b_in[:E], b_in[E:2 * E], b_in[2 * E:]
    q = linear(q_in, wq, bq)
    k = linear(k_in, wk, bk)
    v = linear(v_in, wv, bv)
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(B, Lq, n_heads, d).transpose(1, 2)
    k = k.view(B, Lk, n_heads, d).transpose(1, 2)
    v = v.view(B, Lk, n_heads, d).transpose(1, 2)
    if FAST and not return_probs:
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, Lq, E)
        return linear(o, w_out, b_out), None
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    p = softmax_rows(s)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, E)
    out = linear(o, w_out, b_out)
    return (out, p) if return_probs else (out, None)


# ----------------------------------------------------------------------------------------
# resampling of the two position tables
# ----------------------------------------------------------------------------------------
def _cubic_coeffs(t: torch.Tensor, A: float = -0.75):
    """Cubic-convolution weights (Keys, A=-0.75) as used by ATen upsample_bicubic2d."""
    def c1(x):  # |x| <= 1
        return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0

    def c2(x):  # 1 < |x| < 2
        return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A

    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def bicubic_resize_ac_false(table: torch.Tensor, oh: int, ow: int, step_h=None, step_w=None) -> torch.Tensor:
    """(ih, iw, C) -> (oh, ow, C), bicubic, align_corners=False, border-clamped taps.
    Restates F.interpolate(mode="bicubic", align_corners=False, size=...) used by
    $SP/transformers/models/dinov2/modeling_dinov2.py:86-91 (computed in fp32 there).
    step_h / step_w: source step per output pixel when the caller passed scale_factor= instead of size=
    (ATen area_pixel_compute_scale: 1 / scale_factor); None = ih / oh."""
    ih, iw, _ = table.shape
    dt = table.dtype

    def axis(o, i, step):
        scale = i / o if step is None else step
        src = (torch.arange(o, dtype=dt, device=table.device) + 0.5) * scale - 0.5
        fl = torch.floor(src)
        t = src - fl
        idx = fl.long()
        taps = [torch.clamp(idx + k, 0, i - 1) for k in (-1, 0, 1, 2)]
        return taps, _cubic_coeffs(t)

    ty, wy = axis(oh, ih, step_h)
    tx, wx = axis(ow, iw, step_w)
    out = torch.zeros(oh, ow, table.shape[2], dtype=dt, device=table.device)
    for a in range(4):
        rows = table[ty[a]]  # (oh, iw, C)
        acc = torch.zeros(oh, ow, table.shape[2], dtype=dt, device=table.device)
        for b in range(4):
            acc = acc + rows[:, tx[b]] * wx[b][None, :, None]
        out = out + acc * wy[a][:, None, None]
    return out


def bilinear_resize_ac_true(table: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """(ih, iw, C) -> (oh, ow, C), bilinear, align_corners=True.
    model/positional_encoding.py:61-69 calls F.interpolate with
    scale_factor=((oh+1e-4)/ih, (ow+1e-4)/iw) and align_corners=True; the output size is
    floor(ih*scale)=oh and with align_corners=True the source coordinate is
    dst*(ih-1)/(oh-1) irrespective of the scale factor."""
    ih, iw, _ = table.shape
    dt = table.dtype

    def axis(o, i):
        if o > 1:
            src = torch.arange(o, dtype=dt, device=table.device) * ((i - 1) / (o - 1))
        else:
            src = torch.zeros(o, dtype=dt, device=table.device)
        i0 = torch.clamp(torch.floor(src).long(), 0, i - 1)
        i1 = torch.clamp(i0 + 1, 0, i - 1)
        t = src - i0.to(dt)
        return i0, i1, t

    y0, y1, ty = axis(oh, ih)
    x0, x1, tx = axis(ow, iw)
    top = table[y0][:, x0] * (1 - tx)[None, :, None] + table[y0][:, x1] * tx[None, :, None]
    bot = table[y1][:, x0] * (1 - tx)[None, :, None] + table[y1][:, x1] * tx[None, :, None]
    return top * (1 - ty)[:, None, None] + bot * ty[:, None, None]


# How Dinov2Embeddings.interpolate_pos_encoding resamples the 37x37 position grid for inputs other than 518x518:
#   "scale_factor": transformers 4.33.3, the version the reference pins (environment.yaml:340) --
#       F.interpolate(scale_factor=((ph + 0.1) / 37, (pw + 0.1) / 37), mode="bicubic", align_corners=False);
#       with an explicit scale_factor ATen steps the source coordinate by 1 / scale_factor = 37 / (ph + 0.1).
#       (4.33.3 is not installed here: this follows its published source, and tests pin the arithmetic against
#       F.interpolate(scale_factor=...) itself.)
#   "size": transformers >= 4.4x, incl. the 5.5.0 of this container ($SP/.../modeling_dinov2.py:86-91) --
#       F.interpolate(size=(ph, pw)), i.e. a step of 37 / ph.
POS_INTERP = "scale_factor"


def dinov2_pos_table(sd: Dict[str, torch.Tensor], H: int, W: int, dt, pos_interp: Optional[str] = None) -> torch.Tensor:
    """Dinov2Embeddings.interpolate_pos_encoding ($SP/.../modeling_dinov2.py:57-95)."""
    mode = pos_interp or POS_INTERP
    assert mode in ("scale_factor", "size"), mode
    pos = sd["backbone.embeddings.position_embeddings"][0]  # (1+37*37, C)
    ph, pw = H // PATCH, W // PATCH
    if ph * pw == pos.shape[0] - 1 and H == W:
        return pos.to(dt)
    g = int(round(math.sqrt(pos.shape[0] - 1)))
    steps = (None, None) if mode == "size" else (1.0 / ((ph + 0.1) / g), 1.0 / ((pw + 0.1) / g))
    # the library interpolates in fp32 whatever the model dtype (:86-91)
    grid = bicubic_resize_ac_false(pos[1:].reshape(g, g, -1).to(torch.float32).to(dt), ph, pw, *steps)
    return torch.cat([pos[:1].to(dt), grid.reshape(ph * pw, -1)], dim=0)


def multiview_pe_table(sd: Dict[str, torch.Tensor], H: int, W: int, dt) -> torch.Tensor:
    """MultiViewPosionalEmbeddings (model/positional_encoding.py:42-75): (ph*pw, C) table."""
    PE = sd["pos_enc_fn.PE"][0].to(dt)  # (40, 40, C)
    ph, pw = H // PATCH, W // PATCH
    if ph == PE.shape[0] and pw == PE.shape[1]:
        return PE.reshape(ph * pw, -1)
    return bilinear_resize_ac_true(PE, ph, pw).reshape(ph * pw, -1)


# ----------------------------------------------------------------------------------------
# the path
# ----------------------------------------------------------------------------------------
def dinov2_features(sd: Dict[str, torch.Tensor], imgs: torch.Tensor, dt=torch.float64, pos_interp=None) -> torch.Tensor:
    """Dinov2Model(pixel_values).last_hidden_state: (I,3,H,W) -> (I, 1+P, C).
    $SP/transformers/models/dinov2/modeling_dinov2.py:97-116 (embeddings), :141-149 (patch
    conv k=s=14), :203-234 (attention, 6 heads x 64, scale 1/8), :249-252 (out dense),
    :324-328 (MLP fc1 -> GELU -> fc2), :367-386 (pre-norm layer with LayerScale),
    :473-477 (final LayerNorm)."""
    g = lambda k: sd["backbone." + k].to(dt)
    I, _, H, W = imgs.shape
    ph, pw = H // PATCH, W // PATCH
    x = imgs.to(dt)[:, :, : ph * PATCH, : pw * PATCH]
    # conv k=s=14  ==  per-patch dot with the (384, 3*14*14) kernel matrix
    patches = x.reshape(I, 3, ph, PATCH, pw, PATCH).permute(0, 2, 4, 1, 3, 5).reshape(I, ph * pw, 3 * PATCH * PATCH)
    wpe = g("embeddings.patch_embeddings.projection.weight").reshape(HIDDEN, -1)
    tok = linear(patches, wpe, g("embeddings.patch_embeddings.projection.bias"))
    cls = g("embeddings.cls_token").expand(I, -1, -1)
    h = torch.cat([cls, tok], dim=1) + dinov2_pos_table(sd, H, W, dt, pos_interp)[None]
    d = HIDDEN // DINO_HEADS
    for l in range(DINO_LAYERS):
        p = f"encoder.layer.{l}."
        y = layer_norm(h, g(p + "norm1.weight"), g(p + "norm1.bias"), DINO_EPS)
        q = linear(y, g(p + "attention.attention.query.weight"), g(p + "attention.attention.query.bias"))
        k = linear(y, g(p + "attention.attention.key.weight"), g(p + "attention.attention.key.bias"))
        v = linear(y, g(p + "attention.attention.value.weight"), g(p + "attention.attention.value.bias"))
        T = h.shape[1]
        q = q.view(I, T, DINO_HEADS, d).transpose(1, 2)
        k = k.view(I, T, DINO_HEADS, d).transpose(1, 2)
        v = v.view(I, T, DINO_HEADS, d).transpose(1, 2)
        if FAST:
            a = torch.nn.functional.scaled_dot_product_attention(q, k, v)
        else:
            a = softmax_rows((q @ k.transpose(-1, -2)) / math.sqrt(d)) @ v
        a = a.transpose(1, 2).reshape(I, T, HIDDEN)
        a = linear(a, g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias"))
        h = h + g(p + "layer_scale1.lambda1") * a
        y = layer_norm(h, g(p + "norm2.weight"), g(p + "norm2.bias"), DINO_EPS)
        m = gelu_erf(linear(y, g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias")))
        m = linear(m, g(p + "mlp.fc2.weight"), g(p + "mlp.fc2.bias"))
        h = h + g(p + "layer_scale2.lambda1") * m
    return layer_norm(h, g("layernorm.weight"), g("layernorm.bias"), DINO_EPS)


def get_featmaps(sd, query_img, ref_imgs, dt=torch.float64, pos_interp=None):
    """CrossScoreNet.get_featmaps (task/core.py:119-161): one backbone pass over
    cat([query, refs]); drop CLS (:142); split query / refs (:146-153)."""
    B, _, H, W = query_img.shape
    N = ref_imgs.shape[1]
    allv = torch.cat([query_img[:, None], ref_imgs], dim=1).reshape(B * (1 + N), 3, H, W)
    f = dinov2_features(sd, allv, dt, pos_interp)[:, 1:]
    P = f.shape[1]
    f = f.view(B, 1 + N, P, HIDDEN)
    return f[:, 0], f[:, 1:].reshape(B, N * P, HIDDEN)


def decoder(sd, x, mem, do_self_attn=True, do_short_cut=True, head_id=0, need_probs=False, dt=torch.float64):
    """TransformerDecoderCustomised (model/customised_transformer/transformer.py:246-268)
    over 2 post-norm layers (:157-173): self-attn block :182-192, cross-attn block :195-205,
    FFN (ReLU) :208-210.  The same ``mem`` feeds both layers (:251-261)."""
    probs = None
    for l in range(DEC_LAYERS):
        p = f"ref_cross.attn.layers.{l}."
        g = lambda k: sd[p + k].to(dt)
        if do_self_attn:
            sa, _ = mha(x, x, x, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"),
                        g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"), DEC_HEADS)
            x = layer_norm((x + sa) if do_short_cut else sa, g("norm1.weight"), g("norm1.bias"), DEC_EPS)
        ca, pr = mha(x, mem, mem, g("multihead_attn.in_proj_weight"), g("multihead_attn.in_proj_bias"),
                     g("multihead_attn.out_proj.weight"), g("multihead_attn.out_proj.bias"), DEC_HEADS,
                     return_probs=need_probs)
        if need_probs:
            probs = pr[:, head_id]  # transformer.py:175-178, last layer wins (:266-268)
        x = layer_norm((x + ca) if do_short_cut else ca, g("norm2.weight"), g("norm2.bias"), DEC_EPS)
        ff = linear(torch.relu(linear(x, g("linear1.weight"), g("linear1.bias"))), g("linear2.weight"), g("linear2.bias"))
        x = layer_norm(x + ff, g("norm3.weight"), g("norm3.bias"), DEC_EPS)
    return x, probs


def resolve_power(metric_type="ssim", metric_min=0, power_factor="default") -> float:
    """RegressionLayer._get_pow_fn (model/regression_layer.py:40-62)."""
    if metric_min == 0:
        p = {"ssim": 1, "mae": 2, "mse": 4}[metric_type] if power_factor == "default" else power_factor
    else:
        p = 1
    return float(p)


def head_and_jigsaw(sd, x, ph, pw, metric_type="ssim", metric_min=0, power_factor="default", dt=torch.float64):
    """CrossReferenceNet.head (model/cross_reference.py:45-50,82), RegressionLayer
    (model/regression_layer.py:26-62), jigsaw_to_image (utils/misc/image.py:8-21)."""
    g = lambda k: sd["ref_cross.head." + k].to(dt)
    z = linear(x, g("0.weight"), g("0.bias"))
    z = torch.where(z >= 0, z, 0.01 * z)  # LeakyReLU default slope
    z = linear(z, g("2.weight"), g("2.bias"))
    if metric_min == -1:
        s = torch.tanh(z)
    elif metric_min == 0:
        s = 1.0 / (1.0 + torch.exp(-z))
    else:
        raise ValueError(f"metric_min={metric_min} not supported")
    p = resolve_power(metric_type, metric_min, power_factor)
    if p != 1.0:
        s = s ** p
    B = x.shape[0]
    s = s.view(B, ph, pw, PATCH, PATCH).permute(0, 1, 3, 2, 4).reshape(B, ph * PATCH, pw * PATCH)
    return s


def crossscore_forward(sd, query_img, ref_imgs, *, need_attn_weights=False, head_id=0,
                       do_self_attn=True, do_short_cut=True, metric_type="ssim", metric_min=0,
                       power_factor="default", dt=torch.float64, pos_interp=None):
    """CrossScoreNet.forward with norm_img=False (task/core.py:58-117) ->
    {"score_map_ref_cross": (B, 14*ph, 14*pw), "attn_weights_map_ref_cross": None | (B,ph,pw,N,ph,pw)}."""
    B, _, H, W = query_img.shape
    N = ref_imgs.shape[1]
    ph, pw = H // PATCH, W // PATCH
    fq, fr = get_featmaps(sd, query_img, ref_imgs, dt, pos_interp)
    pe = multiview_pe_table(sd, H, W, dt)
    fq = fq + pe[None]
    fr = (fr.view(B, N, ph * pw, HIDDEN) + pe[None, None]).reshape(B, N * ph * pw, HIDDEN)
    x, probs = decoder(sd, fq, fr, do_self_attn, do_short_cut, head_id, need_attn_weights, dt)
    score = head_and_jigsaw(sd, x, ph, pw, metric_type, metric_min, power_factor, dt)
    if probs is not None:
        probs = probs.reshape(B, ph, pw, N, ph, pw)
    return {"score_map_ref_cross": score, "attn_weights_map_ref_cross": probs,
            "_featmap_query": fq, "_featmap_ref": fr, "_decoder_out": x}


def lse_merge(o_parts, lse_parts):
    """Split-KV identity: O = sum_r exp(LSE_r - LSE) O_r."""
    lse = torch.logsumexp(torch.stack(lse_parts, 0), dim=0)
    o = sum(torch.exp(l - lse)[..., None] * o for o, l in zip(o_parts, lse_parts))
    return o, lse
