"""Host-side plan of the CrossScore forward: packed weights + the kernel sequence over the C ABI.

PyTorch is used for device memory (torch.empty), streams and the one-off weight re-layout; every
operation of the forward itself is a call into libcrossscore_sm100a.so (crossscore_b200._lib).

Data layout in HBM (C = 384, P = patches/image, T = P + 1, I = images in the batch, AT = bf16 | fp32):
  h      (I*T, C)   fp32   DINOv2 residual stream (kept fp32: outlier channels)
  y      (I*T, C)   AT     LayerNorm output = GEMM A operand
  qkv    (I*T, 3C)  AT     fused Q|K|V projection, head h at columns h*64 of each part
  att    (I*T, C)   AT     attention output, heads concatenated
  g      (I*T, 4C)  AT     MLP hidden (GELU fused in the fc1 epilogue)
  xq32/xq (B*P, C)  fp32/AT  decoder stream (post-norm: fp32 copy is the residual, AT copy feeds GEMMs)
  mem    (B*N*P, C) AT     reference tokens (+PE); layer-invariant, so K/V of BOTH decoder layers come
                           from ONE GEMM:  kv (B*N*P, 2 layers x (K|V) x 8 heads x slot)
  decoder heads are 48 wide; the bf16 path stores them in 64-column slots (weights zero-padded) so the
  attention kernel's TMA boxes are 128 bytes; QK^T runs 3 K-steps and PV runs N=48, no padded FLOPs.
"""
from __future__ import annotations

import math
import os
from contextlib import contextmanager
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, DT_BF16, DT_F16, DT_F32, DT_TF32, call

C = 384
PATCH = 14
DINO_HEADS, DINO_LAYERS, DINO_EPS = 6, 12, 1e-6
DEC_HEADS, DEC_LAYERS, DEC_EPS, DEC_D = 8, 2, 1e-5, 48
NUM_SMS_HINT = 148


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest fp32 -> TF32 (10-bit mantissa); the tensor core itself would truncate."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


ATTN_LAYOUT = 1  # the library's default kernel layout (xs_attn_set_layout): 1 = pair kernel, 0 = 64-key kernel


def set_attn_layout(layout: int) -> None:
    """Select the flash-attention kernel layout for the process (library switch + the key-split model below)."""
    global ATTN_LAYOUT
    if layout not in (0, 1):
        raise ValueError("attention layout must be 0 or 1")
    _lib.load().xs_attn_set_layout(int(layout))
    ATTN_LAYOUT = int(layout)


def attn_units(B: int, heads: int, Lq: int, layout: int = None):
    """(work units of one key range, persistent CTA slots, microseconds per 128-key block and unit) of a launch.

    Layout 1 (xs_attn_tc2.cu): one CTA per SM walks pairs of 128-row query tiles -- per (batch): heads * (tiles // 2) pairs
    plus, for an odd tile count, ceil(heads / 2) units made of the last tiles of two heads.  Layout 0 (xs_attn_tc.cu):
    two CTAs per SM, one 128-row tile each.  Block times measured stand-alone on cfg 2's DINOv2 shape (r2t)."""
    layout = ATTN_LAYOUT if layout is None else layout
    nq = (Lq + 127) // 128
    if layout == 1:
        return B * (heads * (nq // 2) + ((heads + 1) // 2 if nq & 1 else 0)), NUM_SMS_HINT, 1.4
    return B * heads * nq, 2 * NUM_SMS_HINT, 1.45


def attn_kv_splits(units: int, nblk: int, slots: int = NUM_SMS_HINT, t_blk: float = 1.4, max_splits: int = 8) -> int:
    """How many key ranges a flash-attention launch is cut into (partials merged by xs_lse_merge).

    The kernel is persistent with `slots` CTAs walking `units * nsplit` equally long work units (attn_units), so a launch
    takes ceil(units * nsplit / slots) waves of (unit time / nsplit).  Few units (a single query: 8 heads x 11 query
    tiles = 44 pair units) leave most SMs idle, and a unit count just above a multiple of `slots` pays a nearly empty last
    wave -- both are fixed by splitting the keys.  A split costs an fp32 partial-output round trip and the merge launch
    (~15 us, measured), so short launches (the DINOv2 attention of a handful of images: ~15 us per unit) are left
    alone.  `t_blk`: microseconds per 128-key block of one unit."""
    t_unit = t_blk * nblk  # microseconds
    best, best_t = 1, None
    for n in range(1, max(1, min(max_splits, nblk // 4)) + 1):
        per = -(-nblk // n)
        if -(-nblk // per) != n:  # every split must own at least one key block
            continue
        t = -(-units * n // slots) * t_unit / n + (15.0 if n > 1 else 0.0)
        if best_t is None or t < best_t * 0.97:  # prefer fewer splits on near-ties
            best, best_t = n, t
    return best


POS_INTERP_MODES = ("scale_factor", "size")


def pos_interp_steps(mode: str, g: int, ph: int, pw: int):
    """Source step per output pixel of the DINOv2 position-embedding resample (0.0 = g / out, the size= form).

    "scale_factor": transformers 4.33.3, the version the reference pins (environment.yaml:340), calls
    F.interpolate(scale_factor=((ph + 0.1) / g, (pw + 0.1) / g)); ATen then maps output to source coordinates with
    1 / scale_factor (computed in double, used in fp32).  "size": transformers >= 4.4x pass size=(ph, pw)
    ($SP/transformers/models/dinov2/modeling_dinov2.py:86-91), i.e. g / ph."""
    if mode == "size":
        return 0.0, 0.0
    return 1.0 / ((ph + 0.1) / g), 1.0 / ((pw + 0.1) / g)


class PackedWeights:
    """Kernel-friendly copies of the reference state_dict (built once per device / precision)."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, precision: str, do_self_attn: bool = True,
                 fold_qscale: bool = False, pos_interp: str = "scale_factor"):
        """fold_qscale: multiply the query projections by softmax_scale * log2(e) (in fp32, before rounding), so the
        attention kernel's logits are already in the log2 domain and its softmax needs no multiply per logit
        (the bf16 product mode; the fp32 parity mode keeps the reference's operation order)."""
        assert precision in ("bf16", "fp32")
        if pos_interp not in POS_INTERP_MODES:
            raise ValueError(f"pos_interp must be one of {POS_INTERP_MODES}, got {pos_interp!r}")
        self.pos_interp = pos_interp
        self.precision = precision
        self.fold_qscale = fold_qscale
        LOG2E = 1.4426950408889634
        qs_dino = LOG2E / math.sqrt(64.0) if fold_qscale else 1.0
        qs_dec = LOG2E / math.sqrt(float(DEC_D)) if fold_qscale else 1.0
        self.dt = DT_BF16 if precision == "bf16" else DT_F32
        self.wdtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.slot = 64 if precision == "bf16" else DEC_D  # decoder head slot width
        self.device = device
        f32 = lambda k: sd[k].detach().to(device=device, dtype=torch.float32).contiguous()
        W = lambda t: t.to(self.wdtype).contiguous()
        # weights of the TF32 GEMMs (patch embed, decoder projections / FFN, head) stay fp32, pre-rounded
        WP = (lambda t: round_tf32(t.float())) if precision == "bf16" else W
        b = "backbone."
        # --- patch embed: Conv2d weight (384,3,14,14) -> (384, 588) [c, ky, kx]
        wpe = f32(b + "embeddings.patch_embeddings.projection.weight").reshape(C, 588)
        self.w_pe, self.b_pe = WP(wpe), f32(b + "embeddings.patch_embeddings.projection.bias")
        self.cls = f32(b + "embeddings.cls_token").reshape(C)
        self.pos = f32(b + "embeddings.position_embeddings")[0].contiguous()  # (1+37*37, C)
        self.layers = []
        for l in range(DINO_LAYERS):
            p = f"{b}encoder.layer.{l}."
            lam1, lam2 = f32(p + "layer_scale1.lambda1"), f32(p + "layer_scale2.lambda1")
            L = dict(
                ln1_g=f32(p + "norm1.weight"), ln1_b=f32(p + "norm1.bias"),
                wqkv=W(torch.cat([f32(p + "attention.attention.query.weight") * qs_dino,
                                  f32(p + "attention.attention.key.weight"),
                                  f32(p + "attention.attention.value.weight")], 0)),
                bqkv=torch.cat([f32(p + "attention.attention.query.bias") * qs_dino,
                                f32(p + "attention.attention.key.bias"),
                                f32(p + "attention.attention.value.bias")], 0).contiguous(),
                # LayerScale folded into the producing Linear in fp32, before any rounding:
                #   h += lam * (a W^T + b)  ==  h += a (lam*W)^T + lam*b
                wo=W(f32(p + "attention.output.dense.weight") * lam1[:, None]),
                bo=(f32(p + "attention.output.dense.bias") * lam1).contiguous(),
                ln2_g=f32(p + "norm2.weight"), ln2_b=f32(p + "norm2.bias"),
                w1=W(f32(p + "mlp.fc1.weight")), b1=f32(p + "mlp.fc1.bias"),
                w2=W(f32(p + "mlp.fc2.weight") * lam2[:, None]), b2=(f32(p + "mlp.fc2.bias") * lam2).contiguous(),
            )
            if precision == "bf16":
                # LayerNorm folded into the Linear that follows it (xs_gemm_ln_folded): gamma into the weight (fp32 product,
                # then one bf16 rounding), c1 = row sums of the ROUNDED weight (the mean correction must match what the MMA
                # sums), c0 = W beta + b.  norm1 -> q|k|v, norm2 -> fc1 (modeling_dinov2.py:367-386).
                wq32 = torch.cat([f32(p + "attention.attention.query.weight") * qs_dino,
                                  f32(p + "attention.attention.key.weight"),
                                  f32(p + "attention.attention.value.weight")], 0)
                L["wqkv_f"] = W(wq32 * L["ln1_g"][None, :])
                L["c1_qkv"] = L["wqkv_f"].float().sum(-1).contiguous()
                L["c0_qkv"] = (wq32 @ L["ln1_b"] + L["bqkv"]).contiguous()
                w1_32 = f32(p + "mlp.fc1.weight")
                L["w1_f"] = W(w1_32 * L["ln2_g"][None, :])
                L["c1_fc1"] = L["w1_f"].float().sum(-1).contiguous()
                L["c0_fc1"] = (w1_32 @ L["ln2_b"] + L["b1"]).contiguous()
            self.layers.append(L)
        self.lnf_g, self.lnf_b = f32(b + "layernorm.weight"), f32(b + "layernorm.bias")
        self.pe_table = f32("pos_enc_fn.PE")[0].contiguous()  # (pe_h, pe_w, C)

        # --- decoder: per-head 48 -> slot padding of the packed in-proj rows
        slot, E = self.slot, DEC_HEADS * self.slot

        def pad_heads(w):  # (8*48, X) -> (8*slot, X)
            if slot == DEC_D:
                return w
            w = w.reshape(DEC_HEADS, DEC_D, -1)
            out = torch.zeros(DEC_HEADS, slot, w.shape[-1], device=w.device, dtype=w.dtype)
            out[:, :DEC_D] = w
            return out.reshape(DEC_HEADS * slot, -1)

        def pad_heads_b(bv):
            return pad_heads(bv[:, None])[:, 0]

        self.dec = []
        kv_w, kv_b = [], []
        for l in range(DEC_LAYERS):
            p = f"ref_cross.attn.layers.{l}."
            D = {}
            if do_self_attn:
                wi, bi = f32(p + "self_attn.in_proj_weight"), f32(p + "self_attn.in_proj_bias")
                qs = (qs_dec, 1.0, 1.0)
                D["sa_win"] = WP(torch.cat([pad_heads(wi[i * C:(i + 1) * C] * qs[i]) for i in range(3)], 0))
                D["sa_bin"] = torch.cat([pad_heads_b(bi[i * C:(i + 1) * C] * qs[i]) for i in range(3)], 0).contiguous()
                D["sa_wo"], D["sa_bo"] = WP(f32(p + "self_attn.out_proj.weight")), f32(p + "self_attn.out_proj.bias")
            wi, bi = f32(p + "multihead_attn.in_proj_weight"), f32(p + "multihead_attn.in_proj_bias")
            D["ca_wq"], D["ca_bq"] = WP(pad_heads(wi[:C] * qs_dec)), pad_heads_b(bi[:C] * qs_dec).contiguous()
            kv_w += [pad_heads(wi[C:2 * C]), pad_heads(wi[2 * C:])]
            kv_b += [pad_heads_b(bi[C:2 * C]), pad_heads_b(bi[2 * C:])]
            D["ca_wo"], D["ca_bo"] = WP(f32(p + "multihead_attn.out_proj.weight")), f32(p + "multihead_attn.out_proj.bias")
            D["w1"], D["b1"] = WP(f32(p + "linear1.weight")), f32(p + "linear1.bias")
            D["w2"], D["b2"] = WP(f32(p + "linear2.weight")), f32(p + "linear2.bias")
            for n in (1, 2, 3):
                D[f"ln{n}_g"], D[f"ln{n}_b"] = f32(p + f"norm{n}.weight"), f32(p + f"norm{n}.bias")
            self.dec.append(D)
        # K/V projection of BOTH layers as one weight: rows [l*2E, l*2E+E) = K_l, [l*2E+E, (l+1)*2E) = V_l
        self.kv_w, self.kv_b = W(torch.cat(kv_w, 0)), torch.cat(kv_b, 0).contiguous()
        self.E = E
        # --- head
        self.h0_w, self.h0_b = WP(f32("ref_cross.head.0.weight")), f32("ref_cross.head.0.bias")
        h2w, h2b = f32("ref_cross.head.2.weight"), f32("ref_cross.head.2.bias")
        if precision == "bf16":  # rows padded 196 -> 224 (one tcgen05 N tile)
            h2w = torch.nn.functional.pad(h2w, (0, 0, 0, 28))
            h2b = torch.nn.functional.pad(h2b, (0, 28))
        self.h2_w, self.h2_b = WP(h2w), h2b.contiguous()
        self._tables = {}

    # resampled tables are cached per patch grid
    def tables(self, ph: int, pw: int, stream):
        key = (ph, pw)
        if key not in self._tables:
            dev = self.device
            g = int(round(math.sqrt(self.pos.shape[0] - 1)))
            if ph == g and pw == g:  # modeling_dinov2.py:71-72: table used as is
                pos = self.pos
            else:
                pos = torch.empty(1 + ph * pw, C, device=dev, dtype=torch.float32)
                pos[0] = self.pos[0]
                step_h, step_w = pos_interp_steps(self.pos_interp, g, ph, pw)
                call("xs_pos_embed_resample_bicubic_steps", _ptr(self.pos[1:]), _ptr(pos[1:]), g, g, ph, pw, C,
                     step_h, step_w, stream)
            pe_h, pe_w = self.pe_table.shape[:2]
            if ph == pe_h and pw == pe_w:  # positional_encoding.py:51-56 shortcut
                pe = self.pe_table.reshape(ph * pw, C)
            else:
                pe = torch.empty(ph * pw, C, device=dev, dtype=torch.float32)
                call("xs_pe_resample_bilinear_ac", _ptr(self.pe_table), _ptr(pe), pe_h, pe_w, ph, pw, C, stream)
            self._tables[key] = (pos, pe)
        return self._tables[key]


class Engine:
    """Runs the forward on one GPU.  One Engine per (module, device, precision)."""

    def __init__(self, sd, device, precision="bf16", do_self_attn=True, do_short_cut=True,
                 use_tanh=False, power=1.0, pos_interp="scale_factor"):
        _lib.load()
        with torch.cuda.device(device):
            call("xs_device_check")
        self.w = PackedWeights(sd, device, precision, do_self_attn, fold_qscale=precision == "bf16",
                               pos_interp=pos_interp)
        self.device = device
        self.dt = self.w.dt
        self.adtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.qdtype = self.adtype   # q / k / v and the decoder K/V cache
        self.attn_dt = self.dt
        self.do_self_attn, self.do_short_cut = do_self_attn, do_short_cut
        self.use_tanh, self.power = bool(use_tanh), float(power)
        self._ws = {}
        # bf16 mode: residual adds of the DINOv2 blocks happen in the GEMM epilogue (XS_FUSE_RESIDUAL=0: separate
        # bf16 delta + add inside the LayerNorm kernel, the pre-fusion plan, kept for A/B measurements)
        self.fuse_residual = os.environ.get("XS_FUSE_RESIDUAL", "1") != "0"
        # XS_FUSE_LN=1 also runs the LayerNorm that follows each residual add in that epilogue (xs_gemm_bias_residual_ln).
        # Parity-green and 23 launches fewer, but the step time is the same in an in-run A/B (28.46-28.58 vs 28.43-28.47
        # ms: the second pass over the row block costs what the LayerNorm kernel did), so the proven plan stays default.
        self.fuse_ln = self.fuse_residual and os.environ.get("XS_FUSE_LN", "0") == "1"
        # XS_FOLD_LN=1 (opt-in, experimental): for batches that fill the GPU the LayerNorms of the DINOv2 blocks are folded into
        # the GEMMs around them -- the residual epilogue emits a bf16 copy of h and per-row (mean, rstd)
        # (xs_gemm_bias_residual_stats), the consuming q|k|v / fc1 GEMM normalises in its epilogue (xs_gemm_ln_folded).
        # Parity-tested, but measured SLOWER (27.94 vs 27.68 ms per step in one run: the 23 LayerNorm launches, 2.3 ms, are
        # gone, the four GEMMs' epilogues grow by 2.6 ms) and less precise with outlier channels (h instead of LayerNorm(h)
        # is what gets rounded to bf16: 5.9e-2 max-abs between the plans on the outlier weights, 5.8e-3 on benign ones).
        self.fold_ln = self.fuse_residual and os.environ.get("XS_FOLD_LN", "0") == "1"
        self.prof = None  # list of (tag, algorithmic flops, algorithmic bytes, start event, stop event) when profiling
        self._last_stream = None   # the engine's scratch buffers are shared by every call: see serialise()
        self._last_event = None

    @contextmanager
    def _op(self, tag, flops=0.0, nbytes=0.0, launches=1):
        """Counts kernel launches; when self.prof is a list also brackets the op with CUDA events on the
        launching stream (used by bench.py for the per-kernel roofline, outside the timed steps)."""
        _lib.count_launch(launches)
        if self.prof is None:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.prof.append((tag, float(flops), float(nbytes), s, e))

    # ---- scratch -----------------------------------------------------------------------------
    def _buf(self, name, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        key = (name, dtype)
        cur = self._ws.get(key)
        if cur is None or cur.numel() < n:
            cur = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self._ws[key] = cur
        return cur[:n].view(*shape)

    # ---- thin op wrappers ----------------------------------------------------------------------
    def _gemm(self, A, Wt, bias, out, act, st, tag="gemm", n_real=None, k_real=None):
        """out = act(A @ Wt^T + bias).  Operand mode follows the tensors: bf16 x bf16 -> tcgen05 kind::f16;
        fp32 x fp32 -> TF32 tensor cores in the bf16 product mode, SIMT FMA in the fp32 parity mode."""
        M, K = A.shape
        N = Wt.shape[0]
        assert A.dtype == Wt.dtype, (A.dtype, Wt.dtype)
        if A.dtype == torch.bfloat16:
            dt = DT_BF16
        else:
            dt = DT_TF32 if self.dt == DT_BF16 else DT_F32
        odt = {torch.bfloat16: DT_BF16, torch.float16: DT_F16, torch.float32: DT_F32}[out.dtype]
        flops = 2.0 * M * (n_real or N) * (k_real or K)  # algorithmic: padding is not counted
        with self._op(tag, flops):
            call("xs_gemm_bias_act", _ptr(A), A.stride(0), _ptr(Wt), Wt.stride(0), _ptr(bias), _ptr(out),
                 out.stride(0), M, N, K, act, dt, odt, st)

    def _gemm_residual(self, A, Wt, bias, h, st, tag="gemm"):
        """h (fp32, in place) += A @ Wt^T + bias: residual add fused into the GEMM epilogue (bf16 mode)."""
        M, K = A.shape
        N = Wt.shape[0]
        assert A.dtype == torch.bfloat16 and Wt.dtype == torch.bfloat16 and h.dtype == torch.float32
        with self._op(tag, 2.0 * M * N * K):
            call("xs_gemm_bias_residual", _ptr(A), A.stride(0), _ptr(Wt), Wt.stride(0), _ptr(bias), _ptr(h),
                 h.stride(0), M, N, K, DT_BF16, st)

    def _gemm_residual_ln(self, A, Wt, bias, h, g, b, eps, y, st, tag="gemm"):
        """h (fp32, in place) += A @ Wt^T + bias;  y (bf16) = LayerNorm(h): one kernel when the batch fills the GPU
        (xs_gemm_bias_residual_ln), the LayerNorm's launch and its HBM read of h are gone."""
        M, K = A.shape
        N = Wt.shape[0]
        assert A.dtype == torch.bfloat16 and h.dtype == torch.float32 and y.dtype == torch.bfloat16
        with self._op(tag + "_ln", 2.0 * M * N * K):
            call("xs_gemm_bias_residual_ln", _ptr(A), A.stride(0), _ptr(Wt), Wt.stride(0), _ptr(bias), _ptr(h),
                 h.stride(0), _ptr(g), _ptr(b), eps, _ptr(y), y.stride(0), M, N, K, DT_BF16, st)

    @staticmethod
    def fold_ln_rows(rows: int) -> bool:
        """Row counts for which the folded-LayerNorm plan is used: the fused statistics epilogue needs every CTA pair to
        own a 256-row block (xs_gemm_tc.cu: num_m2 >= SMs / 2); below that the plain plan is as fast."""
        return (rows + 255) // 256 >= NUM_SMS_HINT // 2

    def _gemm_residual_stats(self, A, Wt, bias, h, hb, stats, eps, st, tag="gemm"):
        """h (fp32, in place) += A @ Wt^T + bias;  hb = bf16(h);  stats = (mean, rstd) per row of h."""
        M, K = A.shape
        N = Wt.shape[0]
        assert A.dtype == torch.bfloat16 and h.dtype == torch.float32 and hb.dtype == torch.bfloat16
        with self._op(tag, 2.0 * M * N * K):
            call("xs_gemm_bias_residual_stats", _ptr(A), A.stride(0), _ptr(Wt), Wt.stride(0), _ptr(bias), _ptr(h),
                 h.stride(0), _ptr(hb), hb.stride(0), _ptr(stats), eps, M, N, K, DT_BF16, st)

    def _gemm_ln_folded(self, hb, Wf, c0, c1, stats, out, act, st, tag="gemm"):
        """out = act(LayerNorm(h) @ W^T + b) from the bf16 copy of h, the gamma-folded weight and the row statistics."""
        M, K = hb.shape
        N = Wf.shape[0]
        with self._op(tag, 2.0 * M * N * K):
            call("xs_gemm_ln_folded", _ptr(hb), hb.stride(0), _ptr(Wf), Wf.stride(0), _ptr(c0), _ptr(c1), _ptr(stats),
                 _ptr(out), out.stride(0), M, N, K, act, DT_BF16, st)

    def _ln(self, res_in, delta, res_out, g, b, eps, y, y32, rows, st):
        dt = DT_BF16 if (delta is not None and delta.dtype == torch.bfloat16) or \
            (y is not None and y.dtype == torch.bfloat16) else DT_F32
        nb = sum(t.element_size() * rows * C for t in (res_in, delta, res_out, y, y32) if t is not None)
        with self._op("layernorm", 0.0, nb):
            call("xs_layernorm", _ptr(res_in), _ptr(delta), _ptr(res_out), _ptr(g), _ptr(b), eps, _ptr(y), _ptr(y32),
                 rows, dt, st)

    def attn_scale(self, d):
        """Softmax scale the attention kernels are given: with the scale folded into the query projection the
        logits are log2-domain already and the kernel's scale * log2(e) must be 1."""
        return math.log(2.0) if self.w.fold_qscale else 1.0 / math.sqrt(d)

    def _attn(self, q, k, v, o, B, heads, Lq, Lk, d, slot, q_rs, q_bs, kv_rs, kv_bs, kv_shared, st,
              lse=None, name="att"):
        """q/k/v are tensor views whose data_ptr is the first column of the respective part; o is bf16 or
        fp32 (B*Lq, heads*d)."""
        scale = self.attn_scale(d)
        units, slots, t_blk = attn_units(B, heads, Lq)
        nsplit = attn_kv_splits(units, (Lk + 127) // 128, slots, t_blk)
        o_is_f32 = 1 if o.dtype == torch.float32 else 0
        flops = 4.0 * B * heads * Lq * Lk * d  # QK^T + PV
        if nsplit == 1:
            with self._op("attn_" + name, flops):
                call("xs_flash_attn", _ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(lse), B, heads, Lq, Lk, d, slot,
                     q_rs, q_bs, kv_rs, kv_bs, int(kv_shared), 1, o_is_f32, scale, self.attn_dt, st)
        else:  # small batch: split the keys across CTAs, then merge (same merge as the multi-GPU path)
            o_parts = self._buf(name + "_oparts", (nsplit, B * Lq, heads * d), torch.float32)
            l_parts = self._buf(name + "_lparts", (nsplit, B, heads, Lq), torch.float32)
            with self._op("attn_" + name, flops):
                call("xs_flash_attn", _ptr(q), _ptr(k), _ptr(v), _ptr(o_parts), _ptr(l_parts), B, heads, Lq, Lk, d,
                     slot, q_rs, q_bs, kv_rs, kv_bs, int(kv_shared), nsplit, 1, scale, self.attn_dt, st)
            with self._op("lse_merge", 0.0, o_parts.numel() * 4.0):
                call("xs_lse_merge", _ptr(o_parts), _ptr(l_parts), _ptr(o), _ptr(lse), nsplit, B, Lq, heads, d, 0, 0,
                     DT_F32 if o_is_f32 else DT_BF16, st)

    # ---- DINOv2 backbone over a list of image tensors (each (n,3,H,W) fp32, same H,W) --------------
    def backbone(self, image_groups, st):
        """Returns (h, delta, n_img, P): the fp32 residual stream BEFORE the last fc2 residual add and the
        pending delta, so the caller fuses `h + delta -> final LN` into its own consumer kernel."""
        w = self.w
        H, Wd = image_groups[0].shape[-2:]
        ph, pw = H // PATCH, Wd // PATCH
        P, T = ph * pw, ph * pw + 1
        I = sum(int(g.shape[0]) for g in image_groups)
        pos, _ = w.tables(ph, pw, st)
        A = self.adtype
        pe_dt = DT_TF32 if self.dt == DT_BF16 else DT_F32  # patch embedding: fp32 operands in both modes
        tok = self._buf("tok", (I * P, C), torch.float32)
        row = 0
        for g in image_groups:
            n = int(g.shape[0])
            assert g.dtype == torch.float32 and g.is_contiguous() and g.shape[1] == 3
            nbytes = _lib.load().xs_workspace_bytes(_lib.OP_PATCH_EMBED, n, H, Wd, pe_dt)
            ws = self._buf("im2col", (nbytes,), torch.uint8)
            with self._op("patch_embed", 2.0 * n * P * 588 * C, launches=2):
                call("xs_patch_embed", _ptr(g), _ptr(w.w_pe), _ptr(w.b_pe), _ptr(tok[row * P:]), _ptr(ws), nbytes, n, H,
                     Wd, pe_dt, st)
            row += n
        R = I * T
        h = self._buf("h", (R, C), torch.float32)
        y = self._buf("y", (R, C), A)
        qkv = self._buf("qkv", (R, 3 * C), self.qdtype)
        att = self._buf("att", (R, C), A)
        d = self._buf("d", (R, C), A)
        g1 = self._buf("g", (R, 4 * C), A)
        L0 = w.layers[0]
        with self._op("embed_ln", 0.0, R * C * (4 + 4 + y.element_size())):
            call("xs_embed_cls_pos_ln", _ptr(tok), DT_F32, _ptr(w.cls), _ptr(pos), _ptr(h), _ptr(L0["ln1_g"]),
                 _ptr(L0["ln1_b"]), DINO_EPS, _ptr(y), I, P, self.dt, st)
        fused = self.dt == DT_BF16 and self.fuse_residual
        fold = fused and self.fold_ln and not self.fuse_ln and self.fold_ln_rows(R)
        stats = self._buf("ln_stats", (R, 2), torch.float32) if fold else None
        for l, L in enumerate(w.layers):
            if fold and l > 0:  # y holds bf16(h) and stats its row statistics (fc2 epilogue of the previous layer)
                self._gemm_ln_folded(y, L["wqkv_f"], L["c0_qkv"], L["c1_qkv"], stats, qkv, ACT_NONE, st, tag="gemm_dino_qkv")
            else:
                self._gemm(y, L["wqkv"], L["bqkv"], qkv, ACT_NONE, st, tag="gemm_dino_qkv")
            self._attn(qkv[:, 0:], qkv[:, C:], qkv[:, 2 * C:], att, I, DINO_HEADS, T, T, 64, 64,
                       3 * C, T * 3 * C, 3 * C, T * 3 * C, False, st, name="dino")
            if fold:
                # h += att Wo^T + bo with y = bf16(h) and the row statistics out of the same epilogue; norm2 happens inside
                # fc1's epilogue; the same for fc2 -> next layer's norm1 -> q|k|v
                self._gemm_residual_stats(att, L["wo"], L["bo"], h, y, stats, DINO_EPS, st, tag="gemm_dino_proj")
                self._gemm_ln_folded(y, L["w1_f"], L["c0_fc1"], L["c1_fc1"], stats, g1, ACT_GELU, st, tag="gemm_dino_fc1")
                if l + 1 < DINO_LAYERS:
                    self._gemm_residual_stats(g1, L["w2"], L["b2"], h, y, stats, DINO_EPS, st, tag="gemm_dino_fc2")
                else:
                    self._gemm_residual(g1, L["w2"], L["b2"], h, st, tag="gemm_dino_fc2")
                continue
            if fused and self.fuse_ln:
                # residual add AND the following LayerNorm in the GEMM epilogue: h += att Wo^T + bo; y = LN2(h)
                self._gemm_residual_ln(att, L["wo"], L["bo"], h, L["ln2_g"], L["ln2_b"], DINO_EPS, y, st,
                                       tag="gemm_dino_proj")
                self._gemm(y, L["w1"], L["b1"], g1, ACT_GELU, st, tag="gemm_dino_fc1")
                if l + 1 < DINO_LAYERS:
                    Ln = w.layers[l + 1]  # h += g W2^T + b2; y = LN1_{l+1}(h)
                    self._gemm_residual_ln(g1, L["w2"], L["b2"], h, Ln["ln1_g"], Ln["ln1_b"], DINO_EPS, y, st,
                                           tag="gemm_dino_fc2")
                else:
                    self._gemm_residual(g1, L["w2"], L["b2"], h, st, tag="gemm_dino_fc2")
                continue
            if fused:
                # the GEMM epilogue adds its tile into the fp32 residual stream; the LayerNorm that follows only reads h
                self._gemm_residual(att, L["wo"], L["bo"], h, st, tag="gemm_dino_proj")
                self._ln(h, None, None, L["ln2_g"], L["ln2_b"], DINO_EPS, y, None, R, st)  # y = LN2(h)
                self._gemm(y, L["w1"], L["b1"], g1, ACT_GELU, st, tag="gemm_dino_fc1")
                self._gemm_residual(g1, L["w2"], L["b2"], h, st, tag="gemm_dino_fc2")
                if l + 1 < DINO_LAYERS:
                    Ln = w.layers[l + 1]
                    self._ln(h, None, None, Ln["ln1_g"], Ln["ln1_b"], DINO_EPS, y, None, R, st)  # y = LN1_{l+1}(h)
                continue
            self._gemm(att, L["wo"], L["bo"], d, ACT_NONE, st, tag="gemm_dino_proj")
            self._ln(h, d, h, L["ln2_g"], L["ln2_b"], DINO_EPS, y, None, R, st)  # h += d ; y = LN2(h)
            self._gemm(y, L["w1"], L["b1"], g1, ACT_GELU, st, tag="gemm_dino_fc1")
            self._gemm(g1, L["w2"], L["b2"], d, ACT_NONE, st, tag="gemm_dino_fc2")
            if l + 1 < DINO_LAYERS:
                Ln = w.layers[l + 1]
                self._ln(h, d, h, Ln["ln1_g"], Ln["ln1_b"], DINO_EPS, y, None, R, st)  # h += d ; y = LN1_{l+1}(h)
        return h, (None if fused else d), I, P

    def features(self, query_img, ref_imgs, st, want_mem=True, pe_table=None):
        """DINOv2 + final LN + CLS drop + PE.  query_img (B,3,H,W) or None, ref_imgs (B,N,3,H,W) or None.
        Returns xq32 (B*P,C) fp32 and mem (B*N*P,C) AT (None where not requested).
        pe_table (P, C) fp32 replaces the resampled multi-view PE (get_featmaps passes zeros)."""
        w = self.w
        groups, nq = [], 0
        if query_img is not None:
            groups.append(query_img)
            nq = int(query_img.shape[0])
        nr = 0
        if ref_imgs is not None:
            r = ref_imgs.reshape(-1, *ref_imgs.shape[-3:])
            groups.append(r)
            nr = int(r.shape[0])
        H, Wd = groups[0].shape[-2:]
        ph, pw = H // PATCH, Wd // PATCH
        P = ph * pw
        _, pe = w.tables(ph, pw, st)
        if pe_table is not None:
            assert pe_table.shape == (P, C) and pe_table.dtype == torch.float32 and pe_table.is_contiguous()
            pe = pe_table
        xq32 = self._buf("xq32", (nq * P, C), torch.float32) if nq else None
        mem = self._buf("mem", (nr * P, C), self.adtype) if nr and want_mem else None
        es = self.adtype.itemsize
        flat = [g.reshape(-1, *g.shape[-3:]) for g in groups]
        h, d, I, _ = self.backbone(flat, st)
        with self._op("final_ln_pe", 0.0, I * (P + 1) * C * (4 + (es if d is not None else 0) + es)):
            call("xs_final_ln_drop_cls_add_pe", _ptr(h), _ptr(d), _ptr(w.lnf_g), _ptr(w.lnf_b), DINO_EPS,
                 _ptr(pe), _ptr(xq32), None, _ptr(mem), I, nq, P, self.dt, st)
        return xq32, mem

    def project_kv(self, mem, st, out=None):
        """K/V of both decoder layers for reference tokens mem (rows, C) -> (rows, 4E)."""
        rows = mem.shape[0]
        kv = out if out is not None else self._buf("kv", (rows, 4 * self.w.E), self.qdtype)
        self._gemm(mem, self.w.kv_w, self.w.kv_b, kv, ACT_NONE, st, tag="gemm_dec_kv", n_real=4 * C)
        return kv

    def decode(self, xq32, kv, B, P, M, ph, pw, st, kv_shared=False, need_attn_weights=False, head_id=0,
               cross_attn_fn=None):
        """2-layer post-norm decoder + head + jigsaw.  xq32 (B*P, C) fp32 decoder stream (updated in place);
        kv: (B*M or M, 4E) projected reference K/V.  The decoder stream and every GEMM input/output except
        Q/K/V stay fp32 (TF32 tensor cores in the bf16 product mode): these 2 % of the FLOPs carry half of the
        bf16 error budget.  cross_attn_fn (optional) replaces the local cross-attention (multi-GPU split-KV)."""
        w, A, E, slot = self.w, self.adtype, self.w.E, self.w.slot
        R = B * P
        f32 = torch.float32
        qkv_s = self._buf("dec_qkv", (R, 3 * E), self.qdtype)
        qc = self._buf("dec_q", (R, E), self.qdtype)
        att = self._buf("dec_att", (R, C), f32)
        d = self._buf("dec_d", (R, C), f32)
        f = self._buf("dec_f", (R, C), f32)
        lse = self._buf("dec_lse", (B, DEC_HEADS, P), f32) if need_attn_weights else None
        sc = self.do_short_cut
        for l, D in enumerate(w.dec):
            if self.do_self_attn:
                self._gemm(xq32, D["sa_win"], D["sa_bin"], qkv_s, ACT_NONE, st, tag="gemm_dec", n_real=3 * C)
                self._attn(qkv_s[:, 0:], qkv_s[:, E:], qkv_s[:, 2 * E:], att, B, DEC_HEADS, P, P, DEC_D, slot,
                           3 * E, P * 3 * E, 3 * E, P * 3 * E, False, st, name="dsa")
                self._gemm(att, D["sa_wo"], D["sa_bo"], d, ACT_NONE, st, tag="gemm_dec")
                self._ln(xq32 if sc else None, d, None, D["ln1_g"], D["ln1_b"], DEC_EPS, None, xq32, R, st)
            self._gemm(xq32, D["ca_wq"], D["ca_bq"], qc, ACT_NONE, st, tag="gemm_dec", n_real=C)
            last = need_attn_weights and l == DEC_LAYERS - 1
            if cross_attn_fn is not None:
                cross_attn_fn(l, qc, att, lse if last else None, st)
            else:
                k_view, v_view = kv[:, l * 2 * E:], kv[:, l * 2 * E + E:]
                self._attn(qc, k_view, v_view, att, B, DEC_HEADS, P, M, DEC_D, slot, E, P * E, 4 * E, M * 4 * E,
                           kv_shared, st, lse=lse if last else None, name="dca")
            self._gemm(att, D["ca_wo"], D["ca_bo"], d, ACT_NONE, st, tag="gemm_dec")
            self._ln(xq32 if sc else None, d, None, D["ln2_g"], D["ln2_b"], DEC_EPS, None, xq32, R, st)
            self._gemm(xq32, D["w1"], D["b1"], f, ACT_RELU, st, tag="gemm_dec")
            self._gemm(f, D["w2"], D["b2"], d, ACT_NONE, st, tag="gemm_dec")
            self._ln(xq32, d, None, D["ln3_g"], D["ln3_b"], DEC_EPS, None, xq32, R, st)
        probs = None
        if need_attn_weights:
            probs = torch.empty(B, P, M, device=self.device, dtype=f32)
            l = DEC_LAYERS - 1
            with self._op("attn_probs", 0.0, probs.numel() * 4.0):
                call("xs_attn_probs_one_head", _ptr(qc), _ptr(kv[:, l * 2 * E:]), _ptr(lse), _ptr(probs), B, DEC_HEADS,
                     head_id, P, M, DEC_D, slot, E, P * E, 4 * E, 0 if kv_shared else M * 4 * E,
                     self.attn_scale(DEC_D), self.attn_dt, st)
        self._gemm(xq32, w.h0_w, w.h0_b, f, ACT_LEAKY, st, tag="gemm_dec")
        score = torch.empty(B, PATCH * ph, PATCH * pw, device=self.device, dtype=f32)
        # K12 algorithmic bytes: token features in + fp32 score map out
        with self._op("head_jigsaw", 2.0 * R * C * 196, R * C * f.element_size() + score.numel() * 4.0):
            call("xs_head_score_jigsaw", _ptr(f), f.stride(0), _ptr(w.h2_w), w.h2_w.stride(0), _ptr(w.h2_b),
                 _ptr(score), B, ph, pw, C, int(self.use_tanh), self.power,
                 DT_TF32 if self.dt == DT_BF16 else DT_F32, st)
        return score, probs

    # ---- split-KV pieces used by crossscore_b200.scene.SplitKVScorer -------------------------------------
    @property
    def kv_width(self):
        return 4 * self.w.E

    @property
    def kv_dtype(self):
        return self.qdtype

    def cross_attn_partial(self, layer, qc, kv_local, B, P, M_local, packed, st):
        """Cross-attention of decoder layer `layer` over THIS rank's keys only.  packed (fp32, 1-D) receives the
        normalised partial O (B*P*C) followed by LSE (B*8*P) -- the unit one all-gather moves per layer."""
        E, slot = self.w.E, self.w.slot
        o = packed[:B * P * C].view(B * P, C)
        lse = packed[B * P * C:B * P * C + B * DEC_HEADS * P].view(B, DEC_HEADS, P)
        self._attn(qc, kv_local[:, layer * 2 * E:], kv_local[:, layer * 2 * E + E:], o, B, DEC_HEADS, P, M_local,
                   DEC_D, slot, E, P * E, 4 * E, M_local * 4 * E, False, st, lse=lse, name="dca_part")

    def merge_partials(self, gathered, n_parts, B, P, att, lse_out, st):
        """gathered: n_parts packed (O_r | LSE_r) buffers back to back -> att (B*P, C) fp32 (+ merged LSE)."""
        part = B * P * C + B * DEC_HEADS * P
        with self._op("lse_merge", 0.0, n_parts * part * 4.0 + att.numel() * 4.0):
            call("xs_lse_merge", _ptr(gathered), gathered.data_ptr() + B * P * C * 4, _ptr(att), _ptr(lse_out),
                 n_parts, B, P, DEC_HEADS, DEC_D, part, part, DT_F32, st)

    def merge_partials_peers(self, ptrs_dev: int, base: int, n_parts, B, P, att, lse_out, st):
        """Same merge, pulling each rank's packed (O_r | LSE_r) buffer through its NVLink peer pointer
        (ptrs_dev: device array of n_parts pointers; base: element offset of this layer's buffer)."""
        part = B * P * C + B * DEC_HEADS * P
        with self._op("lse_merge_peers", 0.0, n_parts * part * 4.0 + att.numel() * 4.0):
            call("xs_lse_merge_peers", ptrs_dev, base, B * P * C, _ptr(att), _ptr(lse_out), n_parts, B, P, DEC_HEADS,
                 DEC_D, DT_F32, st)

    @contextmanager
    def serialise(self):
        """The scratch buffers (h, y, qkv, ...) belong to the engine, not to a call, so two calls on DIFFERENT streams
        must not overlap on the device.  A call made on another stream than the previous one first waits for that one's
        completion event; calls on one stream (the normal case, and every CUDA-graph capture) cost nothing extra."""
        cur = torch.cuda.current_stream(self.device)
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing and self._last_event is not None and self._last_stream != cur.cuda_stream:
            cur.wait_event(self._last_event)
        try:
            yield
        finally:
            if not capturing:
                if self._last_event is None or self._last_stream != cur.cuda_stream:
                    self._last_event = torch.cuda.Event()
                self._last_event.record(cur)
                self._last_stream = cur.cuda_stream

    # ---- the reference-shaped forward ------------------------------------------------------------
    def forward(self, query_img, ref_imgs, need_attn_weights=False, head_id=0):
        if need_attn_weights and not 0 <= head_id < DEC_HEADS:
            raise IndexError(f"need_attn_weights_head_id={head_id} out of range for {DEC_HEADS} heads")
        with self.serialise():
            return self._forward(query_img, ref_imgs, need_attn_weights, head_id)

    def _forward(self, query_img, ref_imgs, need_attn_weights, head_id):
        st = torch.cuda.current_stream(self.device).cuda_stream
        B, _, H, Wd = query_img.shape
        N = ref_imgs.shape[1]
        ph, pw = H // PATCH, Wd // PATCH
        P = ph * pw
        xq32, mem = self.features(query_img, ref_imgs, st)
        kv = self.project_kv(mem, st)
        score, probs = self.decode(xq32, kv, B, P, N * P, ph, pw, st, kv_shared=False,
                                   need_attn_weights=need_attn_weights, head_id=head_id)
        if probs is not None:
            probs = probs.view(B, ph, pw, N, ph, pw)
        return score, probs
