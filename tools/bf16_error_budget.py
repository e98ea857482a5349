"""CPU emulation of the bf16 path's rounding points, to see which ones dominate the score-map error.
(Development tool: uses the oracle's math with toggled bf16 rounding; not part of the product.)"""
import math, sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from oracle import crossscore_oracle as O
from crossscore_b200.synthetic import make_inputs, make_state_dict

ALL = ["w", "img", "tok", "y", "qkv", "p", "att", "d", "g", "x", "dqkv", "dp", "datt", "dd", "df", "hf", "mem", "kv"]


def ln_fold(h, gamma, beta, W, b, eps=1e-6):
    """LayerNorm folded into the GEMM that follows: bf16(h) x bf16(gamma * W)^T in fp32, then per row
    rstd * (acc - mean * c1) + (W beta + b); mean / rstd from the fp32 h."""
    dt = h.dtype
    mean = h.mean(-1, keepdim=True)
    rstd = torch.rsqrt(h.var(-1, unbiased=False, keepdim=True) + eps)
    Wf = (W * gamma[None, :]).bfloat16().to(dt)
    c1 = Wf.sum(-1)
    c0 = W @ beta + b
    acc = h.bfloat16().to(dt) @ Wf.T
    return rstd * (acc - mean * c1) + c0


def run(sd, q, r, active, dt=torch.float32):
    R = lambda x, tag: x.bfloat16().to(dt) if tag in active else x
    g = lambda k: sd[k].to(dt)
    gw = lambda k: R(sd[k].to(dt), "w")
    B, _, H, W = q.shape
    N = r.shape[1]
    ph, pw = H // 14, W // 14
    P = ph * pw
    imgs = torch.cat([q, r.reshape(-1, 3, H, W)], 0).to(dt)
    I = imgs.shape[0]
    x = imgs[:, :, :ph * 14, :pw * 14]
    patches = x.reshape(I, 3, ph, 14, pw, 14).permute(0, 2, 4, 1, 3, 5).reshape(I, P, 588)
    tok = R(R(patches, "img") @ gw("backbone.embeddings.patch_embeddings.projection.weight").reshape(384, -1).T
            + g("backbone.embeddings.patch_embeddings.projection.bias"), "tok")
    h = torch.cat([g("backbone.embeddings.cls_token").expand(I, -1, -1), tok], 1) + O.dinov2_pos_table(sd, H, W, dt)[None]
    T = P + 1
    for l in range(12):
        p = f"backbone.encoder.layer.{l}."
        lam1, lam2 = g(p + "layer_scale1.lambda1"), g(p + "layer_scale2.lambda1")
        wqkv0 = torch.cat([g(p + f"attention.attention.{n}.weight") for n in ("query", "key", "value")], 0)
        bqkv = torch.cat([g(p + f"attention.attention.{n}.bias") for n in ("query", "key", "value")], 0)
        if "fold" in active:
            qkv = R(ln_fold(h, g(p + "norm1.weight"), g(p + "norm1.bias"), wqkv0, bqkv), "qkv")
        else:
            y = R(O.layer_norm(h, g(p + "norm1.weight"), g(p + "norm1.bias"), 1e-6), "y")
            qkv = R(y @ R(wqkv0, "w").T + bqkv, "qkv")
        qq, kk, vv = [t.view(I, T, 6, 64).transpose(1, 2) for t in qkv.split(384, -1)]
        s = qq @ kk.transpose(-1, -2)
        if "s16" in active:  # fp16 logit accumulators (tcgen05 D format f16)
            s = s.half().to(dt)
        s = s / 8
        s = s - s.max(-1, keepdim=True).values
        if "x16" in active:  # fp16 softmax argument with a stale reference max (8 octaves low) and fp16 P
            x = ((s * 1.4426950408889634 + 8.0).half()).to(dt)
            e = torch.exp2(x).half().to(dt)
            a = (e @ vv) / e.sum(-1, keepdim=True)
        else:
            e = torch.exp(s)
            a = (R(e, "p") @ vv) / e.sum(-1, keepdim=True)
        a = R(a.transpose(1, 2).reshape(I, T, 384), "att")
        d = R(a @ R(g(p + "attention.output.dense.weight") * lam1[:, None], "w").T + g(p + "attention.output.dense.bias") * lam1, "d")
        h = h + d
        if "fold" in active:
            m = R(O.gelu_erf(ln_fold(h, g(p + "norm2.weight"), g(p + "norm2.bias"), g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias"))), "g")
        else:
            y = R(O.layer_norm(h, g(p + "norm2.weight"), g(p + "norm2.bias"), 1e-6), "y")
            m = R(O.gelu_erf(y @ gw(p + "mlp.fc1.weight").T + g(p + "mlp.fc1.bias")), "g")
        d = R(m @ R(g(p + "mlp.fc2.weight") * lam2[:, None], "w").T + g(p + "mlp.fc2.bias") * lam2, "d")
        h = h + d
    f = O.layer_norm(h, g("backbone.layernorm.weight"), g("backbone.layernorm.bias"), 1e-6)[:, 1:]
    pe = O.multiview_pe_table(sd, H, W, dt)
    f = f + pe[None]
    x32 = f[:B]
    mem = R(f[B:].reshape(B, N * P, 384), "mem")
    for l in range(2):
        p = f"ref_cross.attn.layers.{l}."
        def mha(xq, kvsrc, pre, tagkv):
            w_in, b_in = g(pre + "in_proj_weight"), g(pre + "in_proj_bias")
            qh = R(R(xq, "x") @ R(w_in[:384], "w").T + b_in[:384], "dqkv")
            kh = R(kvsrc @ R(w_in[384:768], "w").T + b_in[384:768], tagkv)
            vh = R(kvsrc @ R(w_in[768:], "w").T + b_in[768:], tagkv)
            Lq, Lk = qh.shape[1], kh.shape[1]
            qh, kh, vh = [t.view(B, -1, 8, 48).transpose(1, 2) for t in (qh, kh, vh)]
            s = qh @ kh.transpose(-1, -2)
            if "s16" in active:
                s = s.half().to(dt)
            s = s / math.sqrt(48)
            s = s - s.max(-1, keepdim=True).values
            if "x16" in active:
                x = ((s * 1.4426950408889634 + 8.0).half()).to(dt)
                e = torch.exp2(x).half().to(dt)
                a = (e @ vh) / e.sum(-1, keepdim=True)
            else:
                e = torch.exp(s)
                a = (R(e, "dp") @ vh) / e.sum(-1, keepdim=True)
            a = R(a.transpose(1, 2).reshape(B, Lq, 384), "datt")
            return R(a @ gw(pre + "out_proj.weight").T + g(pre + "out_proj.bias"), "dd")
        x32 = O.layer_norm(x32 + mha(x32, R(x32, "x"), p + "self_attn.", "dqkv"), g(p + "norm1.weight"), g(p + "norm1.bias"), 1e-5)
        x32 = O.layer_norm(x32 + mha(x32, mem, p + "multihead_attn.", "kv"), g(p + "norm2.weight"), g(p + "norm2.bias"), 1e-5)
        ff = R(torch.relu(R(x32, "x") @ gw(p + "linear1.weight").T + g(p + "linear1.bias")), "df")
        ff = R(ff @ gw(p + "linear2.weight").T + g(p + "linear2.bias"), "dd")
        x32 = O.layer_norm(x32 + ff, g(p + "norm3.weight"), g(p + "norm3.bias"), 1e-5)
    z = R(R(x32, "x") @ gw("ref_cross.head.0.weight").T + g("ref_cross.head.0.bias"), "hf")
    z = torch.where(z >= 0, z, 0.01 * z)
    z = z @ gw("ref_cross.head.2.weight").T + g("ref_cross.head.2.bias")
    return torch.sigmoid(z)


if __name__ == "__main__":
    torch.manual_seed(0)
    sd = make_state_dict(1)
    q, r = make_inputs(2, 3, 84, 112, seed=0)
    base = run(sd, q, r, set(), torch.float64)
    def err(active):
        d = (run(sd, q, r, set(active)).double() - base).abs()
        return d.max().item(), d.mean().item()
    if len(sys.argv) > 1 and sys.argv[1] == "fold":
        for variant in ("benign", "outlier"):
            sd = make_state_dict(1, variant=variant)
            base = run(sd, q, r, set(), torch.float64)
            print(variant, "all        ", err(ALL))
            print(variant, "all + fold ", err(ALL + ["fold"]))
            print(variant, "only y     ", err(["y", "w"]))
            print(variant, "only fold  ", err(["fold", "w"]))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "f16":
        print("all             ", err(ALL))
        print("all + s16       ", err(ALL + ["s16"]))
        print("all + s16 + x16 ", err(ALL + ["s16", "x16"]))
        print("only s16        ", err(["s16"]))
        print("only s16 + x16  ", err(["s16", "x16"]))
        print("only p          ", err(["p", "dp"]))
        sys.exit(0)
    print("none      ", err([]))
    print("all       ", err(ALL))
    for t in ALL:
        print(f"only {t:5s}", err([t]))
    for t in ALL:
        print(f"all-but {t:5s}", err([a for a in ALL if a != t]))
    dino = ["img", "tok", "y", "qkv", "p", "att", "d", "g"]
    print("dino only ", err(dino + ["w"]))
    print("dec only  ", err([a for a in ALL if a not in dino]))
