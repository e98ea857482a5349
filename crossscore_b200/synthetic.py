"""Deterministic synthetic weights and inputs (no network, no checkpoint on the box).

The real ``CrossScore-v1.0.0.ckpt`` and ``facebook/dinov2-small`` weights are absent
offline, so parity and the benchmark run on seeded weights that follow the
reference's exact 265-tensor ``state_dict`` schema.  The generator
is independent of the reference so the same tensors can be rebuilt on the GPU box; the
spreads are chosen so attention logits, GELU inputs and sigmoid inputs are not degenerate.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Tuple

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # utils/io/images.py:8-11
IMAGENET_STD = (0.229, 0.224, 0.225)

C = 384


def state_dict_spec(pe_h: int = 40, pe_w: int = 40, do_self_attn: bool = True):
    """(name, shape) for every tensor of CrossScoreNet.state_dict(), in the reference order."""
    spec = [("img_mean_std", (6,))]
    b = "backbone."
    spec += [
        (b + "embeddings.cls_token", (1, 1, C)),
        (b + "embeddings.mask_token", (1, C)),
        (b + "embeddings.position_embeddings", (1, 1370, C)),
        (b + "embeddings.patch_embeddings.projection.weight", (C, 3, 14, 14)),
        (b + "embeddings.patch_embeddings.projection.bias", (C,)),
    ]
    for l in range(12):
        p = f"{b}encoder.layer.{l}."
        spec += [
            (p + "norm1.weight", (C,)), (p + "norm1.bias", (C,)),
            (p + "attention.attention.query.weight", (C, C)), (p + "attention.attention.query.bias", (C,)),
            (p + "attention.attention.key.weight", (C, C)), (p + "attention.attention.key.bias", (C,)),
            (p + "attention.attention.value.weight", (C, C)), (p + "attention.attention.value.bias", (C,)),
            (p + "attention.output.dense.weight", (C, C)), (p + "attention.output.dense.bias", (C,)),
            (p + "layer_scale1.lambda1", (C,)),
            (p + "norm2.weight", (C,)), (p + "norm2.bias", (C,)),
            (p + "mlp.fc1.weight", (4 * C, C)), (p + "mlp.fc1.bias", (4 * C,)),
            (p + "mlp.fc2.weight", (C, 4 * C)), (p + "mlp.fc2.bias", (C,)),
            (p + "layer_scale2.lambda1", (C,)),
        ]
    spec += [(b + "layernorm.weight", (C,)), (b + "layernorm.bias", (C,))]
    spec += [("pos_enc_fn.PE", (1, pe_h, pe_w, C))]
    for l in range(2):
        p = f"ref_cross.attn.layers.{l}."
        if do_self_attn:
            spec += [
                (p + "self_attn.in_proj_weight", (3 * C, C)), (p + "self_attn.in_proj_bias", (3 * C,)),
                (p + "self_attn.out_proj.weight", (C, C)), (p + "self_attn.out_proj.bias", (C,)),
            ]
        spec += [
            (p + "multihead_attn.in_proj_weight", (3 * C, C)), (p + "multihead_attn.in_proj_bias", (3 * C,)),
            (p + "multihead_attn.out_proj.weight", (C, C)), (p + "multihead_attn.out_proj.bias", (C,)),
            (p + "linear1.weight", (C, C)), (p + "linear1.bias", (C,)),
            (p + "linear2.weight", (C, C)), (p + "linear2.bias", (C,)),
            (p + "norm1.weight", (C,)), (p + "norm1.bias", (C,)),
            (p + "norm2.weight", (C,)), (p + "norm2.bias", (C,)),
            (p + "norm3.weight", (C,)), (p + "norm3.bias", (C,)),
        ]
    spec += [
        ("ref_cross.head.0.weight", (C, C)), ("ref_cross.head.0.bias", (C,)),
        ("ref_cross.head.2.weight", (196, C)), ("ref_cross.head.2.bias", (196,)),
    ]
    return spec


def _scale_offset(name: str) -> Tuple[float, float]:
    if "norm" in name and name.endswith(".weight"):
        return 0.2, 1.0
    if "norm" in name and name.endswith(".bias"):
        return 0.1, 0.0
    if name.endswith("cls_token"):
        return 0.5, 0.0
    if name.endswith("mask_token"):
        return 0.0, 0.0
    if name.endswith("position_embeddings"):
        return 0.3, 0.0
    if "patch_embeddings.projection.weight" in name:
        return 0.05, 0.0
    if "attention.attention.query.weight" in name or "attention.attention.key.weight" in name:
        return 0.08, 0.0
    if "attention.attention.value.weight" in name or "attention.output.dense.weight" in name:
        return 0.04, 0.0
    if "mlp.fc1.weight" in name:
        return 0.05, 0.0
    if "mlp.fc2.weight" in name:
        return 0.03, 0.0
    if name.endswith("PE"):
        return 1.0, 0.0
    if "in_proj_weight" in name:
        return 0.04, 0.0
    if "out_proj.weight" in name:
        return 0.03, 0.0
    if "linear1.weight" in name or "linear2.weight" in name:
        return 0.05, 0.0
    if name == "ref_cross.head.0.weight":
        return 0.05, 0.0
    if name == "ref_cross.head.2.weight":
        return 0.1, 0.0
    if name.endswith(".bias") or name.endswith("in_proj_bias"):
        return 0.05, 0.0
    raise KeyError(name)


OUTLIER_CHANNELS = (5, 133, 301)


def make_state_dict(seed: int = 1, pe_h: int = 40, pe_w: int = 40, do_self_attn: bool = True,
                    lightning_prefix: bool = False, variant: str = "benign") -> "OrderedDict[str, torch.Tensor]":
    """Seeded fp32 state_dict with the reference schema.  ``lightning_prefix`` adds the
    ``model.`` prefix a Lightning checkpoint carries (task/core.py:173).

    variant "outlier" mimics what trained DINOv2 weights do to a bf16 pipeline: a few
    residual channels carry activations ~100x the rest (the rows of the patch projection, of every fc2 and of the
    position table that write channels OUTLIER_CHANNELS are scaled up), LayerScale spreads up to 1.5, and the
    query / key projections are 1.6x wider so attention logits reach +-35."""
    if variant not in ("benign", "outlier"):
        raise ValueError(f"unknown weight variant {variant!r}")
    sd = OrderedDict()
    for idx, (name, shape) in enumerate(state_dict_spec(pe_h, pe_w, do_self_attn)):
        if name == "img_mean_std":
            t = torch.tensor([*IMAGENET_MEAN, *IMAGENET_STD], dtype=torch.float32)
        elif name.endswith("lambda1"):
            g = torch.Generator().manual_seed(seed * 1000003 + idx)
            t = 0.05 + (1.45 if variant == "outlier" else 0.95) * torch.rand(shape, generator=g, dtype=torch.float32)
        else:
            g = torch.Generator().manual_seed(seed * 1000003 + idx)
            s, o = _scale_offset(name)
            t = torch.randn(shape, generator=g, dtype=torch.float32) * s + o
            if variant == "outlier":
                ch = list(OUTLIER_CHANNELS)
                if "patch_embeddings.projection.weight" in name:
                    t[ch] *= 40.0
                elif name.endswith("position_embeddings"):
                    t[..., ch] *= 40.0
                elif "mlp.fc2.weight" in name:
                    t[ch] *= 25.0
                elif "attention.attention.query.weight" in name or "attention.attention.key.weight" in name:
                    t *= 1.6
        sd[("model." if lightning_prefix else "") + name] = t
    return sd


def make_inputs(B: int, N_ref: int, H: int, W: int, seed: int = 0, shared_refs: bool = False):
    """ImageNet-normalised uniform-random images, the dataloader's contract
    (task/predict.py:69-74): query (B,3,H,W), refs (B,N,3,H,W) fp32."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    q = (torch.rand(B, 3, H, W, generator=g) - mean) / std
    if shared_refs:
        r = (torch.rand(1, N_ref, 3, H, W, generator=g) - mean[None]) / std[None]
        r = r.expand(B, -1, -1, -1, -1)
    else:
        r = (torch.rand(B, N_ref, 3, H, W, generator=g) - mean[None]) / std[None]
    return q.contiguous(), r.contiguous()
