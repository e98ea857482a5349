"""Drop-in ``CrossScoreNet`` for the reference's ``task/core.py:26-161``.

Same constructor / forward signature, same attribute tree and the same 265-tensor ``state_dict`` schema
(``load_state_dict(strict=True)`` accepts the reference checkpoint's ``state_dict``; a Lightning
``.ckpt`` with the ``model.`` prefix is handled by :func:`load_checkpoint`), so it can replace
``self.model = CrossScoreNet(cfg=self.cfg)`` in ``CrossScoreLightningModule`` (task/core.py:173).
The arithmetic runs in the hand-written sm_100a kernels of ``libcrossscore_sm100a.so``; the parameters
held here are only the source of the packed device copies.  There is no PyTorch / CPU fallback: calling
``forward`` without the built library or on a non-B200 device raises.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch
from torch import nn

from .config import default_cfg, resolve_score_activation
from .synthetic import make_state_dict, state_dict_spec


class _Tree(nn.Module):
    """Plain container; children/parameters are attached by dotted name so the state_dict keys equal the
    reference's (e.g. ``backbone.encoder.layer.3.attention.attention.query.weight``)."""

    def attach(self, dotted: str, tensor: torch.Tensor, as_buffer: bool = False):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Tree())
            node = node._modules[p]
        if as_buffer:
            node.register_buffer(parts[-1], tensor)
        else:
            node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class CrossScoreNet(nn.Module):
    """B200-native CrossScore model (inference).  See module docstring.

    Extra keywords (not in the reference): ``precision`` = "bf16" (tcgen05 tensor-core path, default) or
    "fp32" (parity mode, max-abs <= 1e-4 vs the fp32 reference).  ``dinov2_pos_interp`` selects how the DINOv2
    position table is resampled for inputs other than 518x518: "scale_factor" (default) reproduces transformers
    4.33.3, the version the reference pins (environment.yaml:340: ``F.interpolate(scale_factor=(h+0.1)/37)``);
    "size" reproduces transformers >= 4.4x (``F.interpolate(size=(h, w))``).  The two differ by up to ~0.2 in table
    entries on non-square grids, so pick the one the checkpoint's own stack used.
    """

    def __init__(self, cfg=None, precision: str = "bf16", dinov2_pos_interp: str = "scale_factor"):
        super().__init__()
        if dinov2_pos_interp not in ("scale_factor", "size"):
            raise ValueError(f"dinov2_pos_interp must be 'scale_factor' or 'size', got {dinov2_pos_interp!r}")
        self.dinov2_pos_interp = dinov2_pos_interp
        self.cfg = cfg if cfg is not None else default_cfg()
        m = self.cfg.model
        if not m.do_reference_cross:
            # utils/check_config.py:31-36: only the cross-reference predictor exists
            raise ValueError("Reference type must be 'cross'")
        if int(m.patch_size) != 14:
            raise ValueError("patch_size must be 14 (DINOv2 ViT-S/14 backbone)")
        if str(m.pos_enc.multi_view.interpolate_mode) != "bilinear":
            raise ValueError("only interpolate_mode='bilinear' is implemented (config/model/model.yaml:14)")
        if "dinov2-small" not in str(m.backbone.from_pretrained):
            raise ValueError("only the facebook/dinov2-small backbone is implemented")
        metric = m.predict.metric
        self._use_tanh, self._power = resolve_score_activation(metric.type, metric.min, metric.max, metric.power_factor)
        self.precision = precision
        # mirrors Dinov2Config fields the reference reads (task/core.py:52, cross_reference.py:30-33)
        self.dinov2_cfg = SimpleNamespace(hidden_size=384, num_hidden_layers=12, num_attention_heads=6,
                                          patch_size=14, image_size=518, layer_norm_eps=1e-6)
        pe_h, pe_w = int(m.pos_enc.multi_view.h), int(m.pos_enc.multi_view.w)
        init = make_state_dict(seed=0, pe_h=pe_h, pe_w=pe_w, do_self_attn=bool(m.decoder_do_self_attn))
        self.add_module("backbone", _Tree())
        self.add_module("pos_enc_fn", _Tree())
        self.add_module("ref_cross", _Tree())
        for name, _shape in state_dict_spec(pe_h, pe_w, bool(m.decoder_do_self_attn)):
            if name == "img_mean_std":
                self.register_buffer("img_mean_std", init[name])
                continue
            top, rest = name.split(".", 1)
            self._modules[top].attach(rest, init[name])
        self._engines = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.refresh())

    # ---- weights ---------------------------------------------------------------------------------
    def refresh(self):
        """Drop the packed device copies of the weights; the next forward repacks them from the parameters."""
        self._engines.clear()

    def _apply(self, fn, *a, **k):  # .to() / .cuda() move the source parameters; repack lazily
        out = super()._apply(fn, *a, **k)
        self.refresh()
        return out

    def _weights_stamp(self):
        """Cheap fingerprint of the source parameters: in-place ops on a parameter (``p.copy_()``, ``p.add_()``) bump
        its ``_version`` and ``p.data = ...`` changes ``data_ptr``, so the packed copies are rebuilt instead of going
        stale silently.  In-place ops THROUGH ``p.data`` (``p.data.copy_()``) touch neither: call ``refresh()``."""
        stamp = 0
        for t in list(self.parameters()) + list(self.buffers()):
            stamp = (stamp * 1000003 + t._version * 7919 + t.data_ptr()) & 0xFFFFFFFFFFFF
        return stamp

    def _engine(self, device):
        from .engine import Engine
        key = (str(device), self.precision)
        stamp = self._weights_stamp()
        cur = self._engines.get(key)
        if cur is None or cur[0] != stamp:
            cur = (stamp, Engine(self.state_dict(), device, self.precision,
                                 do_self_attn=bool(self.cfg.model.decoder_do_self_attn),
                                 do_short_cut=bool(self.cfg.model.decoder_do_short_cut),
                                 use_tanh=self._use_tanh, power=self._power, pos_interp=self.dinov2_pos_interp))
            self._engines[key] = cur
        return cur[1]

    # ---- reference API ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, query_img, ref_cross_imgs, need_attn_weights=False, need_attn_weights_head_id=0,
                norm_img=False):
        """query_img (B,3,H,W), ref_cross_imgs (B,N,3,H,W) fp32 CUDA tensors ->
        {"score_map_ref_cross": (B,14*(H//14),14*(W//14)) fp32,
         "attn_weights_map_ref_cross": None | (B,ph,pw,N,ph,pw) fp32}   (task/core.py:58-117)."""
        if norm_img:
            # the reference's norm_img=True branch divides by the MEAN (std slice bug, task/core.py:76-81,
            # SURVEY.md F10) and is never taken by its own callers; refuse rather than guess
            raise NotImplementedError("norm_img=True is not supported (unreachable and buggy in the reference)")
        if ref_cross_imgs is None:
            raise ValueError("ref_cross_imgs is required (do_reference_cross=True)")
        self._check_inputs(query_img, ref_cross_imgs)
        eng = self._engine(query_img.device)
        score, probs = eng.forward(query_img.contiguous(), ref_cross_imgs.contiguous(), bool(need_attn_weights),
                                   int(need_attn_weights_head_id))
        return {"score_map_ref_cross": score, "attn_weights_map_ref_cross": probs}

    @torch.no_grad()
    def get_featmaps(self, query_img, ref_cross_imgs):
        """DINOv2 patch features WITHOUT the multi-view PE (task/core.py:119-161):
        {"query": (B,P,C), "ref_cross": (B,N*P,C) | None} fp32."""
        self._check_inputs(query_img, ref_cross_imgs)
        eng = self._engine(query_img.device)
        st = torch.cuda.current_stream(query_img.device).cuda_stream
        B, _, H, W = query_img.shape
        P = (H // 14) * (W // 14)
        refs = None if ref_cross_imgs is None else ref_cross_imgs.contiguous()
        zero_pe = torch.zeros(P, 384, device=query_img.device)  # same kernels, PE table of zeros
        with eng.serialise():
            xq32, mem = eng.features(query_img.contiguous(), refs, st, pe_table=zero_pe)
            return {"query": xq32.view(B, P, 384).clone(),
                    "ref_cross": None if mem is None else mem.float().view(B, -1, 384).clone()}

    @staticmethod
    def _check_inputs(query_img, ref_cross_imgs):
        if query_img.dim() != 4 or query_img.shape[1] != 3:
            raise ValueError(f"query_img must be (B,3,H,W), got {tuple(query_img.shape)}")
        if not query_img.is_cuda:
            raise RuntimeError("crossscore_b200 runs on CUDA (B200) tensors only; there is no CPU path")
        if query_img.dtype != torch.float32:
            raise TypeError("query_img must be float32 (ImageNet-normalised, as the dataloader provides)")
        H, W = query_img.shape[-2:]
        if H < 14 or W < 14:
            raise ValueError("image smaller than one 14x14 patch")
        if ref_cross_imgs is not None:
            if ref_cross_imgs.dim() != 5 or ref_cross_imgs.shape[0] != query_img.shape[0] \
                    or tuple(ref_cross_imgs.shape[2:]) != (3, H, W) or ref_cross_imgs.shape[1] < 1:
                raise ValueError(f"ref_cross_imgs must be (B,N>=1,3,{H},{W}), got {tuple(ref_cross_imgs.shape)}")
            if ref_cross_imgs.dtype != torch.float32 or ref_cross_imgs.device != query_img.device:
                raise TypeError("ref_cross_imgs must be float32 on the same device as query_img")


def strip_lightning_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Lightning stores the net as ``self.model`` (task/core.py:173) -> keys carry ``model.``."""
    if any(k.startswith("model.") for k in sd):
        return {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    return dict(sd)


def load_checkpoint(net: CrossScoreNet, path_or_dict, strict: bool = True, allow_pickle: bool = False):
    """Load ``CrossScore-v1.0.0.ckpt`` (Lightning dict with "state_dict") or a bare state_dict.

    Files are read with ``weights_only=True`` (tensors and plain containers: what a Lightning checkpoint saved with
    ``save_hyperparameters(OmegaConf.to_container(...))`` holds, task/core.py:170).  ``allow_pickle=True`` opts into
    full unpickling, which executes code from the file: only for checkpoints you trust."""
    obj = path_or_dict
    if isinstance(obj, (str, bytes)) or hasattr(obj, "__fspath__"):
        obj = torch.load(obj, map_location="cpu", weights_only=not allow_pickle)
    sd = obj["state_dict"] if isinstance(obj, dict) and "state_dict" in obj else obj
    return net.load_state_dict(strip_lightning_prefix(sd), strict=strict)
