"""Device-side pre- and post-processing around the CrossScore forward.

`preprocess_u8` does what the reference's dataloader does per image on CPU workers -- utils/io/images.py:14-29
(uint8 -> fp32 / 255), torchvision T.Resize(short side, BILINEAR, antialias=True) (task/predict.py:87-92,
dataloading/dataset/nvs_dataset.py:218-225) and T.Normalize(ImageNet) (task/predict.py:69-74) -- in one CUDA kernel,
so only the uint8 pixels cross PCIe.  `postprocess_scores` produces what utils/io/batch_writer.py and
utils/io/score_summariser.py derive from the score maps (frame means, uint16 gray maps, turbo RGB) before the
device -> host copy.  No CPU fallback: both call the C ABI.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call

IMAGENET_MEAN_STD = (0.485, 0.456, 0.406, 0.229, 0.224, 0.225)  # utils/io/images.py:8-11


def resize_output_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision T.Resize(int size): short side -> size, long side -> int(size * long / short)."""
    if size <= 0:
        return h, w
    short, long_ = (h, w) if h <= w else (w, h)
    new_long = int(size * long_ / short)
    return (size, new_long) if h <= w else (new_long, size)


def preprocess_u8(images: torch.Tensor, resize_short_side: int = -1,
                  mean_std: Sequence[float] = IMAGENET_MEAN_STD, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """images (n, H0, W0, 3) or (H0, W0, 3) uint8 on a CUDA device -> (n, 3, H1, W1) fp32 normalised."""
    if images.dim() == 3:
        images = images[None]
    if not images.is_cuda or images.dtype != torch.uint8 or images.shape[-1] != 3:
        raise ValueError("preprocess_u8 expects a CUDA uint8 tensor of shape (n, H, W, 3)")
    images = images.contiguous()
    n, H0, W0, _ = images.shape
    H1, W1 = resize_output_size(H0, W0, resize_short_side)
    if out is None:
        out = torch.empty(n, 3, H1, W1, device=images.device, dtype=torch.float32)
    elif tuple(out.shape) != (n, 3, H1, W1) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous fp32 tensor of shape {(n, 3, H1, W1)}")
    ms = (ctypes.c_float * 6)(*[float(v) for v in mean_std])
    with torch.cuda.device(images.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.count_launch(1)
        call("xs_preprocess_u8_resize_normalize", images.data_ptr(), n, H0, W0, out.data_ptr(), H1, W1, ms, st)
    return out


def postprocess_scores(score: torch.Tensor, mean: bool = True, gray16_vrange: Optional[Sequence[int]] = None,
                       rgb_vrange: Optional[Sequence[float]] = None) -> Dict[str, torch.Tensor]:
    """score (B, H, W) fp32 CUDA -> {"mean": (B,) fp32, "gray16": (B,H,W) uint16, "rgb": (B,H,W,3) uint8}.
    gray16_vrange: [0, 1] or [-1, 1] (utils/io/images.py:49-63; anything else raises ValueError like the reference);
    rgb_vrange: (vmin, vmax) of the turbo colour map (utils/misc/image.py:35-49)."""
    if not score.is_cuda or score.dtype != torch.float32 or score.dim() != 3:
        raise ValueError("postprocess_scores expects a CUDA fp32 tensor of shape (B, H, W)")
    score = score.contiguous()
    B, H, W = score.shape
    dev = score.device
    mode = 0
    if gray16_vrange is not None:
        if list(gray16_vrange) == [0, 1]:
            mode = 0
        elif list(gray16_vrange) == [-1, 1]:
            mode = 1
        else:
            raise ValueError("Invalid range for metric map writing. Must be '[0,1]' or '[-1,1]'")
    res: Dict[str, torch.Tensor] = {}
    mean_t = torch.empty(B, device=dev, dtype=torch.float32) if mean else None
    gray_t = torch.empty(B, H, W, device=dev, dtype=torch.uint16) if gray16_vrange is not None else None
    rgb_t = torch.empty(B, H, W, 3, device=dev, dtype=torch.uint8) if rgb_vrange is not None else None
    vmin, vmax = (float(rgb_vrange[0]), float(rgb_vrange[1])) if rgb_vrange is not None else (0.0, 1.0)
    ws_bytes = _lib.load().xs_workspace_bytes(_lib.OP_SCORE_POSTPROCESS, B, 0, 0, 0) if mean else 0
    ws = torch.empty(max(ws_bytes, 1), device=dev, dtype=torch.uint8)
    p = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _lib.count_launch(2 if mean else 1)
        call("xs_score_postprocess", score.data_ptr(), B, H, W, p(mean_t), p(gray_t), mode, p(rgb_t), vmin, vmax,
             ws.data_ptr(), ws_bytes, st)
    if mean_t is not None:
        res["mean"] = mean_t
    if gray_t is not None:
        res["gray16"] = gray_t
    if rgb_t is not None:
        res["rgb"] = rgb_t
    return res
