"""CPU oracle (TEST INFRASTRUCTURE, not product code) for the steps either side of the CrossScore hot path
(SURVEY.md section 8f rows 2 and 3): image preprocessing in front of the model and score-map post-processing
behind it.  numpy restatements; every function cites the reference lines it follows.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module.

Pinning (tests/golden/make_golden_imgproc.py, tests/test_imgproc_oracle.py):
  * preprocessing is pinned to torchvision's own T.Resize(antialias=True) + T.Normalize -- the calls
    task/predict.py:69-93 makes -- executed in the build container (golden fixtures tests/golden/imgproc_*.npz);
  * uint16 quantisation is pinned to the reference's utils/io/images.py::metric_map_write (imported, with
    imageio.imwrite stubbed to capture the array it is given);
  * the colour map (utils/misc/image.py::gray2rgb -> matplotlib Normalize + cm.get_cmap("turbo")) is
    **parity unpinned**: matplotlib is not installed here and cannot be run; the restatement follows matplotlib's
    published Colormap.__call__ algorithm and the published 256-entry turbo table (recovered from OpenCV's copy of
    the same Google table, oracle/turbo_table.npy; OpenCV's own 8-bit LUT is reproduced exactly by round(255 t)).
"""
from __future__ import annotations

import os

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # utils/io/images.py:8-11
IMAGENET_STD = (0.229, 0.224, 0.225)

_f32 = np.float32


def image_u8_to_f32(img_u8: np.ndarray) -> np.ndarray:
    """utils/io/images.py:14-17 (f32): astype(float32) / 255.0"""
    return img_u8.astype(np.float32) / 255.0


def resize_output_size(h: int, w: int, size: int):
    """torchvision T.Resize(int): the short side becomes `size`, the long side int(size * long / short)
    (call site task/predict.py:87-92)."""
    short, long_ = (h, w) if h <= w else (w, h)
    new_long = int(size * long_ / short)
    return (size, new_long) if h <= w else (new_long, size)


def aa_taps(in_size: int, out_size: int):
    """Per output index: (first input index, normalised triangle-filter weights), fp32 arithmetic.
    ATen `_upsample_bilinear2d_aa` (HelperInterpBase::_compute_indices_min_size_weights_aa, align_corners=False),
    reached from T.Resize(interpolation=BILINEAR, antialias=True)."""
    scale = _f32(in_size) / _f32(out_size)
    support = _f32(1.0) * scale if scale >= 1 else _f32(1.0)
    invscale = _f32(1.0) / scale if scale >= 1 else _f32(1.0)
    taps = []
    for i in range(out_size):
        center = scale * (_f32(i) + _f32(0.5))
        xmin = max(int(center - support + _f32(0.5)), 0)
        xsize = min(int(center + support + _f32(0.5)), in_size) - xmin
        w = np.array([max(_f32(0), _f32(1) - abs((_f32(j + xmin) - center + _f32(0.5)) * invscale))
                      for j in range(xsize)], dtype=np.float32)
        tot = w.sum(dtype=np.float32)
        if tot != 0:
            w = w / tot
        taps.append((xmin, w))
    return taps


def resize_bilinear_aa(x: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """x (C,H,W) fp32 -> (C,oh,ow): separable, width pass first, fp32 intermediate (ATen CPU kernel order)."""
    C, H, W = x.shape
    tmp = np.zeros((C, H, ow), np.float32)
    for i, (x0, w) in enumerate(aa_taps(W, ow)):
        acc = np.zeros((C, H), np.float32)
        for k, wk in enumerate(w):
            acc = acc + wk * x[:, :, x0 + k]
        tmp[:, :, i] = acc
    out = np.zeros((C, oh, ow), np.float32)
    for i, (y0, w) in enumerate(aa_taps(H, oh)):
        acc = np.zeros((C, ow), np.float32)
        for k, wk in enumerate(w):
            acc = acc + wk * tmp[:, y0 + k, :]
        out[:, i, :] = acc
    return out


def normalize(x: np.ndarray, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """torchvision T.Normalize: (x - mean) / std per channel, fp32 (task/predict.py:69-74)."""
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


def preprocess(img_u8: np.ndarray, resize_short_side: int) -> np.ndarray:
    """HWC uint8 -> normalised CHW fp32, the dataloader's per-image work:
    dataloading/dataset/nvs_dataset.py:428-446 (image_read, permute), :218-225 (resize_all), :242-279 (Normalize)."""
    x = image_u8_to_f32(img_u8).transpose(2, 0, 1)
    if resize_short_side > 0:
        oh, ow = resize_output_size(x.shape[1], x.shape[2], resize_short_side)
        if (oh, ow) != x.shape[1:]:
            x = resize_bilinear_aa(np.ascontiguousarray(x), oh, ow)
    return normalize(x)


# ---------------------------------------------------------------------------------------------------------------
def frame_mean(score: np.ndarray) -> np.ndarray:
    """utils/io/score_summariser.py:180-181: score_maps.mean(dim=[-1,-2]); accumulated here in fp64."""
    return score.astype(np.float64).mean(axis=(-1, -2)).astype(np.float32)


def metric_map_quantise(m: np.ndarray, vrange) -> np.ndarray:
    """utils/io/images.py:49-63 (metric_map_write) up to the imageio call: fp32 scale, truncation to int32
    (the PNG holds the low 16 bits)."""
    m = m.astype(np.float32)
    if list(vrange) == [0, 1]:
        m = m * _f32(65535)
    elif list(vrange) == [-1, 1]:
        m = (m + _f32(1)) * _f32(32767)
    else:
        raise ValueError("Invalid range for metric map writing. Must be '[0,1]' or '[-1,1]'")
    return m.astype(np.int32)


_TURBO = None


def turbo_table() -> np.ndarray:
    global _TURBO
    if _TURBO is None:
        _TURBO = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "turbo_table.npy"))
    return _TURBO


def gray2rgb_turbo(img: np.ndarray, vrange) -> np.ndarray:
    """utils/misc/image.py:35-49 (gray2rgb): plt.Normalize(vmin, vmax) in the image's dtype, matplotlib
    Colormap.__call__ (x*N, x==N -> N-1, truncate, <0 -> first entry, >=N -> last, NaN -> bad colour (0,0,0)),
    then u8(): rgb * 255.0 truncated (utils/io/images.py:20-23)."""
    vmin, vmax = vrange
    x = img.astype(np.float32)
    x = x - _f32(vmin)
    x = x / _f32(vmax - vmin)
    xa = x * _f32(256)
    xa = np.where(xa == 256, _f32(255), xa)
    bad = np.isnan(xa)
    idx = np.clip(np.nan_to_num(xa, nan=0.0), -1, 256).astype(np.int64)
    idx = np.where(xa < 0, 0, np.where(xa >= 256, 255, idx))
    rgb = turbo_table()[idx] * 255.0
    out = rgb.astype(np.uint8)
    out[bad] = 0
    return out
