"""Generate golden vectors for the image pre-/post-processing oracle (build container only).

Preprocessing: the exact third-party calls the reference makes -- utils/io/images.py::f32 (imported from
/root/reference), torchvision T.Resize(short side, BILINEAR, antialias=True) and T.Normalize(ImageNet), as wired in
task/predict.py:69-93 and dataloading/dataset/nvs_dataset.py:218-225,242-279 -- run on small seeded uint8 images.
Post-processing: utils/io/images.py::metric_map_write imported from /root/reference with imageio.imwrite stubbed to
capture the integer array it would write.  (gray2rgb needs matplotlib, which is absent: not generated.)

    python tests/golden/make_golden_imgproc.py      # rewrites tests/golden/imgproc_*.npz
"""
import os
import sys
import types

import numpy as np
import torch
from torchvision.transforms import v2 as T

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def boot():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    captured = {}
    sys.modules["imageio"] = types.SimpleNamespace(imwrite=lambda p, m: captured.__setitem__("m", np.array(m)))
    from utils.io import images  # the reference's own module
    return images, captured


def main():
    images, captured = boot()
    mean, std = images.ImageNetMeanStd.mean, images.ImageNetMeanStd.std
    rng = np.random.default_rng(0)
    pre = {}
    # (H, W, resize_short_side): down 2.4x, portrait, identity, upscale, non-integer ratio, no resize
    for k, (H, W, s) in enumerate([(135, 240, 56), (60, 47, 42), (70, 70, 70), (33, 80, 70), (108, 192, 49), (28, 42, -1)]):
        u8 = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        # smooth structure in half of the cases so that the filter taps matter beyond noise
        if k % 2 == 0:
            yy, xx = np.mgrid[0:H, 0:W]
            u8[..., 0] = (127 + 120 * np.sin(xx / 7.0) * np.cos(yy / 5.0)).astype(np.uint8)
        x = torch.tensor(images.f32(u8)).permute(2, 0, 1)
        if s > 0:
            x = T.Resize(s, interpolation=T.InterpolationMode.BILINEAR, antialias=True)(x)
        x = T.Normalize(mean=mean, std=std)(x)
        pre[f"u8_{k}"] = u8
        pre[f"size_{k}"] = np.int64(s)
        pre[f"out_{k}"] = x.numpy().astype(np.float32)
    pre["n"] = np.int64(6)
    np.savez_compressed(os.path.join(HERE, "imgproc_pre.npz"), **pre)

    post = {}
    m01 = rng.random((37, 41), dtype=np.float32)
    m01[0, :4] = [0.0, 1.0, 0.5, 1.0 / 65535]
    m11 = (rng.random((37, 41), dtype=np.float32) * 2 - 1).astype(np.float32)
    m11[0, :3] = [-1.0, 1.0, 0.0]
    for name, m, vr in (("01", m01, [0, 1]), ("11", m11, [-1, 1])):
        images.metric_map_write("unused.png", m.copy(), vr)
        post[f"m_{name}"] = m
        post[f"q_{name}"] = captured["m"].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "imgproc_post.npz"), **post)
    print("wrote imgproc_pre.npz, imgproc_post.npz")


if __name__ == "__main__":
    main()
