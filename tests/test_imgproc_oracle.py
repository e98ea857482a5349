"""CPU: the image pre-/post-processing oracle against the golden vectors produced by the reference's own calls
(tests/golden/make_golden_imgproc.py): torchvision Resize(antialias)+Normalize and utils/io/images.py."""
import os

import numpy as np
import pytest

from oracle import imgproc_oracle as IO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_preprocess_matches_torchvision_golden():
    g = np.load(os.path.join(GOLD, "imgproc_pre.npz"))
    for k in range(int(g["n"])):
        got = IO.preprocess(g[f"u8_{k}"], int(g[f"size_{k}"]))
        want = g[f"out_{k}"]
        assert got.shape == want.shape, (k, got.shape, want.shape)
        # fp32 work: a few ulp of the normalised value (|x| <= 2.7)
        assert np.abs(got - want).max() <= 2e-6, (k, np.abs(got - want).max())


def test_resize_output_size_rule():
    assert IO.resize_output_size(1080, 1920, 518) == (518, 920)
    assert IO.resize_output_size(1920, 1080, 518) == (920, 518)
    assert IO.resize_output_size(518, 518, 518) == (518, 518)
    assert IO.resize_output_size(135, 240, 56) == (56, 99)


@pytest.mark.parametrize("name,vr", [("01", [0, 1]), ("11", [-1, 1])])
def test_quantise_matches_reference_bit_exact(name, vr):
    g = np.load(os.path.join(GOLD, "imgproc_post.npz"))
    got = IO.metric_map_quantise(g[f"m_{name}"], vr)
    assert got.dtype == np.int32 and np.array_equal(got, g[f"q_{name}"])
    with pytest.raises(ValueError):
        IO.metric_map_quantise(g[f"m_{name}"], [0, 2])


def test_turbo_table_and_colour_map_properties():
    t = IO.turbo_table()
    assert t.shape == (256, 3) and t.dtype == np.float64
    # published end points and a mid entry of the turbo table
    assert np.allclose(t[0], [0.18995, 0.07176, 0.23217]) and np.allclose(t[255], [0.47960, 0.01583, 0.01055])
    assert np.allclose(t[128], [0.64362, 0.98999, 0.23356])
    m = np.array([[-0.5, 0.0, 0.25, 0.5, 1.0, 1.5, np.nan]], np.float32)
    rgb = IO.gray2rgb_turbo(m, (0, 1))
    assert rgb.dtype == np.uint8 and rgb.shape == (1, 7, 3)
    first, last = (t[0] * 255.0).astype(np.uint8), (t[255] * 255.0).astype(np.uint8)
    assert np.array_equal(rgb[0, 0], first) and np.array_equal(rgb[0, 1], first)     # under -> first entry
    assert np.array_equal(rgb[0, 4], last) and np.array_equal(rgb[0, 5], last)       # x == 1 and over -> last
    assert np.array_equal(rgb[0, 2], (t[64] * 255.0).astype(np.uint8))
    assert np.array_equal(rgb[0, 3], (t[128] * 255.0).astype(np.uint8))
    assert np.array_equal(rgb[0, 6], [0, 0, 0])                                      # NaN -> bad colour
    # [-1, 1] visual range
    rgb2 = IO.gray2rgb_turbo(np.array([[-1.0, 0.0, 1.0]], np.float32), (-1, 1))
    assert np.array_equal(rgb2[0, 1], (t[128] * 255.0).astype(np.uint8))


def test_frame_mean():
    rng = np.random.default_rng(1)
    s = rng.random((3, 28, 42), dtype=np.float32)
    assert np.allclose(IO.frame_mean(s), s.mean(axis=(1, 2)), rtol=1e-6)


def test_cuda_turbo_lut_header_matches_table():
    """crossscore_b200/csrc/xs_turbo.h (what the kernel indexes) == floor(255 * table) (what gray2rgb + u8 give)."""
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "crossscore_b200", "csrc", "xs_turbo.h")
    body = open(path).read().split("{", 2)[2].split("}")[0]
    vals = np.array([int(v) for v in re.findall(r"\d+", body)], dtype=np.uint8).reshape(256, 3)
    assert np.array_equal(vals, (IO.turbo_table() * 255.0).astype(np.uint8))
