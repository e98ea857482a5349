"""GPU parity of the pre-/post-processing kernels, through the C ABI, against the
golden vectors made by the reference's own calls (torchvision Resize+Normalize, utils/io/images.py) and the numpy
oracle.  Float work (resize / normalise / mean): tolerance stated per test; integer work (uint16 quantisation,
colour indices): bit-exact."""
import os

import numpy as np
import pytest
import torch

from crossscore_b200 import imgproc
from oracle import imgproc_oracle as IO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_preprocess_matches_torchvision_golden():
    g = np.load(os.path.join(GOLD, "imgproc_pre.npz"))
    for k in range(int(g["n"])):
        u8 = torch.from_numpy(g[f"u8_{k}"]).to(DEV)
        got = imgproc.preprocess_u8(u8, int(g[f"size_{k}"]))[0].cpu().numpy()
        want = g[f"out_{k}"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-6, (k, np.abs(got - want).max())  # fp32: a few ulp of |x| <= 2.7


@pytest.mark.parametrize("H,W,size", [(540, 960, 518), (1080, 1920, 518), (518, 518, 518), (300, 200, 518), (777, 1234, 224),
                                      (33, 35, -1), (61, 47, 47), (2160, 3840, 518)])
def test_preprocess_matches_oracle_batched(H, W, size):
    rng = np.random.default_rng(H + W)
    u8 = rng.integers(0, 256, size=(2, H, W, 3), dtype=np.uint8)
    got = imgproc.preprocess_u8(torch.from_numpy(u8).to(DEV), size).cpu().numpy()
    for i in range(2):
        want = IO.preprocess(u8[i], size)
        assert got[i].shape == want.shape
        assert np.abs(got[i] - want).max() <= 2e-6


def test_preprocess_feeds_the_model_shape():
    u8 = torch.randint(0, 256, (3, 270, 480, 3), dtype=torch.uint8, device=DEV)
    x = imgproc.preprocess_u8(u8, 112)
    assert tuple(x.shape) == (3, 3, 112, 199) and x.dtype == torch.float32 and x.is_contiguous()
    with pytest.raises(ValueError):
        imgproc.preprocess_u8(u8.float(), 112)


@pytest.mark.parametrize("shape", [(3, 518, 518), (2, 37, 44), (1, 33, 35)])
def test_postprocess_matches_oracle(shape):
    rng = np.random.default_rng(7)
    m = rng.random(shape, dtype=np.float32)
    m[0, 0, :6] = [0.0, 1.0, 0.5, 1.0 / 65535, 0.999999, 0.25]
    s = torch.from_numpy(m).to(DEV)
    out = imgproc.postprocess_scores(s, mean=True, gray16_vrange=[0, 1], rgb_vrange=(0, 1))
    assert np.allclose(out["mean"].cpu().numpy(), IO.frame_mean(m), rtol=1e-6, atol=0)
    q = out["gray16"].cpu().numpy()
    assert np.array_equal(q.astype(np.int32), IO.metric_map_quantise(m, [0, 1]) & 0xFFFF)   # bit-exact
    assert np.array_equal(out["rgb"].cpu().numpy(), IO.gray2rgb_turbo(m, (0, 1)))          # bit-exact
    # [-1, 1] maps (ssim with min -1)
    m2 = (m * 2 - 1).astype(np.float32)
    out2 = imgproc.postprocess_scores(torch.from_numpy(m2).to(DEV), mean=False, gray16_vrange=[-1, 1], rgb_vrange=(-1, 1))
    assert "mean" not in out2
    assert np.array_equal(out2["gray16"].cpu().numpy().astype(np.int32), IO.metric_map_quantise(m2, [-1, 1]) & 0xFFFF)
    assert np.array_equal(out2["rgb"].cpu().numpy(), IO.gray2rgb_turbo(m2, (-1, 1)))


def test_postprocess_matches_reference_golden_quantisation():
    g = np.load(os.path.join(GOLD, "imgproc_post.npz"))
    for name, vr in (("01", [0, 1]), ("11", [-1, 1])):
        m = g[f"m_{name}"]
        pad = np.zeros((1, 40, 44), np.float32)  # H*W % 4 == 0 is only needed for B > 1; golden maps are 37 x 41
        out = imgproc.postprocess_scores(torch.from_numpy(m[None]).to(DEV), mean=False, gray16_vrange=vr)
        assert np.array_equal(out["gray16"][0].cpu().numpy().astype(np.int32), g[f"q_{name}"] & 0xFFFF)
        del pad
    with pytest.raises(ValueError):
        imgproc.postprocess_scores(torch.zeros(1, 4, 4, device=DEV), gray16_vrange=[0, 2])


def test_postprocess_edge_values():
    m = torch.tensor([[[-0.5, 0.0, 1.0, 1.5], [float("nan"), 0.5, 0.25, 0.75]]], device=DEV)
    out = imgproc.postprocess_scores(m, mean=False, rgb_vrange=(0, 1))
    assert np.array_equal(out["rgb"].cpu().numpy(), IO.gray2rgb_turbo(m.cpu().numpy(), (0, 1)))


def test_abi_argument_errors():
    """Error convention of the new entries: negative status + message, never a crash (include/crossscore_b200.h)."""
    import ctypes
    from crossscore_b200 import _lib
    from crossscore_b200._lib import call
    st = torch.cuda.current_stream().cuda_stream
    u8 = torch.zeros(1, 600, 600, 3, dtype=torch.uint8, device=DEV)
    out = torch.empty(1, 3, 40, 40, device=DEV)
    ms = (ctypes.c_float * 6)(0.5, 0.5, 0.5, 0.2, 0.2, 0.2)
    with pytest.raises(_lib.XsError, match="too large"):       # 15x down-scaling: beyond the tap table
        call("xs_preprocess_u8_resize_normalize", u8.data_ptr(), 1, 600, 600, out.data_ptr(), 40, 40, ms, st)
    with pytest.raises(_lib.XsError):
        call("xs_preprocess_u8_resize_normalize", u8.data_ptr(), 0, 600, 600, out.data_ptr(), 40, 40, ms, st)
    s = torch.zeros(1, 8, 8, device=DEV)
    with pytest.raises(_lib.XsError, match="vrange_mode"):
        call("xs_score_postprocess", s.data_ptr(), 1, 8, 8, None, None, 2, None, 0.0, 1.0, None, 0, st)
    with pytest.raises(_lib.XsError, match="workspace"):       # frame means need the partial-sum workspace
        m = torch.empty(1, device=DEV)
        call("xs_score_postprocess", s.data_ptr(), 1, 8, 8, m.data_ptr(), None, 0, None, 0.0, 1.0, None, 0, st)
    with pytest.raises(_lib.XsError, match="colour range"):
        rgb = torch.empty(1, 8, 8, 3, dtype=torch.uint8, device=DEV)
        call("xs_score_postprocess", s.data_ptr(), 1, 8, 8, None, None, 0, rgb.data_ptr(), 1.0, 1.0, None, 0, st)
    ptrs = torch.zeros(2, dtype=torch.int64, device=DEV)
    o = torch.empty(4, 384, device=DEV)
    with pytest.raises(_lib.XsError, match="overlaps"):
        call("xs_lse_merge_peers", ptrs.data_ptr(), 0, 10, o.data_ptr(), None, 2, 1, 4, 8, 48, 1, st)


def test_lse_merge_peers_single_process():
    """The peer-pointer merge on local buffers (pointer array of two ordinary allocations) == xs_lse_merge."""
    from crossscore_b200._lib import call
    torch.manual_seed(1)
    R, B, P, Hh, d = 2, 2, 37, 8, 48
    Cc = Hh * d
    part = B * P * Cc + B * Hh * P
    bufs = [torch.randn(2 * part, device=DEV) for _ in range(R)]          # two "layers" per buffer; use layer 1
    for b in bufs:
        b[part + B * P * Cc:] *= 3
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    got, lse_got = torch.empty(B * P, Cc, device=DEV), torch.empty(B, Hh, P, device=DEV)
    call("xs_lse_merge_peers", ptrs.data_ptr(), part, B * P * Cc, got.data_ptr(), lse_got.data_ptr(), R, B, P, Hh, d, 1, st)
    packed = torch.stack([b[part:] for b in bufs]).contiguous()
    want, lse_want = torch.empty(B * P, Cc, device=DEV), torch.empty(B, Hh, P, device=DEV)
    call("xs_lse_merge", packed.data_ptr(), packed.data_ptr() + B * P * Cc * 4, want.data_ptr(), lse_want.data_ptr(),
         R, B, P, Hh, d, part, part, 1, st)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(lse_got, lse_want)
