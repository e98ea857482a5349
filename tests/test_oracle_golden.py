"""The oracle is pinned against outputs of the reference itself (tests/golden/*.npz,
produced by tests/golden/make_golden.py from /root/reference in the build container)."""
import numpy as np
import pytest
import torch

from oracle import crossscore_oracle as O
from helpers import GOLDEN_CASES, compare_to_golden, golden_pos_interp, golden_problem, load_golden, oracle_kwargs

SMALL = [c for c in GOLDEN_CASES if "518" not in c and "1036" not in c]


@pytest.mark.parametrize("case", SMALL)
def test_oracle_matches_reference_fp64(case):
    rec = load_golden(case)
    sd, q, r = golden_problem(rec)
    out = O.crossscore_forward(sd, q, r, need_attn_weights=bool(rec["need_w"]), head_id=int(rec["head_id"]),
                               dt=torch.float64, pos_interp=golden_pos_interp(rec), **oracle_kwargs(rec["cfg_over"]))
    mx, mean = compare_to_golden(out["score_map_ref_cross"], rec)
    # the reference ran in fp32; fp64 restatement agrees to fp32 round-off
    assert mx < 2e-5 and mean < 2e-6, (mx, mean)
    assert np.allclose(out["score_map_ref_cross"].mean(dim=(-1, -2)).numpy(), rec["score_mean"], atol=2e-6)
    if rec["need_w"]:
        a = out["attn_weights_map_ref_cross"].float().numpy()
        assert a.shape == rec["attn"].shape
        assert np.abs(a - rec["attn"]).max() < 1e-6
    if "feat_query" in rec:
        assert np.abs(out["_featmap_query"].numpy() - 0).max() > 0  # sanity
        fq, fr = O.get_featmaps(sd, q, r, torch.float64, golden_pos_interp(rec))
        # fp32 reference noise scales with the feature magnitude (outlier-channel weights: |f| up to ~100)
        ftol = 5e-5 * max(1.0, float(np.abs(rec["feat_query"]).max()) / 10.0)
        assert np.abs(fq.float().numpy() - rec["feat_query"]).max() < ftol
        assert np.abs(fr.float().numpy() - rec["feat_ref"]).max() < ftol


def test_oracle_matches_reference_518_fp32():
    """Headline shape (cfg 1: 1 query + 5 refs, 518x518), fp32 oracle vs fp32 reference."""
    rec = load_golden("g3_518_n5")
    sd, q, r = golden_problem(rec)
    out = O.crossscore_forward(sd, q, r, dt=torch.float32)
    mx, mean = compare_to_golden(out["score_map_ref_cross"], rec)
    assert mx < 1e-4 and mean < 1e-5, (mx, mean)


def test_fast_cpu_variant_matches_reference():
    """bench.py times the oracle with torch's fused CPU ops (O.FAST); same results as the spelled-out path."""
    rec = load_golden("g2_nonsquare_84x117_n3_attn")
    sd, q, r = golden_problem(rec)
    O.FAST = True
    try:
        out = O.crossscore_forward(sd, q, r, dt=torch.float32, pos_interp=golden_pos_interp(rec))
    finally:
        O.FAST = False
    mx, mean = compare_to_golden(out["score_map_ref_cross"], rec)
    assert mx < 1e-4 and mean < 1e-5, (mx, mean)


def test_resamplers_match_torch_interpolate():
    """The spelled-out bicubic / bilinear restatements equal F.interpolate (the library call
    the reference makes: positional_encoding.py:61-69, modeling_dinov2.py:86-91)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    t = torch.randn(37, 37, 16, generator=g, dtype=torch.float64)
    for oh, ow in [(5, 5), (6, 8), (37, 49), (74, 74), (12, 11)]:
        want = F.interpolate(t.permute(2, 0, 1)[None], size=(oh, ow), mode="bicubic", align_corners=False)[0].permute(1, 2, 0)
        assert (O.bicubic_resize_ac_false(t, oh, ow) - want).abs().max() < 1e-12
    # transformers 4.33.3 (the reference's pinned version): scale_factor=((oh + 0.1) / 37, (ow + 0.1) / 37)
    for oh, ow in [(5, 5), (6, 8), (37, 49), (74, 74), (12, 11), (37, 65)]:
        sf = ((oh + 0.1) / 37, (ow + 0.1) / 37)
        want = F.interpolate(t.permute(2, 0, 1)[None], scale_factor=sf, mode="bicubic", align_corners=False)[0].permute(1, 2, 0)
        assert want.shape[:2] == (oh, ow)
        got = O.bicubic_resize_ac_false(t, oh, ow, 1.0 / sf[0], 1.0 / sf[1])
        assert (got - want).abs().max() < 1e-12
        assert (got - O.bicubic_resize_ac_false(t, oh, ow)).abs().max() > 1e-3  # and it is NOT the size= form
    t = torch.randn(40, 40, 16, generator=g, dtype=torch.float64)
    for oh, ow in [(5, 5), (6, 8), (37, 37), (37, 49), (74, 74)]:
        want = F.interpolate(t.permute(2, 0, 1)[None], scale_factor=((oh + 1e-4) / 40, (ow + 1e-4) / 40),
                             mode="bilinear", align_corners=True)[0].permute(1, 2, 0)
        assert want.shape[:2] == (oh, ow)
        assert (O.bilinear_resize_ac_true(t, oh, ow) - want).abs().max() < 1e-12


def test_regression_layer_behaviour_table():
    """model/regression_layer.py:65-81 prints this grid; the accept/reject pattern is pinned."""
    from crossscore_b200.config import resolve_score_activation
    assert resolve_score_activation("ssim", 0, 1, "default") == (False, 1.0)
    assert resolve_score_activation("mae", 0, 1, "default") == (False, 2.0)
    assert resolve_score_activation("mse", 0, 1, "default") == (False, 4.0)
    assert resolve_score_activation("ssim", -1, 1, 5) == (True, 1.0)
    assert resolve_score_activation("ssim", 0, 1, 1.5) == (False, 1.5)
    for bad in [("mae", -1, 1, "default"), ("mse", -1, 1, 1), ("psnr", 0, 1, 1), ("ssim", 0, 2, 1)]:
        with pytest.raises(ValueError):
            resolve_score_activation(*bad)
    with pytest.raises(ValueError):
        resolve_score_activation("ssim", 0, 1, "some_typo")


def test_lse_merge_identity():
    g = torch.Generator().manual_seed(0)
    s = torch.randn(7, 50, generator=g, dtype=torch.float64)
    v = torch.randn(50, 4, generator=g, dtype=torch.float64)
    full = torch.softmax(s, -1) @ v
    parts, lses = [], []
    for a, b in [(0, 13), (13, 14), (14, 50)]:
        lses.append(torch.logsumexp(s[:, a:b], -1))
        parts.append(torch.softmax(s[:, a:b], -1) @ v[a:b])
    o, lse = O.lse_merge(parts, lses)
    assert (o - full).abs().max() < 1e-12
    assert (lse - torch.logsumexp(s, -1)).abs().max() < 1e-12
