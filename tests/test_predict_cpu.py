"""CPU tests of the predict runner's host logic (crossscore_b200/predict.py): scene listing, reference sampling,
sharding, output naming, the reference-image cache and the shared-reference scheduling, driven through a fake
backend (the real one needs a GPU and is covered by tests/test_predict_gpu.py)."""
import csv
import os
import sys
import types

import numpy as np
import pytest

from crossscore_b200 import predict as P


def test_select_references_matches_reference_sampler():
    """Same draws as utils/neighbour/sampler.py::SamplerRandom under the same numpy seed (imported from the
    reference when it is mounted; the restated expectations otherwise)."""
    refs = [f"r{i}.png" for i in range(9)]
    if os.path.isdir("/root/reference"):
        sys.path.insert(0, "/root/reference")
        try:
            from utils.neighbour.sampler import SamplerFactory
        finally:
            sys.path.pop(0)
        for det in (False, True):
            np.random.seed(7)
            want = [SamplerFactory("random", 5, det)(None, refs) for _ in range(4)]
            rng = np.random.RandomState(7)
            got = [P.select_references(refs, 5, det, rng) for _ in range(4)]
            assert got == want
        np.random.seed(3)
        want = SamplerFactory("random", 12, False)(None, refs)
        assert P.select_references(refs, 12, False, np.random.RandomState(3)) == want
    assert P.select_references(refs, 3, True) == refs[:3]
    padded = P.select_references(refs[:2], 5, False, np.random.RandomState(0))
    assert sorted(padded) == sorted(refs[:2] + [P.EMPTY] * 3)


def test_naming_and_out_dir():
    assert P.score_map_file_name("/data/a/b/c/d/e/frame_0007.png", 1, 12, 3) == "r1_B0012_b003_b_c_d_e_frame_0007.png"
    assert P.predict_out_dir(None, None, "", now="T") == "log/T/predict_empty_ckpt/T"
    assert P.predict_out_dir("log/run1/ckpt/last.ckpt", None, "x", now="T") == "log/run1/predict/T_x"
    assert P.predict_out_dir("log/run1/ckpt/last.ckpt", "/tmp/o", "", now="T") == "/tmp/o"
    assert P.metric_type_str("ssim", 0) == "ssim_0_1" and P.metric_type_str("ssim", -1) == "ssim_-1_1"
    assert P.metric_type_str("mae", 0) == "mae"
    assert P.intrinsic_vrange("ssim") == [-1, 1] and P.intrinsic_vrange("mse") == [0, 1]
    with pytest.raises(ValueError):
        P.intrinsic_vrange("psnr")
    assert P.summary_row("m/ds/scene/split/ours/renders/frame_00012.png", 0.5) == \
        ["ds", "m/ds/scene/split/ours", "00012.png", 0.5][0:0] + ["scene", "m/ds/scene/split/ours", "00012.png", 0.5]


def test_shard_indices_cover_all_queries_once():
    for n in (0, 1, 7, 16):
        for world in (1, 2, 8):
            allidx = sorted(i for r in range(world) for i in P.shard_indices(n, r, world))
            assert allidx == list(range(n))


def test_byte_lru():
    c = P.ByteLRU(100)
    c.put("a", 1, 40); c.put("b", 2, 40)
    assert c.value("a") == 1
    c.put("c", 3, 40)            # evicts b (least recently used)
    assert c.value("b") is None and c.value("a") == 1 and c.value("c") == 3
    c.put("huge", 4, 1000)       # larger than the cache: not stored
    assert c.value("huge") is None and c.used == 80


class FakeBackend:
    """numpy stand-in: 'preprocess' = mean over a 2x2 pooling, 'forward' = query mean + reference mean."""

    def __init__(self):
        self.pre_calls, self.fwd_calls, self.scene_builds = 0, 0, 0

    def preprocess(self, u8, size):
        self.pre_calls += u8.shape[0]
        return u8.astype(np.float32).transpose(0, 3, 1, 2) / 255.0

    def stack(self, ts):
        return np.stack(list(ts), 0)

    def forward(self, q, r):
        self.fwd_calls += 1
        B, _, H, W = q.shape
        return np.broadcast_to((q.mean(axis=(1, 2, 3)) + r.mean(axis=(1, 2, 3, 4)))[:, None, None] / 2, (B, H, W)).copy()

    def scene_scorer(self):
        be = self

        class S:
            def build_reference_cache(self, refs):
                be.scene_builds += 1
                self.m = refs.mean()

            def score(self, q):
                B, _, H, W = q.shape
                return types.SimpleNamespace(clone=lambda: np.broadcast_to(
                    ((q.mean(axis=(1, 2, 3)) + self.m) / 2)[:, None, None], (B, H, W)).copy())
        return S()

    def postprocess(self, score, gray_vrange, rgb_vrange):
        out = {"mean": score.mean(axis=(1, 2)).astype(np.float32)}
        if gray_vrange is not None:
            out["gray16"] = (score * 65535).astype(np.uint16)
        if rgb_vrange is not None:
            out["rgb"] = np.repeat((score * 255).astype(np.uint8)[..., None], 3, -1)
        return out

    def nbytes(self, t):
        return t.nbytes


def _scene(tmp_path, nq=5, nr=4):
    from PIL import Image
    rng = np.random.default_rng(0)
    qd, rd = tmp_path / "m" / "ds" / "scene" / "test" / "ours" / "renders", tmp_path / "m" / "ds" / "scene" / "train" / "ours" / "gt"
    qd.mkdir(parents=True); rd.mkdir(parents=True)
    for i in range(nq):
        Image.fromarray(rng.integers(0, 256, (28, 42, 3), dtype=np.uint8)).save(qd / f"frame_{i:05}.png")
    for i in range(nr):
        Image.fromarray(rng.integers(0, 256, (28, 42, 3), dtype=np.uint8)).save(rd / f"frame_{i:05}.png")
    return str(qd), str(rd)


@pytest.mark.parametrize("colour_mode", ["rgb", "gray"])
def test_runner_writes_maps_and_summary(tmp_path, colour_mode):
    qd, rd = _scene(tmp_path)
    q, r = P.list_scene(qd, rd)
    assert [os.path.basename(x) for x in q] == [f"frame_{i:05}.png" for i in range(5)]
    be = FakeBackend()
    run = P.PredictRunner(be, str(tmp_path / "out"), "ssim", 0, 1, batch_size=2, num_refs=3, resize_short_side=-1,
                          colour_mode=colour_mode, seed=1)
    rows = run.run(q, r)
    assert len(rows) == 5 and be.fwd_calls == 3
    # every reference image was decoded + preprocessed at most once (device cache), every query exactly once
    assert be.pre_calls <= 5 + 4 and run.cache.hits > 0
    maps = sorted(os.listdir(tmp_path / "out" / "batch" / "score_map_ref_cross"))
    assert len(maps) == 5 and maps[0].startswith("r0_B0000_b000_") and maps[-1].startswith("r0_B0002_b000_")
    from PIL import Image
    im = Image.open(tmp_path / "out" / "batch" / "score_map_ref_cross" / maps[0])
    assert im.mode == ("RGB" if colour_mode == "rgb" else "I;16") and im.size == (42, 28)
    path = run.write_summary()
    got = list(csv.reader(open(path)))
    assert got[0] == ["scene_name", "rendered_dir", "image_name", "pred_ssim_0_1"]
    assert [g[2] for g in got[1:]] == [f"{i:05}.png" for i in range(5)] and all(len(g[3].split(".")[1]) == 4 for g in got[1:])
    assert got[1][0] == "scene"


def test_runner_shared_reference_scene_and_ranks(tmp_path):
    qd, rd = _scene(tmp_path, nq=7, nr=6)
    q, r = P.list_scene(qd, rd)
    rows = []
    for rank in range(2):
        be = FakeBackend()
        run = P.PredictRunner(be, str(tmp_path / "out"), "mae", 0, 1, batch_size=4, num_refs=5, deterministic_refs=True,
                              resize_short_side=-1, colour_mode="gray", rank=rank, world=2, write_maps=False)
        rows += run.run(q, r)
        assert be.scene_builds == 1 and be.fwd_calls == 0        # one reference encode per scene, not per batch
        assert be.pre_calls == len(P.shard_indices(7, rank, 2)) + 5
        assert run.write_summary().name == f"scores_r{rank}.csv"
    assert sorted(x[2] for x in rows) == [f"{i:05}.png" for i in range(7)]
    # more ranks than query frames: a rank with an EMPTY shard still takes part in the one collective
    # step of the scene -- the reference cache build -- instead of leaving its peers waiting for its K/V slice
    for rank in range(9):
        be = FakeBackend()
        run = P.PredictRunner(be, str(tmp_path / "o9"), "mae", 0, 1, batch_size=4, num_refs=5, deterministic_refs=True,
                              resize_short_side=-1, colour_mode="gray", rank=rank, world=9, write_maps=False)
        got = run.run(q, r)
        assert be.scene_builds == 1, rank
        assert len(got) == len(P.shard_indices(7, rank, 9))
    # fewer references than requested: empty_image padding -> per-query path (no shared cache)
    be = FakeBackend()
    run = P.PredictRunner(be, str(tmp_path / "o2"), "mae", 0, 1, batch_size=4, num_refs=8, deterministic_refs=True,
                          resize_short_side=-1, write_maps=False)
    run.run(q, r)
    assert be.scene_builds == 0 and be.fwd_calls == 2
