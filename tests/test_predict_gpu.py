"""GPU: the predict runner end to end (decode -> device preprocessing -> CrossScoreNet -> device post-processing ->
PNG / CSV) against the CPU oracle pipeline on the same files, for the per-query-reference path and the
shared-reference (scene cache) path."""
import csv
import os

import numpy as np
import pytest
import torch

from crossscore_b200 import CrossScoreNet, default_cfg
from crossscore_b200 import predict as P
from crossscore_b200.synthetic import make_state_dict
from oracle import crossscore_oracle as O
from oracle import imgproc_oracle as IO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _scene(tmp_path, nq, nr, hw=(60, 90)):
    from PIL import Image
    rng = np.random.default_rng(5)
    qd = tmp_path / "m" / "ds" / "scene" / "test" / "ours" / "renders"
    rd = tmp_path / "m" / "ds" / "scene" / "train" / "ours" / "gt"
    qd.mkdir(parents=True); rd.mkdir(parents=True)
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    for d, n in ((qd, nq), (rd, nr)):
        for i in range(n):
            a = rng.integers(0, 256, (*hw, 3), dtype=np.uint8)
            a[..., 1] = (127 + 100 * np.sin(xx / (3.0 + i)) * np.cos(yy / 4.0)).astype(np.uint8)
            Image.fromarray(a).save(d / f"frame_{i:05}.png")
    return str(qd), str(rd)


@pytest.mark.parametrize("deterministic", [False, True])
def test_predict_runner_matches_oracle_pipeline(tmp_path, deterministic):
    qd, rd = _scene(tmp_path, nq=3, nr=4)
    q_paths, r_paths = P.list_scene(qd, rd)
    sd = make_state_dict(2)
    net = CrossScoreNet(default_cfg(), precision="fp32")
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    size = 42  # 60x90 -> 42x63: 3 x 4 patches
    run = P.PredictRunner(P.DeviceBackend(net, DEV), str(tmp_path / "out"), "ssim", 0, 1, batch_size=2, num_refs=2,
                          deterministic_refs=deterministic, resize_short_side=size, colour_mode="gray", seed=11)
    with torch.inference_mode():
        rows = run.run(q_paths, r_paths)
    path = run.write_summary()
    # the oracle pipeline on the same files, same reference draws
    rng = np.random.RandomState(11)
    refs_all = [P.select_references(r_paths, 2, deterministic, rng) for _ in q_paths]
    pre = lambda p: torch.from_numpy(IO.preprocess(P.read_image_u8(p), size))
    want_means = []
    maps = sorted(os.listdir(tmp_path / "out" / "batch" / "score_map_ref_cross"))
    from PIL import Image
    for i, qp in enumerate(q_paths):
        q = pre(qp)[None]
        r = torch.stack([pre(p) for p in refs_all[i]])[None]
        score = O.crossscore_forward(sd, q, r, dt=torch.float64)["score_map_ref_cross"][0].numpy()
        want_means.append(score.mean())
        name = P.score_map_file_name(qp, 0, i // 2, i % 2)
        assert name in maps
        got_q = np.array(Image.open(tmp_path / "out" / "batch" / "score_map_ref_cross" / name)).astype(np.int64)
        want_q = IO.metric_map_quantise(score.astype(np.float32), [-1, 1]).astype(np.int64)   # ssim: intrinsic range
        assert got_q.shape == want_q.shape == (42, 56)
        assert np.abs(got_q - want_q).max() <= 8        # fp32 mode: 1e-4 on the map = 3.3 counts of 32767 (+ rounding)
    got_means = {r[2]: r[3] for r in rows}
    for i, qp in enumerate(q_paths):
        assert abs(got_means[os.path.basename(qp).replace("frame_", "")] - want_means[i]) <= 1e-4
    table = list(csv.reader(open(path)))
    assert table[0][3] == "pred_ssim_0_1" and len(table) == 4
    if deterministic:
        assert run.cache.misses == 2 and run.be.net is net     # references decoded once for the whole scene
