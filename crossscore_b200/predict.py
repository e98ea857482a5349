"""Lightning-free predict runner: the caller of the hot path.

Mirrors what `task/predict.py` + `CrossScoreLightningModule.predict_step / on_predict_batch_end` do for one scene
(a directory of query images scored against a directory of reference images), without Lightning / Hydra:

  * scene listing and reference sampling: dataloading/dataset/simple_reference.py:44-60 (sorted listdir),
    utils/neighbour/sampler.py:19-35 (N random references per query, `empty_image` padding, `deterministic`
    = the first N), config/data/SimpleReference.yaml:16-19;
  * per-image work: uint8 decode on the host, then ON THE DEVICE /255 -> resize short side (antialiased bilinear) ->
    ImageNet normalise (crossscore_b200.imgproc, replacing the dataloader workers), CrossScoreNet.forward
    (task/core.py:266-272 call signature), per-frame mean / uint16 / turbo maps (utils/io/score_summariser.py:180-181,
    utils/io/batch_writer.py:114-135,263-270) before the device -> host copy;
  * outputs: `<out>/batch/score_map_ref_cross/r{rank}_B{batch:04}_b{b:03}_{name}.png` (batch_writer.py:114-135 naming)
    and `<out>/score_summary/scores.csv` with the columns of score_summariser.py:150-154 (`%.4f`);
  * out-dir naming of task/predict.py:46-66 (`log/<now>/predict_empty_ckpt/<now>[_alias]`, or next to the ckpt);
  * multi-GPU: one process per GPU, queries strided over ranks like a DistributedSampler without shuffling
    (task/predict.py:119-124); file names carry the rank.

Scene scheduler (row 4): decoded + preprocessed reference images are cached on the device by path (the reference
re-reads and re-normalises them for every query); with `deterministic` references every query of the scene sees the
SAME reference set, so their features and decoder K/V are computed once (crossscore_b200.scene.SceneScorer) and
only the queries run through the backbone.

The device work sits behind a small backend object so the host logic is testable without a GPU; the shipped
backend (`DeviceBackend`) has no CPU fallback.
"""
from __future__ import annotations

import argparse
import csv
import os
from collections import OrderedDict
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

EMPTY = "empty_image"  # utils/neighbour/sampler.py:24


# ------------------------------------------------------------------------------------------------------------------
# scene listing / reference selection (pure host logic)
# ------------------------------------------------------------------------------------------------------------------
def list_scene(query_dir: str, reference_dir: str):
    """simple_reference.py:52-58: sorted directory listings, `~` expanded."""
    query_dir, reference_dir = os.path.expanduser(query_dir), os.path.expanduser(reference_dir)
    q = [os.path.join(query_dir, p) for p in sorted(os.listdir(query_dir))]
    r = [os.path.join(reference_dir, p) for p in sorted(os.listdir(reference_dir))]
    return q, r


def select_references(ref_list: Sequence[str], n_sample: int, deterministic: bool, rng=np.random) -> List[str]:
    """utils/neighbour/sampler.py:19-35 (SamplerRandom.sample), including its use of numpy's global RNG
    (pass `rng` = a seeded np.random.RandomState to reproduce `lightning.seed_everything(seed)` streams)."""
    ref_list = list(ref_list)
    if n_sample > len(ref_list):
        result = ref_list + [EMPTY] * (n_sample - len(ref_list))
        return rng.permutation(result).tolist()
    if deterministic:
        return ref_list[:n_sample]
    return rng.choice(ref_list, n_sample, replace=False).tolist()


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Indices of the queries rank `rank` scores (stride `world`, unshuffled, no duplicate padding)."""
    return list(range(rank, n, world))


def predict_out_dir(ckpt_path: Optional[str], out_dir: Optional[str] = None, alias: str = "", now: Optional[str] = None):
    """task/predict.py:46-66."""
    now = now or datetime.now().strftime("%Y%m%d_%H%M%S.%f")
    if ckpt_path is None:
        log_dir = Path("log") / now / "predict_empty_ckpt"
    else:
        log_dir = Path(ckpt_path).parents[1] / "predict"
    out = out_dir if out_dir is not None else f"{log_dir}/{now}"
    if alias != "":
        out += f"_{alias}"
    return out


def score_map_file_name(query_path: str, rank: int, batch_idx: int, b: int) -> str:
    """utils/io/batch_writer.py:117-131."""
    stem = str(Path(*Path(query_path).parts[-5:])).replace("/", "_").replace(".png", "")
    return f"r{rank}_B{batch_idx:04}_b{b:03}_{stem}.png"


def metric_type_str(metric_type: str, metric_min: int) -> str:
    """utils/io/score_summariser.py:157-165."""
    if metric_type == "ssim":
        return f"{metric_type}_-1_1" if metric_min == -1 else f"{metric_type}_0_1"
    return f"{metric_type}"


def intrinsic_vrange(metric_type: str):
    """utils/io/batch_writer.py:9-22 (get_vrange): the uint16 maps always use the metric's intrinsic range."""
    if metric_type == "ssim":
        return [-1, 1]
    if metric_type in ("mse", "mae"):
        return [0, 1]
    raise ValueError(f"metric_type {metric_type} not supported")


def summary_row(query_path: str, score: float):
    """utils/io/score_summariser.py:183-193 (falls back to shorter paths than the reference's dataset layout)."""
    parts = query_path.split("/")
    scene = parts[-5] if len(parts) >= 5 else (parts[-2] if len(parts) >= 2 else "")
    rendered_dir = os.path.join(*parts[:-2]) if len(parts) > 2 else ""
    image_name = parts[-1].replace("frame_", "")
    return [scene, rendered_dir, image_name, score]


class ByteLRU:
    """Device cache of preprocessed reference images keyed by path, bounded in bytes."""

    def __init__(self, capacity_bytes: int):
        self.capacity, self.used = int(capacity_bytes), 0
        self._d: "OrderedDict[str, object]" = OrderedDict()
        self.hits = self.misses = 0

    def get(self, key):
        if key in self._d:
            self._d.move_to_end(key)
            self.hits += 1
            return self._d[key]
        self.misses += 1
        return None

    def put(self, key, value, nbytes: int):
        if nbytes > self.capacity:
            return
        while self.used + nbytes > self.capacity and self._d:
            _, (old, ob) = self._d.popitem(last=False)
            self.used -= ob
        self._d[key] = (value, nbytes)
        self.used += nbytes

    def value(self, key):
        hit = self.get(key)
        return None if hit is None else hit[0]


# ------------------------------------------------------------------------------------------------------------------
# device backend
# ------------------------------------------------------------------------------------------------------------------
class DeviceBackend:
    """CUDA implementation of the per-batch work (no CPU fallback)."""

    def __init__(self, net, device="cuda:0"):
        import torch
        self.torch = torch
        self.net = net
        self.device = torch.device(device)

    def preprocess(self, images_u8: np.ndarray, resize_short_side: int):
        """(n, H, W, 3) uint8 host array -> (n, 3, H1, W1) fp32 device tensor."""
        from . import imgproc
        t = self.torch.from_numpy(np.ascontiguousarray(images_u8)).pin_memory().to(self.device, non_blocking=True)
        return imgproc.preprocess_u8(t, resize_short_side)

    def stack(self, tensors):
        return self.torch.stack(list(tensors), 0)

    def forward(self, q, r):
        return self.net(q, r, False, 0, False)["score_map_ref_cross"]

    def scene_scorer(self):
        from .scene import SceneScorer
        return SceneScorer(self.net._engine(self.device), self.device)

    def postprocess(self, score, gray_vrange, rgb_vrange):
        from . import imgproc
        out = imgproc.postprocess_scores(score, mean=True, gray16_vrange=gray_vrange, rgb_vrange=rgb_vrange)
        return {k: v.cpu().numpy() for k, v in out.items()}

    def nbytes(self, t):
        return t.numel() * t.element_size()


def read_image_u8(path: str) -> np.ndarray:
    """utils/io/images.py:26-29 (image_read) without the float conversion, which happens on the device."""
    from PIL import Image
    a = np.array(Image.open(path))
    if a.ndim == 2:
        a = np.repeat(a[:, :, None], 3, 2)
    return np.ascontiguousarray(a[:, :, :3]).astype(np.uint8, copy=False)


def write_png(path: Path, arr: np.ndarray):
    from PIL import Image
    Image.fromarray(arr).save(path)  # uint16 (H, W) -> 16-bit gray PNG, uint8 (H, W, 3) -> RGB


# ------------------------------------------------------------------------------------------------------------------
# the runner
# ------------------------------------------------------------------------------------------------------------------
class PredictRunner:
    def __init__(self, backend, out_dir: str, metric_type: str = "ssim", metric_min: int = 0, metric_max: int = 1,
                 batch_size: int = 8, num_refs: int = 5, deterministic_refs: bool = False,
                 resize_short_side: int = 518, colour_mode: str = "rgb", zero_reference: bool = False,
                 rank: int = 0, world: int = 1, seed: Optional[int] = 1, cache_bytes: int = 8 << 30,
                 write_maps: bool = True, reader=read_image_u8):
        if colour_mode not in ("gray", "rgb"):
            raise ValueError(f"colour_mode {colour_mode} not supported")  # batch_writer.py:263-270
        self.be, self.out_dir = backend, Path(out_dir)
        self.metric_type, self.metric_min, self.metric_max = metric_type, metric_min, metric_max
        self.vrange_intrinsic = intrinsic_vrange(metric_type)
        self.batch_size, self.num_refs, self.deterministic = batch_size, num_refs, deterministic_refs
        self.resize, self.colour_mode, self.zero_reference = resize_short_side, colour_mode, zero_reference
        self.rank, self.world = rank, world
        self.rng = np.random.RandomState(seed) if seed is not None else np.random
        self.cache = ByteLRU(cache_bytes)
        self.write_maps, self.reader = write_maps, reader
        self.rows: List[list] = []

    # -- images -----------------------------------------------------------------------------------------------
    def _load(self, path: str, like_shape=None, cache: bool = False):
        """One preprocessed image (3, H1, W1) on the device; `empty_image` / zero_reference -> zeros BEFORE the
        normalisation (nvs_dataset.py:459-468)."""
        if cache:
            hit = self.cache.value(path)
            if hit is not None:
                return hit
        if path == EMPTY or (self.zero_reference and like_shape is not None and cache):
            u8 = np.zeros(like_shape, np.uint8)
        else:
            u8 = self.reader(path)
        t = self.be.preprocess(u8[None], self.resize)[0]
        if cache:
            self.cache.put(path, t, self.be.nbytes(t))
        return t

    # -- scoring ----------------------------------------------------------------------------------------------
    def run(self, query_paths: Sequence[str], ref_paths: Sequence[str]):
        mine = shard_indices(len(query_paths), self.rank, self.world)
        # references are drawn for EVERY query in dataset order (one RNG stream, like a single-process dataloader
        # with num_workers = 0), then this rank keeps its share
        refs_all = [select_references(ref_paths, self.num_refs, self.deterministic, self.rng)
                    for _ in range(len(query_paths))]
        shared = self.deterministic and self.num_refs <= len(ref_paths) and not self.zero_reference
        scorer = None
        map_dir = self.out_dir / "batch" / "score_map_ref_cross"
        if self.write_maps:
            map_dir.mkdir(parents=True, exist_ok=True)
        if shared and len(query_paths) > 0:
            # The whole scene shares one reference set: encode it once.  build_reference_cache() exchanges K/V slices
            # with one collective per owning rank, so EVERY rank must get here -- also a rank whose query shard is
            # empty (more GPUs than frames) -- and before anything that can raise on a single rank.
            refs = self.be.stack([self._load(p, None, cache=True) for p in refs_all[0]])
            scorer = self.be.scene_scorer()
            scorer.build_reference_cache(refs)
            shared_hw = tuple(refs.shape[-2:])
        for bi, lo in enumerate(range(0, len(mine), self.batch_size)):
            idx = mine[lo:lo + self.batch_size]
            q_u8 = [self.reader(query_paths[i]) for i in idx]
            raw_shape = q_u8[0].shape
            if any(a.shape != raw_shape for a in q_u8):
                raise ValueError("query images of one batch must have the same size")
            q = self.be.preprocess(np.stack(q_u8, 0), self.resize)
            if shared:
                if shared_hw != tuple(q.shape[-2:]):
                    raise ValueError("reference and query images must have the same size after resizing")
                score = scorer.score(q).clone()
            else:
                r = self.be.stack([self.be.stack([self._load(p, raw_shape, cache=True) for p in refs_all[i]]) for i in idx])
                if tuple(r.shape[-2:]) != tuple(q.shape[-2:]):
                    raise ValueError("reference and query images must have the same size after resizing")
                score = self.be.forward(q, r)
            want_rgb = self.write_maps and self.colour_mode == "rgb"
            want_gray = self.write_maps and self.colour_mode == "gray"
            out = self.be.postprocess(score, self.vrange_intrinsic if want_gray else None,
                                      (self.metric_min, self.metric_max) if want_rgb else None)
            for b, i in enumerate(idx):
                self.rows.append(summary_row(query_paths[i], float(out["mean"][b])))
                if self.write_maps:
                    name = score_map_file_name(query_paths[i], self.rank, bi, b)
                    write_png(map_dir / name, out["rgb"][b] if want_rgb else out["gray16"][b])
        return self.rows

    def write_summary(self):
        d = self.out_dir / "score_summary"
        d.mkdir(parents=True, exist_ok=True)
        name = "scores.csv" if self.world == 1 else f"scores_r{self.rank}.csv"
        with open(d / name, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["scene_name", "rendered_dir", "image_name",
                        f"pred_{metric_type_str(self.metric_type, self.metric_min)}"])
            for row in sorted(self.rows, key=lambda r: (r[0], r[1], r[2])):
                w.writerow(row[:3] + [f"{row[3]:.4f}"])
        return d / name


def main(argv=None):
    ap = argparse.ArgumentParser(description="CrossScore predict (one scene) on B200")
    ap.add_argument("--query-dir", required=True)
    ap.add_argument("--reference-dir", required=True)
    ap.add_argument("--ckpt", default=None, help="Lightning checkpoint (state_dict with the `model.` prefix)")
    ap.add_argument("--out-dir", default=None)
    ap.add_argument("--alias", default="")
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--num-refs", type=int, default=5)
    ap.add_argument("--deterministic-refs", action="store_true")
    ap.add_argument("--resize-short-side", type=int, default=518)
    ap.add_argument("--colour-mode", default="rgb", choices=["rgb", "gray"])
    ap.add_argument("--zero-reference", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args(argv)

    import torch
    from . import CrossScoreNet, default_cfg
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        from .runner import bind_to_gpu_numa
        bind_to_gpu_numa(local)  # decode threads and pinned buffers on the GPU's own socket
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = default_cfg()
    net = CrossScoreNet(cfg, precision=args.precision)
    if args.ckpt is not None:
        from .model import load_checkpoint
        load_checkpoint(net, args.ckpt)
    else:  # task/predict.py:46-50 "predict_empty_ckpt": random weights, still runs
        from .synthetic import make_state_dict
        net.load_state_dict(make_state_dict(args.seed))
    net = net.to(dev).eval()
    out_dir = predict_out_dir(args.ckpt, args.out_dir, args.alias)
    m = cfg.model.predict.metric
    runner = PredictRunner(DeviceBackend(net, dev), out_dir, m.type, m.min, m.max, args.batch_size, args.num_refs,
                           args.deterministic_refs, args.resize_short_side, args.colour_mode, args.zero_reference,
                           rank, world, args.seed)
    q, r = list_scene(args.query_dir, args.reference_dir)
    with torch.inference_mode():
        runner.run(q, r)
    path = runner.write_summary()
    torch.cuda.synchronize()
    print(f"[rank {rank}] wrote {len(runner.rows)} scores to {path}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
