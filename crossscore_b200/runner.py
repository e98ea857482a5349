"""Host-buffer front end: score batches that live in (pinned) HOST memory.

This is the call a user of the reference makes per batch -- Lightning moves ``batch["query/img"]`` and
``batch["reference/cross/imgs"]`` to the device, runs ``CrossScoreNet.forward`` and the writers read the
score map back (task/core.py:266-272, utils/io/batch_writer.py:133-135).  ``HostScorer`` does the same with
double-buffered device inputs: the H2D copy of batch i+1 runs on a copy stream while batch i computes, and
the score maps of batch i are read back to pinned host memory on a third stream while batch i+1 computes (the
forward's output buffer is copied aside on the device first; a read-back queued on the compute stream would hold
up the next forward for the 34 MB PCIe transfer).
"""
from __future__ import annotations

import os

import torch


def bind_to_gpu_numa(device_index: int) -> int:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (its NUMA node / PCIe root).

    One process per GPU: pinned host buffers are then allocated on (first touched from) the socket the GPU hangs off, so
    the H2D copies of eight ranks do not all cross the inter-socket link.  Opt-in (XS_NUMA_BIND=1): on this pool's
    single-socket VMs it changes nothing for the copies (measured) and the narrower CPU mask is not free for NCCL's helper
    threads.  Returns the number of CPUs bound to; 0 when not enabled, when NVML or the affinity call is unavailable or
    the mask is empty (then nothing changes)."""
    if os.environ.get("XS_NUMA_BIND", "0") != "1" or not hasattr(os, "sched_setaffinity"):
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:  # NVML enumerates the physical devices
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))  # stay inside the cgroup / container mask
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


class _ReadBack:
    """Device -> pinned host copies on their own stream, double-buffered through device staging tensors."""

    def __init__(self, device, depth):
        self.device, self.depth = device, depth
        self.stream = torch.cuda.Stream(device)
        self._staged = [torch.cuda.Event() for _ in range(depth)]
        self._done = [torch.cuda.Event() for _ in range(depth)]
        self._used = [False] * depth
        self._last = None

    def push(self, b, pairs):
        """pairs: [(device result, device staging, pinned host)].  Called on the compute stream right after the forward."""
        main = torch.cuda.current_stream(self.device)
        if self._used[b]:
            main.wait_event(self._done[b])  # the staging tensors' previous read-back (depth submits ago) has finished
        for src, stage, _ in pairs:
            stage.copy_(src, non_blocking=True)
        self._staged[b].record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._staged[b])
            for _, stage, host in pairs:
                host.copy_(stage, non_blocking=True)
            self._done[b].record(self.stream)
        self._used[b] = True
        self._last = self._done[b]

    def fence(self):
        """Make the current stream wait for every read-back queued so far (then synchronising it is enough)."""
        if self._last is not None:
            torch.cuda.current_stream(self.device).wait_event(self._last)

    def synchronize(self):
        if self._last is not None:
            self._last.synchronize()


class HostScorer:
    def __init__(self, net, device="cuda:0", depth: int = 2):
        self.net = net
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        self._rb = _ReadBack(self.device, depth)
        self._bufs = None
        self._copied = [torch.cuda.Event() for _ in range(depth)]
        self._consumed = [torch.cuda.Event() for _ in range(depth)]
        self._step = 0

    def _ensure(self, q, r):
        key = (tuple(q.shape), tuple(r.shape))
        if self._bufs is None or self._bufs[0] != key:
            self.synchronize()
            dev = self.device
            qs = [torch.empty(q.shape, dtype=torch.float32, device=dev) for _ in range(self.depth)]
            rs = [torch.empty(r.shape, dtype=torch.float32, device=dev) for _ in range(self.depth)]
            B, H, W = q.shape[0], 14 * (q.shape[-2] // 14), 14 * (q.shape[-1] // 14)
            stage = [torch.empty(B, H, W, dtype=torch.float32, device=dev) for _ in range(self.depth)]
            outs = [torch.empty(B, H, W, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
            self._bufs = (key, qs, rs, stage, outs)
            self._rb = _ReadBack(self.device, self.depth)
            self._step = 0
        return self._bufs[1:]

    def h2d_bytes(self, q, r):
        return q.numel() * 4 + r.numel() * 4

    def fence(self):
        """The current stream waits for all queued read-backs."""
        self._rb.fence()

    def synchronize(self):
        """Block the host until every returned host tensor is complete."""
        self._rb.synchronize()

    def submit(self, q_host: torch.Tensor, r_host: torch.Tensor) -> torch.Tensor:
        """Enqueue one batch (pinned fp32 host tensors).  Returns the pinned host tensor that will hold the score maps
        after ``synchronize()`` (or ``fence()`` + a synchronise of the current stream); it is reused `depth` submits
        later."""
        qs, rs, stage, outs = self._ensure(q_host, r_host)
        b = self._step % self.depth
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self._step >= self.depth:
                self.copy_stream.wait_event(self._consumed[b])
            qs[b].copy_(q_host, non_blocking=True)
            rs[b].copy_(r_host, non_blocking=True)
            self._copied[b].record(self.copy_stream)
        main.wait_event(self._copied[b])
        score = self.net(qs[b], rs[b], False, 0, False)["score_map_ref_cross"]
        self._consumed[b].record(main)
        self._rb.push(b, [(score, stage[b], outs[b])])
        self._step += 1
        return outs[b]


class HostPipeline:
    """uint8 host images in, frame means + uint16 score maps out: the whole chain of SURVEY.md section 8f around the
    model on the device (decode excluded).  Per batch: H2D of the uint8 pixels (a quarter of the fp32 bytes the
    reference's dataloader hands to Lightning) on a copy stream, xs_preprocess_u8_resize_normalize,
    CrossScoreNet.forward, xs_score_postprocess, D2H of (B,) means and (B,H,W) uint16 maps (half of the fp32 map).
    Double-buffered like HostScorer."""

    def __init__(self, net, device="cuda:0", resize_short_side: int = -1, gray16_vrange=(0, 1), depth: int = 2):
        from . import imgproc
        self.imgproc = imgproc
        self.net, self.device, self.depth = net, torch.device(device), depth
        self.resize, self.vrange = resize_short_side, list(gray16_vrange)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._rb = _ReadBack(self.device, depth)
        self._bufs = None
        self._copied = [torch.cuda.Event() for _ in range(depth)]
        self._consumed = [torch.cuda.Event() for _ in range(depth)]
        self._step = 0

    def fence(self):
        self._rb.fence()

    def synchronize(self):
        self._rb.synchronize()

    def _ensure(self, q, r):
        key = (tuple(q.shape), tuple(r.shape))
        if self._bufs is None or self._bufs[0] != key:
            self.synchronize()
            self._rb = _ReadBack(self.device, self.depth)
            dev = self.device
            qs = [torch.empty(q.shape, dtype=torch.uint8, device=dev) for _ in range(self.depth)]
            rs = [torch.empty(r.shape, dtype=torch.uint8, device=dev) for _ in range(self.depth)]
            B, H0, W0 = q.shape[0], q.shape[1], q.shape[2]
            H1, W1 = self.imgproc.resize_output_size(H0, W0, self.resize)
            H, W = 14 * (H1 // 14), 14 * (W1 // 14)
            means = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
            maps = [torch.empty(B, H, W, dtype=torch.uint16).pin_memory() for _ in range(self.depth)]
            self._stage = [(torch.empty(B, dtype=torch.float32, device=dev),
                            torch.empty(B, H, W, dtype=torch.uint16, device=dev)) for _ in range(self.depth)]
            self._bufs = (key, qs, rs, means, maps)
            self._step = 0
        return self._bufs[1:]

    def h2d_bytes(self, q, r):
        return q.numel() + r.numel()

    def d2h_bytes(self, q):
        _, _, _, means, maps = self._bufs
        return means[0].numel() * 4 + maps[0].numel() * 2

    def submit(self, q_host: torch.Tensor, r_host: torch.Tensor):
        """q_host (B,H,W,3), r_host (B,N,H,W,3) pinned uint8.  Returns (means, maps16) pinned host tensors, valid
        after ``synchronize()`` (or ``fence()`` + a synchronise of the current stream) until `depth` later submits."""
        qs, rs, means, maps = self._ensure(q_host, r_host)
        b = self._step % self.depth
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self._step >= self.depth:
                self.copy_stream.wait_event(self._consumed[b])
            qs[b].copy_(q_host, non_blocking=True)
            rs[b].copy_(r_host, non_blocking=True)
            self._copied[b].record(self.copy_stream)
        main.wait_event(self._copied[b])
        B, N = r_host.shape[0], r_host.shape[1]
        q = self.imgproc.preprocess_u8(qs[b], self.resize)
        r = self.imgproc.preprocess_u8(rs[b].view(B * N, *r_host.shape[2:]), self.resize)
        self._consumed[b].record(main)
        score = self.net(q, r.view(B, N, *r.shape[1:]), False, 0, False)["score_map_ref_cross"]
        out = self.imgproc.postprocess_scores(score, mean=True, gray16_vrange=self.vrange)
        sm, sg = self._stage[b]
        self._rb.push(b, [(out["mean"], sm, means[b]), (out["gray16"], sg, maps[b])])
        self._step += 1
        return means[b], maps[b]


class GraphedScorer:
    """The forward of one fixed input shape captured into a CUDA graph and replayed.

    A single query with its references (BASELINE cfg 1) is launch-bound: ~115 kernel launches of a few microseconds
    each.  Every C-ABI entry only enqueues on the caller's stream and the engine's workspaces are persistent, so the
    whole forward captures as is; replaying it costs one launch.  Inputs are copied into the graph's static buffers
    (device-to-device), the returned score map is the graph's static output (valid until the next call)."""

    def __init__(self, net, device="cuda:0", warmup: int = 2):
        self.net, self.device, self.warmup = net, torch.device(device), warmup
        self._key = None
        self._graph = None

    def _capture(self, q, r):
        self._q = torch.empty_like(q, device=self.device)
        self._r = torch.empty_like(r, device=self.device)
        self._q.copy_(q)
        self._r.copy_(r)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up off the capture: builds workspaces, tables and packed weights
            for _ in range(self.warmup):
                self.net(self._q, self._r, False, 0, False)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = self.net(self._q, self._r, False, 0, False)["score_map_ref_cross"]
        self._key = (tuple(q.shape), tuple(r.shape))

    def __call__(self, q: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
        if self._key != (tuple(q.shape), tuple(r.shape)):
            self._capture(q, r)
        else:
            self._q.copy_(q, non_blocking=True)
            self._r.copy_(r, non_blocking=True)
        self._graph.replay()
        return self._out
