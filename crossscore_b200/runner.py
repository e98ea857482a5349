"""Host-buffer front end: score batches that live in (pinned) HOST memory.

This is the call a user of the reference makes per batch -- Lightning moves ``batch["query/img"]`` and
``batch["reference/cross/imgs"]`` to the device, runs ``CrossScoreNet.forward`` and the writers read the
score map back (task/core.py:266-272, utils/io/batch_writer.py:133-135).  ``HostScorer`` does the same with
double-buffered device inputs: the H2D copy of batch i+1 runs on a copy stream while batch i computes, and
the score maps are read back to pinned host memory.
"""
from __future__ import annotations

import torch


class HostScorer:
    def __init__(self, net, device="cuda:0", depth: int = 2):
        self.net = net
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs = None
        self._copied = [torch.cuda.Event() for _ in range(depth)]
        self._consumed = [torch.cuda.Event() for _ in range(depth)]
        self._step = 0

    def _ensure(self, q, r):
        key = (tuple(q.shape), tuple(r.shape))
        if self._bufs is None or self._bufs[0] != key:
            dev = self.device
            qs = [torch.empty(q.shape, dtype=torch.float32, device=dev) for _ in range(self.depth)]
            rs = [torch.empty(r.shape, dtype=torch.float32, device=dev) for _ in range(self.depth)]
            B, H, W = q.shape[0], 14 * (q.shape[-2] // 14), 14 * (q.shape[-1] // 14)
            outs = [torch.empty(B, H, W, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
            self._bufs = (key, qs, rs, outs)
            self._step = 0
        return self._bufs[1:]

    def h2d_bytes(self, q, r):
        return q.numel() * 4 + r.numel() * 4

    def submit(self, q_host: torch.Tensor, r_host: torch.Tensor) -> torch.Tensor:
        """Enqueue one batch (pinned fp32 host tensors).  Returns the pinned host tensor that will hold the
        score maps once the current stream has been synchronised (valid until `depth` later submits)."""
        qs, rs, outs = self._ensure(q_host, r_host)
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
        outs[b].copy_(score, non_blocking=True)
        self._step += 1
        return outs[b]
