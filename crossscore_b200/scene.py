"""Multi-GPU scheduling of the CrossScore hot path: one process per GPU.

The reference scales predict with Lightning DDP + a DistributedSampler (task/predict.py:118-124): queries
shard across ranks and there is NO collective on the data path.  Two exchange steps exist once the work of
a scene is organised around its reference views, and both go through NCCL (torch.distributed):

``SceneScorer``  (BASELINE cfg 3)  many query frames share ONE reference set.  The reference tokens are
    layer-invariant inputs of both decoder layers (model/customised_transformer/transformer.py:251-261 passes
    the same ``memory`` to every layer), so their K/V projections are a per-scene cache:
    rank r encodes its contiguous slice of the reference views, projects K/V of both layers, and the slices
    are exchanged once (one ``broadcast`` per owning rank into the full (N*P, 4E) cache: 21 MB bf16 at 5
    refs).  Queries then shard across ranks with no further communication.

``SplitKVScorer``  (BASELINE cfg 4)  few queries, many reference views.  Rank r owns the K/V of its slice
    of the reference views; the (small) query stream is replicated.  Per decoder layer every rank runs flash
    cross-attention over its local keys and emits a normalised partial (O_r fp32, LSE_r); ONE
    ``all_gather_into_tensor`` of the packed (O_r | LSE_r) buffer per layer and the merge kernel
    (xs_lse_merge: O = sum_r exp(LSE_r - LSE) O_r) reproduce the single-GPU softmax exactly.

Everything numerical is delegated to an engine object (crossscore_b200.engine.Engine: CUDA kernels only);
this file is host logic -- partitioning, buffer ownership and the collectives -- and is exercised on CPU
with the gloo backend in tests/test_scene_cpu.py.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import os

import torch
import torch.distributed as dist

PATCH = 14
C = 384
DEC_HEADS = 8
DEC_LAYERS = 2


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of n items: the first n % world ranks own one extra item.
    (cfg 3 = contiguous blocks of queries per rank, cfg 4 = contiguous reference views.)"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _dist_info(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _serialised(engine):
    """Engine.serialise() (scratch buffers are per engine: calls on different streams are ordered); engines without it
    (the CPU fakes of the host-logic tests) need nothing."""
    import contextlib
    fn = getattr(engine, "serialise", None)
    return fn() if fn is not None else contextlib.nullcontext()


def _stream(device):
    if torch.device(device).type == "cuda":
        return torch.cuda.current_stream(device).cuda_stream
    return None


class SceneScorer:
    """Scores the query frames of one scene against a shared reference set (cfg 3).

    ``engine`` needs: ``features(query_img|None, ref_imgs|None, st) -> (xq32, mem)``,
    ``project_kv(mem, st, out=...)``, ``decode(xq32, kv, B, P, M, ph, pw, st, kv_shared=True) -> (score, _)``,
    ``kv_width`` (columns of the K/V cache) and ``adtype`` (its dtype).
    """

    def __init__(self, engine, device, group=None):
        self.engine = engine
        self.device = torch.device(device)
        self.group = group
        self.world, self.rank = _dist_info(group)
        self.kv: Optional[torch.Tensor] = None
        self.grid: Optional[Tuple[int, int]] = None
        self.n_ref = 0
        self.cache_bytes_received = 0
        self._graphs = {}  # query batch shape -> (CUDAGraph, static input, static output)

    def build_reference_cache(self, ref_imgs: torch.Tensor) -> torch.Tensor:
        """ref_imgs (N,3,H,W) fp32, identical on every rank.  Encodes this rank's slice of the views, projects
        K/V for both decoder layers and exchanges the slices; returns the full (N*P, kv_width) cache."""
        if ref_imgs.dim() != 4 or ref_imgs.shape[1] != 3:
            raise ValueError(f"ref_imgs must be (N,3,H,W), got {tuple(ref_imgs.shape)}")
        N, _, H, W = ref_imgs.shape
        ph, pw = H // PATCH, W // PATCH
        P = ph * pw
        eng, st = self.engine, _stream(self.device)
        kv = torch.empty(N * P, eng.kv_width, device=self.device, dtype=getattr(eng, "kv_dtype", eng.adtype))
        lo, hi = shard_range(N, self.world, self.rank)
        if hi > lo:
            with _serialised(eng):
                _, mem = eng.features(None, ref_imgs[lo:hi].contiguous(), st)
                eng.project_kv(mem, st, out=kv[lo * P:hi * P])
        self.cache_bytes_received = 0
        if self.world > 1:
            for owner in range(self.world):
                a, b = shard_range(N, self.world, owner)
                if b > a:
                    rows = kv[a * P:b * P]
                    dist.broadcast(rows, src=dist.get_global_rank(self.group, owner) if self.group else owner,
                                   group=self.group)
                    if owner != self.rank:
                        self.cache_bytes_received += rows.numel() * rows.element_size()
        self.kv, self.grid, self.n_ref = kv, (ph, pw), N
        self._graphs.clear()  # captured graphs hold the previous cache's pointers
        return kv

    def score(self, query_imgs: torch.Tensor, graph: bool = False) -> torch.Tensor:
        """query_imgs (Bq,3,H,W): THIS rank's queries -> (Bq, 14ph, 14pw) fp32 score maps.

        graph=True replays the batch as ONE CUDA graph (captured per batch shape against the current cache): a batch of
        32 queries is ~115 launches in ~7 ms, so with eager launches the schedule is at the mercy of the host -- eight
        ranks driving their GPUs from one shared CPU lose 15 % to it (bench.py cfg3).  The returned tensor is the graph's
        static output: valid until the next graphed call with the same shape."""
        if self.kv is None:
            raise RuntimeError("build_reference_cache() must run before score()")
        Bq, _, H, W = query_imgs.shape
        ph, pw = H // PATCH, W // PATCH
        if (ph, pw) != self.grid:
            raise ValueError(f"query patch grid {(ph, pw)} differs from the cached reference grid {self.grid}")
        if graph and self.device.type == "cuda":
            return self._score_graphed(query_imgs)
        return self._score_eager(query_imgs.contiguous())

    def _score_eager(self, query_imgs: torch.Tensor) -> torch.Tensor:
        Bq, _, H, W = query_imgs.shape
        ph, pw = H // PATCH, W // PATCH
        P = ph * pw
        eng, st = self.engine, _stream(self.device)
        with _serialised(eng):
            xq32, _ = eng.features(query_imgs, None, st)
            score, _ = eng.decode(xq32, self.kv, Bq, P, self.n_ref * P, ph, pw, st, kv_shared=True)
        return score

    def _score_graphed(self, query_imgs: torch.Tensor) -> torch.Tensor:
        key = tuple(query_imgs.shape)
        ent = self._graphs.get(key)
        if ent is None:
            q_static = torch.empty_like(query_imgs, device=self.device).contiguous()
            q_static.copy_(query_imgs)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up off the capture: workspaces, tables, lazily created buffers
                self._score_eager(q_static)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._score_eager(q_static)
            ent = (g, q_static, out)
            self._graphs[key] = ent
        g, q_static, out = ent
        q_static.copy_(query_imgs, non_blocking=True)
        g.replay()
        return out

    def score_scene(self, all_query_imgs: torch.Tensor, ref_imgs: torch.Tensor, batch: int = 32,
                    graph: bool = False) -> torch.Tensor:
        """Convenience driver: every rank passes the same (Q,3,H,W) queries, scores its contiguous shard in
        batches and returns the shard's maps (rank r owns queries shard_range(Q, world, r))."""
        self.build_reference_cache(ref_imgs)
        lo, hi = shard_range(all_query_imgs.shape[0], self.world, self.rank)
        outs: List[torch.Tensor] = []
        for s in range(lo, hi, batch):
            outs.append(self.score(all_query_imgs[s:min(hi, s + batch)], graph=graph).clone())
        if not outs:
            H, W = all_query_imgs.shape[-2:]
            return torch.empty(0, PATCH * (H // PATCH), PATCH * (W // PATCH), device=self.device)
        return torch.cat(outs, 0)


class SplitKVScorer:
    """Scores queries against MANY reference views with the reference tokens sharded across ranks (cfg 4).

    Extra engine methods used: ``cross_attn_partial(layer, qc, kv_local, B, P, M_local, packed, st)`` writes
    the normalised partial O (B*P*C floats) followed by LSE (B*8*P floats) into ``packed``;
    ``merge_partials(gathered, n_parts, B, P, att, lse_out, st)`` merges ``n_parts`` packed parts.
    """

    def __init__(self, engine, device, group=None, exchange: Optional[str] = None):
        """exchange: how the per-rank partials meet.  "p2p" (default with more than one rank): each rank writes its
        packed (O_r | LSE_r) into a symmetric-memory buffer and the merge kernel PULLS all partials through NVLink peer
        pointers after one cross-rank barrier -- compute and collective in one kernel (xs_lse_merge_peers).
        "nccl": one all_gather_into_tensor per decoder layer, then xs_lse_merge (also the automatic choice when
        symmetric memory cannot be set up, e.g. no peer access)."""
        self.engine = engine
        self.device = torch.device(device)
        self.group = group
        self.world, self.rank = _dist_info(group)
        self.allgather_bytes = 0
        self.exchange = exchange or os.environ.get("XS_SPLITKV_EXCHANGE", "p2p")
        if self.exchange not in ("p2p", "nccl"):
            raise ValueError(f"exchange must be 'p2p' or 'nccl', got {self.exchange!r}")
        if self.world == 1 or self.device.type != "cuda":
            self.exchange = "nccl"  # nothing to exchange / host-side tests of the schedule (gloo)
        self._symm = None       # (part elements, tensor, handle)
        self.peer_bytes_pulled = 0

    def _symm_buffers(self, part: int):
        """Two symmetric packed buffers (one per decoder layer) of `part` floats each, rendezvoused across the ranks."""
        if self._symm is None or self._symm[0] != part:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(DEC_LAYERS * part, dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            self._symm = (part, t, hdl)
        return self._symm[1], self._symm[2]

    def forward(self, query_img: torch.Tensor, ref_imgs: torch.Tensor) -> torch.Tensor:
        """query_img (B,3,H,W), ref_imgs (B,N,3,H,W) identical on every rank -> (B,14ph,14pw) score maps,
        identical on every rank (the decoder stream is replicated; only cross-attention keys are sharded)."""
        B, _, H, W = query_img.shape
        N = ref_imgs.shape[1]
        ph, pw = H // PATCH, W // PATCH
        P = ph * pw
        eng, st = self.engine, _stream(self.device)
        lo, hi = shard_range(N, self.world, self.rank)
        n_loc = hi - lo
        M_loc = n_loc * P
        refs_loc = ref_imgs[:, lo:hi].contiguous() if n_loc else None
        xq32, mem = eng.features(query_img.contiguous(), refs_loc, st)
        kv_loc = eng.project_kv(mem, st) if n_loc else None
        part = B * P * C + B * DEC_HEADS * P
        self.allgather_bytes = 0
        self.peer_bytes_pulled = 0
        symm_t = hdl = None
        if self.exchange == "p2p":
            try:
                symm_t, hdl = self._symm_buffers(part)
            except Exception as e:  # no peer access / symmetric memory backend: use the NCCL collective
                import warnings
                warnings.warn(f"split-KV: symmetric memory unavailable ({e}); exchanging partials with NCCL all-gather")
                self.exchange = "nccl"
        if self.exchange == "nccl":
            packed_nccl = torch.empty(part, device=self.device, dtype=torch.float32)
            gathered = torch.empty(self.world * part, device=self.device, dtype=torch.float32) if self.world > 1 else None

        def cross_attn(layer, qc, att, lse_out, st_):
            packed = symm_t[layer * part:(layer + 1) * part] if hdl is not None else packed_nccl
            if n_loc:
                eng.cross_attn_partial(layer, qc, kv_loc, B, P, M_loc, packed, st_)
            else:  # this rank owns no reference view: neutral element of the merge
                packed[:B * P * C].zero_()
                packed[B * P * C:].fill_(float("-inf"))
            if hdl is not None:
                # every rank's partial of this layer is in place after the barrier; the buffer of this layer is not
                # rewritten before the next query's barrier of the OTHER layer has been passed by all ranks
                hdl.barrier(channel=layer)
                eng.merge_partials_peers(hdl.buffer_ptrs_dev, layer * part, self.world, B, P, att, lse_out, st_)
                self.peer_bytes_pulled += (self.world - 1) * part * 4
            elif self.world > 1:
                dist.all_gather_into_tensor(gathered, packed, group=self.group)
                self.allgather_bytes += gathered.numel() * 4
                eng.merge_partials(gathered, self.world, B, P, att, lse_out, st_)
            else:
                eng.merge_partials(packed, 1, B, P, att, lse_out, st_)

        score, _ = eng.decode(xq32, None, B, P, N * P, ph, pw, st, cross_attn_fn=cross_attn)
        return score
