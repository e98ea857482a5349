"""Host logic of the multi-GPU schedules (crossscore_b200/scene.py) on CPU: world_size-2 gloo runs.

The CUDA engine is replaced by a stand-in built from the CPU oracle (fp64) that exposes the same methods,
so partitioning, cache exchange (broadcast) and the split-KV exchange (all-gather + LSE merge) are checked
end to end against the single-process oracle forward.  The GPU equivalents live in tests/test_scene_gpu.py.
"""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crossscore_b200.scene import SceneScorer, SplitKVScorer, shard_range
from crossscore_b200.synthetic import make_inputs, make_state_dict
from oracle import crossscore_oracle as O

C, HEADS, D = 384, 8, 48
DT = torch.float64


class OracleEngine:
    """Same surface as crossscore_b200.engine.Engine, arithmetic from oracle/crossscore_oracle.py (CPU, fp64)."""
    adtype = DT
    kv_width = 4 * C

    def __init__(self, sd):
        self.sd = sd

    def features(self, query_img, ref_imgs, st, want_mem=True):
        groups = [g for g in (query_img, None if ref_imgs is None else ref_imgs.reshape(-1, *ref_imgs.shape[-3:]))
                  if g is not None]
        H, W = groups[0].shape[-2:]
        pe = O.multiview_pe_table(self.sd, H, W, DT)
        f = O.dinov2_features(self.sd, torch.cat(groups, 0), DT)[:, 1:] + pe[None]
        nq = 0 if query_img is None else query_img.shape[0]
        xq = f[:nq].reshape(-1, C) if nq else None
        mem = f[nq:].reshape(-1, C) if ref_imgs is not None else None
        return xq, mem

    def project_kv(self, mem, st, out=None):
        cols = []
        for l in range(2):
            w = self.sd[f"ref_cross.attn.layers.{l}.multihead_attn.in_proj_weight"].to(DT)
            b = self.sd[f"ref_cross.attn.layers.{l}.multihead_attn.in_proj_bias"].to(DT)
            cols += [O.linear(mem, w[C:2 * C], b[C:2 * C]), O.linear(mem, w[2 * C:], b[2 * C:])]
        kv = torch.cat(cols, 1)
        if out is not None:
            out.copy_(kv)
            return out
        return kv

    @staticmethod
    def _attend(q, k, v):
        """q (B,P,C), k/v (B,M,C) -> normalised O (B,P,C), LSE (B,8,P)."""
        B, P, _ = q.shape
        M = k.shape[1]
        qh = q.view(B, P, HEADS, D).transpose(1, 2)
        kh = k.view(B, M, HEADS, D).transpose(1, 2)
        vh = v.view(B, M, HEADS, D).transpose(1, 2)
        s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
        lse = torch.logsumexp(s, -1)
        o = torch.exp(s - lse[..., None]) @ vh
        return o.transpose(1, 2).reshape(B, P, C), lse

    def decode(self, xq, kv, B, P, M, ph, pw, st, kv_shared=False, need_attn_weights=False, head_id=0,
               cross_attn_fn=None):
        x = xq.view(B, P, C)
        for l in range(2):
            g = lambda k: self.sd[f"ref_cross.attn.layers.{l}." + k].to(DT)
            sa, _ = O.mha(x, x, x, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"),
                          g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"), HEADS)
            x = O.layer_norm(x + sa, g("norm1.weight"), g("norm1.bias"), 1e-5)
            qc = O.linear(x, g("multihead_attn.in_proj_weight")[:C], g("multihead_attn.in_proj_bias")[:C])
            att = torch.empty(B * P, C, dtype=DT)
            if cross_attn_fn is not None:
                cross_attn_fn(l, qc.reshape(B * P, C), att, None, st)
            else:
                kvb = kv.view(1, M, 4 * C).expand(B, -1, -1) if kv_shared else kv.view(B, M, 4 * C)
                o, _ = self._attend(qc, kvb[..., l * 2 * C:l * 2 * C + C], kvb[..., l * 2 * C + C:(l + 1) * 2 * C])
                att.copy_(o.reshape(B * P, C))
            ca = O.linear(att.view(B, P, C), g("multihead_attn.out_proj.weight"), g("multihead_attn.out_proj.bias"))
            x = O.layer_norm(x + ca, g("norm2.weight"), g("norm2.bias"), 1e-5)
            ff = O.linear(torch.relu(O.linear(x, g("linear1.weight"), g("linear1.bias"))), g("linear2.weight"),
                          g("linear2.bias"))
            x = O.layer_norm(x + ff, g("norm3.weight"), g("norm3.bias"), 1e-5)
        return O.head_and_jigsaw(self.sd, x, ph, pw, dt=DT), None

    def cross_attn_partial(self, layer, qc, kv_local, B, P, M_local, packed, st):
        kvb = kv_local.view(B, M_local, 4 * C)
        o, lse = self._attend(qc.view(B, P, C), kvb[..., layer * 2 * C:layer * 2 * C + C],
                              kvb[..., layer * 2 * C + C:(layer + 1) * 2 * C])
        packed[:B * P * C] = o.reshape(-1).float()
        packed[B * P * C:] = lse.reshape(-1).float()

    def merge_partials(self, gathered, n_parts, B, P, att, lse_out, st):
        part = B * P * C + B * HEADS * P
        g = gathered.view(n_parts, part).double()
        o_parts = [g[r, :B * P * C].view(B, P, HEADS, D).transpose(1, 2) for r in range(n_parts)]
        l_parts = [g[r, B * P * C:].view(B, HEADS, P) for r in range(n_parts)]
        lse = torch.logsumexp(torch.stack(l_parts, 0), 0)
        o = sum(torch.exp(l - lse)[..., None] * o for o, l in zip(o_parts, l_parts))  # -inf part: weight 0
        att.copy_(o.transpose(1, 2).reshape(B * P, C))


# ---------------------------------------------------------------------------------------------------
def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 8, 64, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert shard_range(5, 8, 6) == (5, 5)  # cfg 3 with 5 refs on 8 ranks: ranks 5..7 own no view
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


H, W = 28, 42  # 2 x 3 patches


def _problem(n_ref, n_query):
    sd = make_state_dict(1)
    q, _ = make_inputs(n_query, 1, H, W, seed=3)
    _, r = make_inputs(1, n_ref, H, W, seed=4)
    return sd, q, r[0]


def _worker(rank, world, port, mode, n_ref, n_query, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd, q, refs = _problem(n_ref, n_query)
        eng = OracleEngine(sd)
        if mode == "scene":
            sc = SceneScorer(eng, "cpu")
            out = sc.score_scene(q, refs, batch=2)
            lo, hi = shard_range(n_query, world, rank)
            ret[rank] = (lo, hi, out.clone(), sc.cache_bytes_received)
        else:
            sk = SplitKVScorer(eng, "cpu")
            out = sk.forward(q, refs[None].expand(n_query, -1, -1, -1, -1).contiguous())
            ret[rank] = (0, n_query, out.clone(), sk.allgather_bytes)
    finally:
        dist.destroy_process_group()


def _run(mode, n_ref, n_query, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, n_ref, n_query, ret), nprocs=world, join=True)
    return dict(ret)


def _want(n_ref, n_query):
    sd, q, refs = _problem(n_ref, n_query)
    return O.crossscore_forward(sd, q, refs[None].expand(n_query, -1, -1, -1, -1), dt=DT)["score_map_ref_cross"]


@pytest.mark.parametrize("n_ref,n_query", [(3, 5), (1, 2)])
def test_scene_cache_two_ranks_gloo(n_ref, n_query):
    """cfg 3 shape: reference views sharded 2+1 (or 1+0), K/V cache broadcast, queries sharded 3+2."""
    want = _want(n_ref, n_query)
    ret = _run("scene", n_ref, n_query)
    got = torch.cat([ret[r][2] for r in range(2)], 0)
    assert [ret[r][:2] for r in range(2)] == [shard_range(n_query, 2, r) for r in range(2)]
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 1e-9
    P = (H // 14) * (W // 14)
    a, b = shard_range(n_ref, 2, 0)
    assert ret[1][3] == (b - a) * P * 4 * C * 8  # rank 1 received exactly rank 0's rows (fp64 stand-in)


@pytest.mark.parametrize("n_ref,n_query", [(3, 1), (1, 2)])
def test_split_kv_two_ranks_gloo(n_ref, n_query):
    """cfg 4 shape: keys sharded by reference view; n_ref=1 leaves rank 1 with no keys (LSE = -inf part)."""
    want = _want(n_ref, n_query)
    ret = _run("split", n_ref, n_query)
    for r in range(2):
        assert (ret[r][2] - want).abs().max().item() < 2e-6  # partials cross the wire as fp32
    P = (H // 14) * (W // 14)
    assert ret[0][3] == 2 * 2 * (n_query * P * C + n_query * HEADS * P) * 4  # 2 layers x 2 parts


def test_single_process_paths_equal_oracle():
    """world = 1 (no process group): both schedulers reduce to the plain forward."""
    sd, q, refs = _problem(2, 2)
    eng = OracleEngine(sd)
    want = _want(2, 2)
    got = SceneScorer(eng, "cpu").score_scene(q, refs)
    assert (got - want).abs().max().item() < 1e-9
    got = SplitKVScorer(eng, "cpu").forward(q, refs[None].expand(2, -1, -1, -1, -1).contiguous())
    assert (got - want).abs().max().item() < 2e-6
