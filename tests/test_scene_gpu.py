"""GPU tests of the multi-GPU schedules (crossscore_b200/scene.py): shared reference K/V cache (cfg 3) and
split-KV cross-attention with LSE merge (cfg 4).  Single-GPU variants always run; the 2-rank NCCL variants
run when the box has >= 2 GPUs (`gpurun --gpus 2`)."""
import math
import os
import socket

import pytest
import torch

from crossscore_b200 import CrossScoreNet, default_cfg
from crossscore_b200._lib import DT_F32, call
from crossscore_b200.scene import SceneScorer, SplitKVScorer, shard_range
from crossscore_b200.synthetic import make_inputs, make_state_dict
from oracle import crossscore_oracle as O

pytestmark = pytest.mark.gpu
H, W = 84, 112  # 6 x 8 patches
TOL = {"bf16": (1e-2, 1e-3), "fp32": (1e-4, 1e-4)}


def _problem(n_ref, n_query):
    sd = make_state_dict(1)
    q, _ = make_inputs(n_query, 1, H, W, seed=3)
    _, r = make_inputs(1, n_ref, H, W, seed=4)
    return sd, q, r[0]


def _want(sd, q, refs):
    n = q.shape[0]
    return O.crossscore_forward(sd, q, refs[None].expand(n, -1, -1, -1, -1), dt=torch.float64)["score_map_ref_cross"]


def _net(sd, precision, dev):
    net = CrossScoreNet(default_cfg(), precision=precision)
    net.load_state_dict(sd)
    return net.to(dev).eval()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_scene_cache_single_gpu(precision):
    sd, q, refs = _problem(3, 5)
    want = _want(sd, q, refs)
    dev = torch.device("cuda", 0)
    eng = _net(sd, precision, dev)._engine(dev)
    got = SceneScorer(eng, dev).score_scene(q.to(dev), refs.to(dev), batch=2)
    torch.cuda.synchronize()
    err = (got.cpu().double() - want).abs()
    assert err.max().item() <= TOL[precision][0] and err.mean().item() <= TOL[precision][1]


def test_scene_scorer_graph_replay_matches_eager():
    """score(graph=True) replays the batch as one CUDA graph: same bits as the eager launches, for new inputs of the same
    shape, for a second batch shape, and after the reference cache has been rebuilt (graphs are dropped with it)."""
    sd, q, refs = _problem(3, 6)
    dev = torch.device("cuda", 0)
    eng = _net(sd, "bf16", dev)._engine(dev)
    sc = SceneScorer(eng, dev)
    sc.build_reference_cache(refs.to(dev))
    q = q.to(dev)
    for batch in (q[:4], q[2:6], q[:2], q[:4]):
        want = sc.score(batch).clone()
        got = sc.score(batch, graph=True).clone()
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    _, refs2 = make_inputs(1, 3, H, W, seed=44)
    sc.build_reference_cache(refs2[0].to(dev))
    want = sc.score(q[:4]).clone()
    got = sc.score(q[:4], graph=True).clone()
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_split_kv_single_gpu(precision):
    sd, q, refs = _problem(4, 2)
    want = _want(sd, q, refs)
    dev = torch.device("cuda", 0)
    eng = _net(sd, precision, dev)._engine(dev)
    got = SplitKVScorer(eng, dev).forward(q.to(dev), refs[None].expand(2, -1, -1, -1, -1).contiguous().to(dev))
    torch.cuda.synchronize()
    err = (got.cpu().double() - want).abs()
    assert err.max().item() <= TOL[precision][0] and err.mean().item() <= TOL[precision][1]


def test_lse_merge_packed_parts_with_empty_part():
    """xs_lse_merge over a packed (O_r | LSE_r) all-gather buffer; one part saw no keys (LSE = -inf, O = 0)."""
    torch.manual_seed(0)
    dev = "cuda"
    R, B, P, Hh, d = 3, 2, 50, 8, 48
    Cc = Hh * d
    part = B * P * Cc + B * Hh * P
    o = torch.randn(R, B, P, Hh, d, device=dev)
    lse = torch.randn(R, B, Hh, P, device=dev) * 3
    o[1] = 0
    lse[1] = float("-inf")
    packed = torch.empty(R, part, device=dev)
    packed[:, :B * P * Cc] = o.reshape(R, -1)
    packed[:, B * P * Cc:] = lse.reshape(R, -1)
    out = torch.empty(B * P, Cc, device=dev)
    lse_out = torch.empty(B, Hh, P, device=dev)
    call("xs_lse_merge", packed.data_ptr(), packed.data_ptr() + B * P * Cc * 4, out.data_ptr(), lse_out.data_ptr(),
         R, B, P, Hh, d, part, part, DT_F32, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref_lse = torch.logsumexp(lse.double(), 0)
    w = torch.exp(lse.double() - ref_lse)  # (R,B,H,P)
    ref = (w.permute(0, 1, 3, 2)[..., None] * o.double()).sum(0).reshape(B * P, Cc)
    assert torch.isfinite(out).all()
    assert (out.double() - ref).abs().max().item() < 1e-5
    assert (lse_out.double() - ref_lse).abs().max().item() < 1e-5


# ---- 2 ranks over NCCL ------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, precision, n_ref, n_query, ret, exchange="p2p"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sd, q, refs = _problem(n_ref, n_query)
        eng = _net(sd, precision, dev)._engine(dev)
        if mode == "scene":
            sc = SceneScorer(eng, dev)
            out = sc.score_scene(q.to(dev), refs.to(dev), batch=2)
        else:
            sk = SplitKVScorer(eng, dev, exchange=exchange)
            rr = refs[None].expand(n_query, -1, -1, -1, -1).contiguous().to(dev)
            out = sk.forward(q.to(dev), rr).clone()
            out2 = sk.forward(q.to(dev), rr)  # second query: the symmetric buffers are reused
            assert torch.equal(out, out2)
            ret[f"exchange{rank}"] = sk.exchange
        torch.cuda.synchronize()
        ret[rank] = out.cpu()
    finally:
        dist.destroy_process_group()


def _run(mode, precision, n_ref, n_query, exchange="p2p"):
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), mode, precision, n_ref, n_query, ret, exchange), nprocs=2, join=True)
    return dict(ret)


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")


@needs2
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_scene_cache_two_gpus_nccl(precision):
    sd, q, refs = _problem(3, 5)
    want = _want(sd, q, refs)
    ret = _run("scene", precision, 3, 5)
    got = torch.cat([ret[0], ret[1]], 0)
    assert ret[0].shape[0] == shard_range(5, 2, 0)[1]
    err = (got.double() - want).abs()
    assert err.max().item() <= TOL[precision][0] and err.mean().item() <= TOL[precision][1]


@needs2
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("precision,n_ref", [("fp32", 3), ("bf16", 3), ("bf16", 1)])
def test_split_kv_two_gpus(precision, n_ref, exchange):
    """exchange = p2p: partials pulled through NVLink peer pointers inside the merge kernel (symmetric memory);
    nccl: all-gather + merge.  Both must give the single-GPU answer."""
    sd, q, refs = _problem(n_ref, 2)
    want = _want(sd, q, refs)
    ret = _run("split", precision, n_ref, 2, exchange)
    assert ret["exchange0"] == ret["exchange1"] == exchange  # p2p must really run as p2p on an NVLink box
    assert torch.equal(ret[0], ret[1])  # the replicated decoder stream ends identical on both ranks
    err = (ret[0].double() - want).abs()
    assert err.max().item() <= TOL[precision][0] and err.mean().item() <= TOL[precision][1]
