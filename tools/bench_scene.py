"""Throughput of the two multi-GPU schedules of BASELINE.json (development / evidence tool, run under torchrun or alone):
  cfg 3: a scene of Q query frames sharing ONE set of 5 references: reference K/V cache built once (views sharded over
         the ranks, slices exchanged over NCCL), then every rank scores its Q/world queries in batches of 32;
  cfg 4: 1 query x 64 references, reference tokens sharded over the ranks (split-KV cross-attention, one all-gather of
         packed (O, LSE) per decoder layer) against the same problem on one GPU.
Prints one JSON line on rank 0.  Times are CUDA events on the launching stream, max over ranks."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as dist
from crossscore_b200 import CrossScoreNet, default_cfg
from crossscore_b200.scene import SceneScorer, SplitKVScorer, shard_range
from crossscore_b200.synthetic import make_inputs, make_state_dict

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
Q = int(os.environ.get("Q", 1024))
net = CrossScoreNet(default_cfg(), precision="bf16")
net.load_state_dict(make_state_dict(1))
net = net.to(dev).eval()
eng = net._engine(dev)

def tmax(ms):
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {"n_gpus": world}
with torch.inference_mode():
    # ---------------- cfg 3 ----------------
    if os.environ.get("SKIP_CFG3") != "1":
      _, refs = make_inputs(1, 5, 518, 518, seed=7)
      refs = refs[0].to(dev)
      lo, hi = shard_range(Q, world, rank)
      q_mine, _ = make_inputs(32, 1, 518, 518, seed=100 + rank)   # one resident batch reused for the rank's share
      q_mine = q_mine.to(dev)
      sc = SceneScorer(eng, dev)
      sc.build_reference_cache(refs); sc.score(q_mine)            # warm-up
      barrier(); e0.record()
      sc.build_reference_cache(refs)
      e1.record(); barrier()
      ms_cache = tmax(e0.elapsed_time(e1))
      n_batches = (hi - lo + 31) // 32
      barrier(); e0.record()
      for _ in range(n_batches):
          s = sc.score(q_mine)
      e1.record(); barrier()
      ms_score = tmax(e0.elapsed_time(e1))
      out["cfg3"] = {"queries": Q, "refs": 5, "cache_build_ms": ms_cache, "cache_bytes_received_per_rank": sc.cache_bytes_received,
                     "score_ms": ms_score, "maps_per_s_excl_cache": Q / (ms_score * 1e-3),
                     "maps_per_s_incl_cache": Q / ((ms_score + ms_cache) * 1e-3),
                     "note": "per-map work 135.0 GF (queries only; SURVEY 8d), inputs device-resident"}
    # ---------------- cfg 4 ----------------
    q1, r64 = make_inputs(1, 64, 518, 518, seed=9)
    q1, r64 = q1.to(dev), r64.to(dev)
    out["cfg4"] = {"refs": 64}
    for mode in (("p2p", "nccl") if world > 1 else ("nccl",)):
        sk = SplitKVScorer(eng, dev, exchange=mode)
        got = sk.forward(q1, r64).clone()
        sk.forward(q1, r64)
        barrier(); e0.record()
        for _ in range(10):
            sk.forward(q1, r64)
        e1.record(); barrier()
        out["cfg4"][f"ms_per_query_split_kv_{sk.exchange}"] = tmax(e0.elapsed_time(e1)) / 10
        out["cfg4"][f"exchange_bytes_per_query_{sk.exchange}"] = sk.allgather_bytes or sk.peer_bytes_pulled
    if rank == 0:
        full = net(q1, r64, False, 0, False)["score_map_ref_cross"]
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            net(q1, r64, False, 0, False)
        e1.record(); torch.cuda.synchronize()
        out["cfg4"]["ms_per_query_one_gpu"] = e0.elapsed_time(e1) / 3
        out["cfg4"]["max_abs_diff_vs_one_gpu"] = float((full - got).abs().max())
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
