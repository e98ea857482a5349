#!/usr/bin/env python
"""Benchmark of the CrossScore inference hot path (BASELINE.json metric: score maps/s, 518x518, 5 refs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--precision bf16|fp32]

One "step" = one forward of B=32 queries x 5 reference views at 518x518 (BASELINE.json configs[1]) per GPU.
N>1 is launched by torchrun (one rank per GPU): the headline `value` shards queries across ranks with no data-path
collective (weak scaling).  The same line carries, for N>1, the two schedules that DO exchange
data over NCCL / NVLink -- `cfg3` (1024-frame scene, shared reference K/V cache broadcast once) and `cfg4` (1 query x
64 references, split-KV cross-attention merged per decoder layer) -- each with a parity field against the
single-GPU path; and for N=1 `cfg5` (1036x1036 x 16 refs), `fp32_mode` (the parity mode's speed) and `torch_gpu`
(the oracle's torch ops on the same B200 under bf16 autocast: the same-box library comparator).
Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle port of the reference path on the
host cores (the reference is pure Python/PyTorch and is not present on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 518
N_REF = 5
C = 384


def flops_per_map(P=1369, N=N_REF, query_only=False):
    """Algorithmic FLOPs per score map (SURVEY.md section 8d convention).  query_only: the reference views' backbone
    passes and K/V projections are cached per scene (cfg 3), so a map costs one backbone pass + the decoder."""
    T, M = P + 1, N * P
    f_dino = 2 * P * 588 * C + 12 * (2 * T * C * 3 * C + 4 * T * T * C + 2 * T * C * C + 4 * T * C * 4 * C)
    f_dec = 2 * (2 * P * C * 3 * C + 4 * P * P * C + 2 * P * C * C + 2 * P * C * C + 2 * M * C * 2 * C
                 + 4 * P * M * C + 2 * P * C * C + 4 * P * C * C)
    f_head = 2 * P * C * C + 2 * P * C * 196
    if query_only:
        return f_dino + (f_dec - 2 * 2 * M * C * 2 * C) + f_head
    return (1 + N) * f_dino + f_dec + f_head


def attn_flops_per_map(P=1369, N=N_REF):
    T, M = P + 1, N * P
    return (1 + N) * 12 * 4 * T * T * C + 2 * (4 * P * P * C + 4 * P * M * C)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def ncu_traffic(kernel_tag):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture of this same command (profiles/traffic.json, written from profiles/*_ncu_*.txt), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(kernel_tag)
        except Exception:
            return None
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tensor_burst": p["bf16_tflops"], "tensor_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"tensor_burst": 1590.0, "tensor_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(steps, warmup, min_seconds=0.0):
    """maps/s of the CPU oracle port on cfg 1 (1 query + 5 refs, 518x518), all host threads, fp32."""
    import torch
    from crossscore_b200.synthetic import make_inputs, make_state_dict
    from oracle import crossscore_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    O.FAST = True  # torch's fused CPU kernels (SDPA / layer_norm / gelu): what the reference itself calls on CPU
    sd = make_state_dict(1)
    q, r = make_inputs(1, N_REF, H, W, seed=0)
    with torch.inference_mode():
        for _ in range(warmup):
            O.crossscore_forward(sd, q, r, dt=torch.float32)
        times = []
        t_all = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            O.crossscore_forward(sd, q, r, dt=torch.float32)
            times.append(time.perf_counter() - t0)
        while time.perf_counter() - t_all < min_seconds:
            t0 = time.perf_counter()
            O.crossscore_forward(sd, q, r, dt=torch.float32)
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return len(times) / total, total / len(times), torch.get_num_threads(), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    rate, sec, cores, n = cpu_oracle_rate(steps, min(args.warmup, 1))
    sample = (f"{n} forwards of cfg-1 (1 query + {N_REF} refs, {H}x{W}), fp32, oracle port of the reference on torch's fused CPU "
              f"ops (F.linear / SDPA / layer_norm / gelu: the calls the reference makes; within 5 % of the reference itself "
              f"on the build container, DESIGN.md section 6), {sec:.2f} s each; batch 1 because a batch-32 step takes ~40 s")
    line = {
        "impl": "reference", "metric": "score maps/s (518x518, 5 refs)", "value": rate, "unit": "maps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg1: 1 query x 5 refs, 518x518 per step (bounded sample of cfg2 batch-32)",
                   "global_batch": 1, "image": [H, W], "n_ref": N_REF},
        "cpu_baseline": {"value": rate, "unit": "maps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
def _time_ms(fn, iters, warm, sync):
    """Mean CUDA-event milliseconds of fn() on the current stream (warm untimed calls first)."""
    import torch
    for _ in range(warm):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / iters


def profile_kernels(net, dev, fn, pk):
    """One instrumented pass of fn(): per-tag launch count, ms, achieved TFLOP/s or GB/s (CUDA events per operator)."""
    import torch
    eng = net._engine(dev)
    eng.prof = []
    fn()
    torch.cuda.synchronize()
    agg = {}
    for tag, fl, nb, s, e in eng.prof:
        a = agg.setdefault(tag, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += fl
        a[2] += nb
        a[3] += s.elapsed_time(e)
    eng.prof = None
    step_ms = sum(a[3] for a in agg.values())
    kernels = {}
    for tag, (n, fl, nb, ms) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        ent = {"launches": n, "ms": round(ms, 4), "share": round(ms / step_ms, 4)}
        if fl > 0:
            ent["tflops"] = round(fl / (ms * 1e-3) / 1e12, 2)
            ent["frac_of_tensor_peak_sustained"] = round(ent["tflops"] / pk["tensor_sustained"], 4)
        if nb > 0:  # HBM-bound kernels (LayerNorm, PE add, head + jigsaw: SURVEY 8d) report GB/s against the HBM peak
            ent["gbs"] = round(nb / (ms * 1e-3) / 1e9, 1)
            ent["frac_of_hbm_peak"] = round(ent["gbs"] / pk["hbm"], 4)
        kernels[tag] = ent
    return agg, kernels, step_ms


def extra_single_gpu(net, dev, args, pk):
    """N = 1 only: BASELINE cfg 5, the fp32 parity mode's speed, and the same-box torch comparator."""
    import torch
    from crossscore_b200 import CrossScoreNet, default_cfg
    from crossscore_b200.synthetic import make_inputs, make_state_dict
    sync = torch.cuda.synchronize
    out = {}
    # ---- cfg 5: 1 query x 16 refs at 1036x1036 (74x74 patches, T = 5477, 87 616 reference tokens), bf16 ----
    q5, r5 = (t.to(dev) for t in make_inputs(1, 16, 1036, 1036, seed=5))
    f5 = lambda: net(q5, r5, False, 0, False)
    ms = _time_ms(f5, 5, 2, sync)
    fpm = flops_per_map(5476, 16)
    agg, kernels, _ = profile_kernels(net, dev, f5, pk)
    a = agg["attn_dino"]
    ach = a[1] / (a[3] * 1e-3) / 1e12
    afl = sum(v[1] for t, v in agg.items() if t.startswith("attn_"))
    ams = sum(v[3] for t, v in agg.items() if t.startswith("attn_"))
    out["cfg5"] = {"workload": "cfg5: 1 query x 16 refs, 1036x1036 (P = 5476), bf16, device-resident", "ms_per_map": ms,
                   "maps_per_s": 1e3 / ms, "tflops": fpm / (ms * 1e-3) / 1e12,
                   "frac_of_tensor_peak_sustained": fpm / (ms * 1e-3) / 1e12 / pk["tensor_sustained"],
                   "attention_share_of_flops": attn_flops_per_map(5476, 16) / fpm,
                   "attention_tflops": afl / (ams * 1e-3) / 1e12,
                   "roofline": {"bound": "tensor", "kernel": "attn_dino", "achieved": ach, "peak": pk["tensor_sustained"],
                                "unit": "TFLOP/s", "frac": ach / pk["tensor_sustained"],
                                "algorithmic_flops_per_launch": a[1] / a[0], "avg_launch_ms": a[3] / a[0]},
                   "kernels": {k: kernels[k] for k in list(kernels)[:6]}}
    del q5, r5
    # ---- fp32 parity mode on the headline shape (SIMT fp32 kernels; max-abs <= 1e-4 vs the reference) ----
    B32 = 4
    net32 = CrossScoreNet(default_cfg(), precision="fp32")
    net32.load_state_dict(make_state_dict(1))
    net32 = net32.to(dev).eval()
    q, r = (t.to(dev) for t in make_inputs(B32, N_REF, H, W, seed=100))
    ms = _time_ms(lambda: net32(q, r, False, 0, False), 2, 1, sync)
    got32 = net32(q, r, False, 0, False)["score_map_ref_cross"]
    got16 = net(q, r, False, 0, False)["score_map_ref_cross"]
    d = (got32 - got16).abs()
    out["fp32_mode"] = {"workload": f"cfg2 shape, batch {B32} x {N_REF} refs, {H}x{W}, precision=fp32", "ms_per_step": ms,
                        "maps_per_s": B32 * 1e3 / ms, "tflops": B32 * flops_per_map() / (ms * 1e-3) / 1e12,
                        "bf16_vs_fp32_max_abs": float(d.max()), "bf16_vs_fp32_mean_abs": float(d.mean())}
    del net32
    # ---- same-box library comparator: the oracle's torch ops on this B200, bf16 autocast, eager ----
    try:
        from oracle import crossscore_oracle as O
        sd_dev = {k: v.to(dev) for k, v in make_state_dict(1).items()}
        Bt = args.batch
        qt, rt = (t.to(dev) for t in make_inputs(Bt, N_REF, H, W, seed=100))
        O.FAST = True

        def torch_fwd():
            with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
                return O.crossscore_forward(sd_dev, qt, rt, dt=torch.float32)["score_map_ref_cross"]
        ms = _time_ms(torch_fwd, 3, 2, sync)
        ours = net(qt, rt, False, 0, False)["score_map_ref_cross"]
        dd = (torch_fwd().float() - ours).abs()
        out["torch_gpu"] = {"what": "oracle restatement run by torch on this GPU: F.linear / F.scaled_dot_product_attention / "
                                    "F.layer_norm / F.gelu under torch.autocast(bfloat16), eager, device-resident inputs",
                            "workload": f"cfg2: batch {Bt} x {N_REF} refs, {H}x{W}", "ms_per_step": ms,
                            "value": Bt * 1e3 / ms, "unit": "maps/s",
                            "max_abs_vs_ours": float(dd.max()), "mean_abs_vs_ours": float(dd.mean())}
        O.FAST = False
    except Exception as e:  # the comparator must never take the bench line down
        out["torch_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return out


def extra_multi_gpu(net, dev, world, rank, pk, tmax, barrier):
    """Every N (at N = 1 without the exchange): the two schedules of north-star item (4) that exchange data between GPUs, with parity fields.
    cfg3: a scene of 1024 query frames sharing ONE set of 5 references (reference views sharded over the ranks, K/V of
    both decoder layers exchanged once over NCCL, then queries shard).  cfg4: 1 query x 64 references, reference
    tokens sharded (split-KV cross-attention; partial O + LSE merged per decoder layer over NVLink peer memory or
    one NCCL all-gather)."""
    import torch
    from crossscore_b200.scene import SceneScorer, SplitKVScorer, shard_range
    from crossscore_b200.synthetic import make_inputs
    eng = net._engine(dev)
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.inference_mode():
        # ---------------- cfg 3 ----------------
        Q, Bq = 1024, 32
        _, refs = make_inputs(1, N_REF, H, W, seed=7)
        refs = refs[0].to(dev)
        lo, hi = shard_range(Q, world, rank)
        q_mine, _ = make_inputs(Bq, 1, H, W, seed=100 + rank)  # one resident batch reused for the rank's share
        q_mine = q_mine.to(dev)
        sc = SceneScorer(eng, dev)
        sc.build_reference_cache(refs)
        sc.score(q_mine)
        sc.score(q_mine, graph=True)
        torch.cuda.synchronize()
        eager_ms = _time_ms(lambda: sc.score(q_mine), 8, 2, torch.cuda.synchronize) if rank == 0 else 0.0
        barrier()
        e0.record()
        sc.build_reference_cache(refs)
        e1.record()
        barrier()
        ms_cache = tmax(e0.elapsed_time(e1))
        n_batches = (hi - lo + Bq - 1) // Bq
        # a rank's share of the scene is only a few batches at N = 8 (35 ms): repeat it so the clocks are in their
        # steady state for the timed region, report the time of ONE pass over the scene
        reps = max(1, -(-16 // max(n_batches, 1)))
        for _ in range(n_batches):
            sc.score(q_mine, graph=True)
        barrier()
        e0.record()
        for _ in range(reps * n_batches):
            s = sc.score(q_mine, graph=True)  # one CUDA graph per batch of 32 queries (115 launches eager)
        e1.record()
        barrier()
        ms_score = tmax(e0.elapsed_time(e1)) / reps
        # parity: the same 4 queries through the plain forward (every query carries its own copy of the references)
        want = net(q_mine[:4], refs[None].expand(4, -1, -1, -1, -1).contiguous(), False, 0, False)["score_map_ref_cross"]
        d3 = tmax(float((sc.score(q_mine[:4], graph=True) - want).abs().max()))
        f3 = flops_per_map(query_only=True)
        out["cfg3"] = {"workload": f"cfg3: {Q} query frames sharing one set of {N_REF} refs, {H}x{W}, queries sharded over "
                                   f"{world} GPUs in batches of {Bq}", "cache_build_ms": ms_cache,
                       "cache_bytes_received_per_rank": sc.cache_bytes_received, "score_ms": ms_score,
                       "maps_per_s_excl_cache": Q / (ms_score * 1e-3),
                       "maps_per_s_incl_cache": Q / ((ms_score + ms_cache) * 1e-3),
                       "launch_mode": "one CUDA graph per batch of 32 queries",
                       "ms_per_batch": ms_score / n_batches, "ms_per_batch_eager_rank0": eager_ms,
                       "gflop_per_map": f3 / 1e9, "tflops_excl_cache": Q * f3 / (ms_score * 1e-3) / 1e12,
                       "frac_of_tensor_peak_sustained": Q * f3 / (ms_score * 1e-3) / 1e12 / world / pk["tensor_sustained"],
                       "parity": {"max_abs_vs_plain_forward": d3, "max_over": "ranks, 4 queries each"}}
        # ---------------- cfg 4 ----------------
        q1, r64 = (t.to(dev) for t in make_inputs(1, 64, H, W, seed=9))
        c4 = {"workload": f"cfg4: 1 query x 64 refs, {H}x{W}; reference views sharded over {world} GPUs (split-KV)"}
        got = None
        for mode in (("p2p", "nccl") if world > 1 else ("nccl",)):
            sk = SplitKVScorer(eng, dev, exchange=mode)
            g = sk.forward(q1, r64).clone()
            got = g if got is None else got
            sk.forward(q1, r64)
            barrier()
            e0.record()
            for _ in range(10):
                sk.forward(q1, r64)
            e1.record()
            barrier()
            c4[f"ms_per_query_{sk.exchange}"] = tmax(e0.elapsed_time(e1)) / 10
            c4[f"exchange_bytes_per_query_{sk.exchange}"] = sk.allgather_bytes or sk.peer_bytes_pulled
        full = net(q1, r64, False, 0, False)["score_map_ref_cross"]
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            net(q1, r64, False, 0, False)
        e1.record()
        torch.cuda.synchronize()
        one = e0.elapsed_time(e1) / 3
        c4["ms_per_query_one_gpu"] = tmax(one)
        best = min(v for k, v in c4.items() if k.startswith("ms_per_query_") and k != "ms_per_query_one_gpu")
        c4["speedup_vs_one_gpu"] = c4["ms_per_query_one_gpu"] / best
        if world == 1:
            c4["note"] = "one GPU: the split-KV schedule degenerates to a single part (no exchange); ms_per_query_nccl " \
                         "is that schedule, ms_per_query_one_gpu the plain forward"
        c4["parity"] = {"max_abs_vs_one_gpu": tmax(float((full - got).abs().max())), "max_over": "ranks"}
        out["cfg4"] = c4
    return out

# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from crossscore_b200 import CrossScoreNet, _lib, default_cfg
    from crossscore_b200.runner import HostScorer
    from crossscore_b200.synthetic import make_inputs, make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from crossscore_b200.runner import bind_to_gpu_numa
    # multi-GPU: before any pinned allocation, so that host buffers live on the GPU's own socket (N = 1 keeps every core
    # for the CPU baseline leg)
    numa_cpus = bind_to_gpu_numa(local) if int(os.environ.get("WORLD_SIZE", "1")) > 1 else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, args.warmup

    net = CrossScoreNet(default_cfg(), precision=args.precision)
    net.load_state_dict(make_state_dict(1))
    net = net.to(dev).eval()
    q, r = make_inputs(B, N_REF, H, W, seed=100 + rank)
    q_dev, r_dev = q.to(dev), r.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput --------------------------------------------------------------
    for _ in range(Wm):
        out = net(q_dev, r_dev, False, 0, False)["score_map_ref_cross"]
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = net(q_dev, r_dev, False, 0, False)["score_map_ref_cross"]
    e1.record()
    barrier()
    launches = _lib.launches() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    assert torch.isfinite(out).all()
    value = world * B * K / (ms_total * 1e-3)

    # ---- end to end: pinned host inputs -> public API -> host score maps ---------------------------------
    qh, rh = q.pin_memory(), r.pin_memory()
    scorer = HostScorer(net, dev)
    for _ in range(2):
        scorer.submit(qh, rh)
    barrier()
    e0.record()
    for _ in range(K):
        host_out = scorer.submit(qh, rh)
    scorer.fence()  # the timed region ends when the last score map is in host memory
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * K / (ms_e2e * 1e-3)
    assert bool(torch.isfinite(host_out).all())

    # ---- full chain around the model: uint8 host images -> device preprocessing -> forward
    #      -> device post-processing -> host frame means + uint16 maps ------------------------------------------
    from crossscore_b200.runner import HostPipeline
    g = torch.Generator().manual_seed(200 + rank)
    q8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).pin_memory()
    r8 = torch.randint(0, 256, (B, N_REF, H, W, 3), generator=g, dtype=torch.uint8).pin_memory()
    pipe = HostPipeline(net, dev)
    for _ in range(2):
        pipe.submit(q8, r8)
    barrier()
    e0.record()
    for _ in range(K):
        means_h, maps_h = pipe.submit(q8, r8)
    pipe.fence()
    e1.record()
    barrier()
    ms_pipe = max_over_ranks(e0.elapsed_time(e1))
    pipe_value = world * B * K / (ms_pipe * 1e-3)
    assert bool(torch.isfinite(means_h).all())
    pipeline = {"value": pipe_value, "unit": "maps/s", "ms_per_step": ms_pipe / K,
                "h2d_bytes_per_step": pipe.h2d_bytes(q8, r8), "d2h_bytes_per_step": pipe.d2h_bytes(q8),
                "stages": "uint8 HWC host images -> H2D -> xs_preprocess_u8_resize_normalize -> CrossScoreNet.forward -> "
                          "xs_score_postprocess (frame mean + uint16 map) -> D2H"}

    # ---- single-query latency (BASELINE cfg 1: 1 query + 5 refs): eager launches vs one CUDA graph replay ----------
    from crossscore_b200.runner import GraphedScorer
    q1, r1 = q_dev[:1].contiguous(), r_dev[:1].contiguous()
    graphed = GraphedScorer(net, dev)
    graphed(q1, r1)
    lat = {}
    for name, fn in (("eager", lambda: net(q1, r1, False, 0, False)), ("cuda_graph", lambda: graphed(q1, r1))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        lat[name + "_ms"] = e0.elapsed_time(e1) / 20
    latency = {"workload": "cfg1: 1 query x 5 refs, 518x518, device-resident inputs", **lat}

    # ---- per-kernel roofline (instrumented pass, CUDA events on the launching stream) --------------------
    pk = peaks()
    agg, kernels, step_ms = profile_kernels(net, dev, lambda: net(q_dev, r_dev, False, 0, False), pk)
    top = next(iter(kernels))
    tent = agg[top]
    if tent[1] > 0:
        ach = tent[1] / tent[0] / (tent[3] / tent[0] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor_sustained"], "traffic": ncu_traffic(top),
                "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "algorithmic_flops_per_launch": tent[1] / tent[0], "avg_launch_ms": tent[3] / tent[0]}
    else:
        ach = tent[2] / tent[0] / (tent[3] / tent[0] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": ncu_traffic(top), "peak_source": pk["source"]}
    attn_fl = sum(a[1] for t, a in agg.items() if t.startswith("attn_") and a[1] > 0)
    attn_ms = sum(a[3] for t, a in agg.items() if t.startswith("attn_") and a[1] > 0)
    attn_tf = attn_fl / (attn_ms * 1e-3) / 1e12 if attn_ms > 0 else 0.0

    # ---- the other BASELINE configs (kept out of the headline numbers above) ----------------------------------
    extra = {}
    if not args.no_extra:
        extra = extra_multi_gpu(net, dev, world, rank, pk, max_over_ranks, barrier)  # cfg3 / cfg4 at every N
        if world == 1:
            extra.update(extra_single_gpu(net, dev, args, pk))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) -------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores, n = cpu_oracle_rate(2, 1)
        cpu = {"value": rate, "unit": "maps/s", "cores": cores, "kind": "port",
               "sample": f"{n} forwards of cfg-1 (1 query + {N_REF} refs, {H}x{W}), fp32, oracle port of the reference on torch's "
                         f"fused CPU ops (within 5 % of the reference itself on the build container), {sec:.2f} s each, after 1 "
                         f"warm-up; batch 1 because a batch-32 step takes ~12 s per forward"}

    fpm = flops_per_map()
    line = {
        "metric": "score maps/s (518x518, 5 refs)", "value": value, "unit": "maps/s", "n_gpus": world,
        "steps": K, "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"cfg2: batch {B} queries x {N_REF} refs, {H}x{W}, per GPU (queries shard across ranks)",
                   "global_batch": world * B, "image": [H, W], "n_ref": N_REF, "precision": args.precision,
                   "l2": "inputs exceed L2 (%.0f MB of images per step per GPU)" % ((q.numel() + r.numel()) * 4 / 1e6),
                   "weights": "seeded synthetic state_dict (real ckpt is a git-lfs pointer offline)"},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "maps/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": scorer.h2d_bytes(qh, rh), "d2h_bytes_per_step": int(host_out.numel() * 4),
                "cpus_bound_to_gpu_numa_node": numa_cpus},
        "pipeline": pipeline,
        "latency_cfg1": latency,
        "gpu_launches": launches,
        "clocks": clocks,
        "model_tflops": value * fpm / 1e12,
        "model_frac_of_tensor_peak_sustained": value * fpm / 1e12 / world / pk["tensor_sustained"],
        "attention_tflops": attn_tf, "attention_frac_of_bf16_peak_sustained": attn_tf / pk["tensor_sustained"],
        "attention_frac_of_bf16_peak_burst": attn_tf / pk["tensor_burst"],
        "kernels": kernels,
        "kernel_time_share_of_step": step_ms / (ms_total / K),
        **extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg3/cfg4 (N>1) and cfg5/fp32/torch (N=1) blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun when called directly with --gpus N
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)]
            cmd += [a for a in sys.argv[1:]]
            sys.exit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()
