"""cfg 4 on 8 GPUs: what one rank computes (1 query + 8 of the 64 references), eager vs one CUDA graph (development tool)."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200 import CrossScoreNet, default_cfg
from crossscore_b200.runner import GraphedScorer
from crossscore_b200.synthetic import make_inputs, make_state_dict
dev = torch.device("cuda:0")
net = CrossScoreNet(default_cfg(), precision="bf16"); net.load_state_dict(make_state_dict(1)); net = net.to(dev).eval()
def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for nref in (8, 16, 64):
    q, r = (t.to(dev) for t in make_inputs(1, nref, 518, 518, seed=9))
    eager = timeit(lambda: net(q, r, False, 0, False))
    g = GraphedScorer(net, dev); g(q, r)
    graph = timeit(lambda: g(q, r))
    eng = net._engine(dev)
    eng.prof = []
    net(q, r, False, 0, False); torch.cuda.synchronize()
    agg = {}
    for tag, fl, nb, s, e in eng.prof:
        a = agg.setdefault(tag, [0, 0.0]); a[0] += 1; a[1] += s.elapsed_time(e)
    eng.prof = None
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])
    print(f"1 query x {nref} refs: eager {eager:.3f} ms, graph {graph:.3f} ms; per-op ms: " + ", ".join(f"{t} {v[1]:.3f}({v[0]})" for t, v in top))
