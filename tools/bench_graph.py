"""Eager launches vs one CUDA-graph replay of the B=32 forward (development tool): how much of the step is launch gaps?"""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200 import CrossScoreNet, default_cfg
from crossscore_b200.synthetic import make_inputs, make_state_dict
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 32))
net = CrossScoreNet(default_cfg(), precision="bf16"); net.load_state_dict(make_state_dict(1)); net = net.to(dev).eval()
q, r = (t.to(dev) for t in make_inputs(B, 5, 518, 518, seed=100))
def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
eager = timeit(lambda: net(q, r, False, 0, False))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    net(q, r, False, 0, False)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    out = net(q, r, False, 0, False)["score_map_ref_cross"]
graph = timeit(g.replay)
eager2 = timeit(lambda: net(q, r, False, 0, False))
print(f"B={B}: eager {eager:.3f} ms, graph replay {graph:.3f} ms, eager again {eager2:.3f} ms -> {B / graph * 1e3:.1f} maps/s graphed")
