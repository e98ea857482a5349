"""One cfg-3 batch (32 queries against a cached 5-reference K/V) per-operator times (development tool)."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200 import CrossScoreNet, default_cfg
from crossscore_b200.scene import SceneScorer
from crossscore_b200.synthetic import make_inputs, make_state_dict
dev = torch.device("cuda:0")
net = CrossScoreNet(default_cfg(), precision="bf16"); net.load_state_dict(make_state_dict(1)); net = net.to(dev).eval()
eng = net._engine(dev)
_, refs = make_inputs(1, 5, 518, 518, seed=7); refs = refs[0].to(dev)
q, _ = make_inputs(32, 1, 518, 518, seed=100); q = q.to(dev)
sc = SceneScorer(eng, dev); sc.build_reference_cache(refs)
for _ in range(3): sc.score(q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(16): sc.score(q)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 16
eng.prof = []
sc.score(q); torch.cuda.synchronize()
agg = {}
for tag, fl, nb, s, e in eng.prof:
    a = agg.setdefault(tag, [0, 0.0]); a[0] += 1; a[1] += s.elapsed_time(e)
eng.prof = None
print(f"{os.path.basename(os.environ.get('XS_LIB_PATH','default'))}: {ms:.3f} ms per batch of 32 -> {32e3/ms:.0f} maps/s; " +
      ", ".join(f"{t} {v[1]:.3f}({v[0]})" for t, v in sorted(agg.items(), key=lambda kv: -kv[1][1])))
