"""Time the pre-/post-processing kernels alone (development tool): achieved GB/s of algorithmic bytes."""
import os, sys, json, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200 import imgproc

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
for (n, H, W, s) in [(192, 540, 960, 518), (48, 1080, 1920, 518), (192, 518, 518, 518)]:
    u8 = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    H1, W1 = imgproc.resize_output_size(H, W, s)
    out = torch.empty(n, 3, H1, W1, device="cuda")
    ms = timeit(lambda: imgproc.preprocess_u8(u8, s, out=out))
    b = u8.numel() + out.numel() * 4
    res[f"pre_{n}x{H}x{W}->{H1}x{W1}"] = dict(ms=round(ms, 4), gbs=round(b / ms / 1e6, 1), images_per_s=round(n / ms * 1e3))
sc = torch.rand(32, 518, 518, device="cuda")
for name, kw in [("post_mean", dict(mean=True)), ("post_mean_gray16", dict(mean=True, gray16_vrange=[0, 1])),
                 ("post_all", dict(mean=True, gray16_vrange=[0, 1], rgb_vrange=(0, 1)))]:
    ms = timeit(lambda: imgproc.postprocess_scores(sc, **kw))
    b = sc.numel() * (4 + (2 if "gray" in name or "all" in name else 0) + (3 if "all" in name else 0))
    res[name + "_32x518x518"] = dict(ms=round(ms, 4), gbs=round(b / ms / 1e6, 1))
print(json.dumps(res, indent=1))
