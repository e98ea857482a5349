"""Time the tensor-core attention kernel alone (development tool).
  python tools/prof_attn.py [dino|dino192|dec|dsa|cfg5]      env: SCALE1=0/1 (scale folded into q), LAYOUT=0/1
"""
import math, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200 import _lib
from crossscore_b200._lib import call, DT_BF16
shape = sys.argv[1] if len(sys.argv) > 1 else "dino"
I, H, T, Lk, d = {"dino": (48, 6, 1370, 1370, 64), "dino192": (192, 6, 1370, 1370, 64), "dec": (32, 8, 1369, 6845, 48),
                  "dsa": (32, 8, 1369, 1369, 48), "cfg5": (17, 6, 5477, 5477, 64)}[shape]
SCALE1 = os.environ.get("SCALE1", "1") == "1"
LAYOUT = int(os.environ.get("LAYOUT", "1"))
if hasattr(_lib.load(), "xs_attn_set_layout"):
    _lib.load().xs_attn_set_layout(LAYOUT)
torch.manual_seed(0)
qs = (1.4426950408889634 / math.sqrt(d)) if SCALE1 else 1.0
q = (torch.randn(I, T, H * 64, device="cuda") * qs).to(torch.bfloat16)
kv = torch.randn(I, Lk, 2 * H * 64, device="cuda").to(torch.bfloat16)
o = torch.empty(I * T, H * d, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
scale = math.log(2.0) if SCALE1 else 1 / math.sqrt(d)
def run():
    call("xs_flash_attn", q.data_ptr(), kv.data_ptr(), kv.data_ptr() + H * 64 * 2, o.data_ptr(), None, I, H, T, Lk, d, 64,
         H * 64, T * H * 64, 2 * H * 64, Lk * 2 * H * 64, 0, 1, 0, scale, DT_BF16, st)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 4.0 * I * H * T * Lk * d
print(f"{os.path.basename(os.environ.get('XS_LIB_PATH', 'default')):32s} layout={LAYOUT} {shape:8s} I={I} H={H} Lq={T} Lk={Lk} d={d}: "
      f"{ms:.4f} ms  {fl/ms/1e9:.1f} TFLOP/s")
