"""Time the tensor-core attention kernel alone on the DINOv2 shape (development tool)."""
import math, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200 import _lib
from crossscore_b200._lib import call, DT_BF16, DT_F16
F16 = os.environ.get("F16", "0") == "1"
ADT, DTF = (torch.float16, DT_F16) if F16 else (torch.bfloat16, DT_BF16)
I, H, T, d = int(os.environ.get("I", 48)), 6, 1370, 64
if len(sys.argv) > 1 and sys.argv[1] == "dec":
    I, H, T, d = 32, 8, 1369, 48
Lk = int(os.environ.get("LK", T))
torch.manual_seed(0)
qkv = (torch.randn(I, T, 3 * H * 64, device="cuda") * (0.35 if F16 else 1.0)).to(ADT)
kv = torch.randn(I, Lk, 2 * H * 64, device="cuda").to(ADT)
o = torch.empty(I * T, H * d, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
def run():
    call("xs_flash_attn", qkv.data_ptr(), kv.data_ptr(), kv.data_ptr() + H * 64 * 2, o.data_ptr(), None, I, H, T, Lk, d, 64,
         3 * H * 64, T * 3 * H * 64, 2 * H * 64, Lk * 2 * H * 64, 0, 1, 0, math.log(2.0) if F16 else 1 / math.sqrt(d), DTF, st)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 4.0 * I * H * T * Lk * d
print(f"f16={int(F16)} dbg={os.environ.get('XS_ATTN_DBG','0')} I={I} H={H} Lq={T} Lk={Lk} d={d}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s  "
      f"clk/tile-iter/SM ~ {ms*1e-3*1.9e9/ (I*H*math.ceil(T/128)*math.ceil(Lk/128)/148):.0f}")
