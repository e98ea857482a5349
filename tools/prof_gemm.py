"""Time one tensor-core GEMM shape alone (development tool):  python tools/prof_gemm.py M N K [act]"""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from crossscore_b200._lib import call, DT_BF16
M, N, K = (int(x) for x in sys.argv[1:4])
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
def run():
    call("xs_gemm_bias_act", A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), out.data_ptr(), N, M, N, K, act, DT_BF16, DT_BF16, st)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"PAIR={os.environ.get('XS_GEMM_PAIR','1')} M={M} N={N} K={K} act={act}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s  "
      f"HBM {(M*K*2 + M*N*2)/ms/1e6:.0f} GB/s")
