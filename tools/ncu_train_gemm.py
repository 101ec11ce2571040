"""The training GEMM in its roles at the shapes of the batch-512 step (for ncu --set full -k regex:tc_gemm):
forward fc with GELU16 (M=5120, N=1536, K=384), forward d-wide (N=384, K=1536), dgrad MN-major W (out 384, reduce 1536), dgrad with the
GELU-backward epilogue (out 1536, reduce 384), wgrad MN-major both with split-K ([1536, 384] and [384, 384] from M=5120 rows)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdt_policy_b200 import train_ops as O

torch.manual_seed(0)
M, d, F = 5120, 384, 1536
x, wfc, wpr = torch.randn(M, d, device="cuda"), torch.randn(F, d, device="cuda") / d ** 0.5, torch.randn(d, F, device="cuda") / F ** 0.5
x16, wfc16, wpr16 = O.split(x), O.split(wfc), O.split(wpr)
for _ in range(2):
    h, g16 = O.gemm16(0, x16, wfc16, M, F, d, bias=torch.zeros(F, device="cuda"), epi=O.EPI_GELU16)          # 1 fc forward + GELU16
    f = O.gemm16(0, g16, wpr16, M, d, F)                                                                       # 2 proj forward
    df16 = O.split(torch.randn(M, d, device="cuda"))
    dh16 = O.dgrad_gelu_bwd16(df16, wpr16, h, M, d, F)                                                         # 3 dgrad + GELU backward epilogue
    da = O.gemm16(1, dh16, wfc16, M, F, d)                                                                     # 4 dgrad (reduce 1536)
    dwp = O.gemm16(2, df16, g16, M, d, F)                                                                      # 5 wgrad [384, 1536], split-K
    dwf = O.gemm16(2, dh16, x16, M, F, d)                                                                      # 6 wgrad [1536, 384], split-K
    dwo = O.gemm16(2, df16, x16, M, d, d)                                                                      # 7 wgrad [384, 384], split-K
torch.cuda.synchronize()
print("ok")
