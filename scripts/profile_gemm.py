"""Two head-MLP GEMM shapes for `ncu --set full`: 64->1024 forward (bias+ReLU, bf16 out) and the dX
of 1024->512 (gate, bf16 out) at M rows."""
import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 28672
for K, N in ((64, 1024), (1024, 512)):
    A = torch.randn(M, ops.ld8(K), device="cuda").to(torch.bfloat16)
    W = (torch.randn(K, ops.ld8(N), device="cuda") * 0.1).to(torch.bfloat16)
    dY = torch.randn(M, ops.ld8(N), device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ob = torch.empty(M, ops.ld8(N), device="cuda", dtype=torch.bfloat16)
    dxb = torch.empty(M, ops.ld8(K), device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(A, 0, W, 1, M, N, K, bias=bias, relu=True, out_bf16=ob)
        ops.gemm(dY, 0, W, 0, M, K, N, gate=A, out_bf16=dxb)
torch.cuda.synchronize()
print("ok")
