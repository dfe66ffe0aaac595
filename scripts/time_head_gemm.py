"""Head-MLP GEMM shapes (M = masked rows of the bench step): forward, dX and dW."""
import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 28672

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

tot = 0.0
for K, N in ((64, 1024), (1024, 512), (512, 256), (256, 128)):
    A = torch.randn(M, ops.ld8(K), device="cuda").to(torch.bfloat16)
    W = (torch.randn(K, ops.ld8(N), device="cuda") * 0.1).to(torch.bfloat16)
    dY = torch.randn(M, ops.ld8(N), device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ob = torch.empty(M, ops.ld8(N), device="cuda", dtype=torch.bfloat16)
    dxb = torch.empty(M, ops.ld8(K), device="cuda", dtype=torch.bfloat16)
    dW = torch.empty(K, N, device="cuda")
    fl = 2.0 * M * K * N
    f = t(lambda: ops.gemm(A, 0, W, 1, M, N, K, bias=bias, relu=True, out_bf16=ob))
    bx = t(lambda: ops.gemm(dY, 0, W, 0, M, K, N, gate=A, out_bf16=dxb))
    bw = t(lambda: ops.gemm_splitk(A, 1, dY, 1, K, N, M, dW))
    tot += f + bx + bw
    print(f"{K:5d}->{N:5d}: fwd {f:6.1f} us ({fl/f/1e6:5.0f} TF/s)  dX {bx:6.1f} us ({fl/bx/1e6:5.0f})  dW {bw:6.1f} us ({fl/bw/1e6:5.0f})")
print(f"total {tot:.0f} us")
