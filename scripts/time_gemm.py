import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
M = 212992
def bench(N, K, f32=False, addend=False, relu=False, n=20):
    A = (torch.randn(M, ops.ld8(K), device="cuda")).to(torch.bfloat16)
    W = (torch.randn(K, ops.ld8(N), device="cuda") * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    of = torch.empty(M, N, device="cuda") if f32 else None
    ob = None if f32 else torch.empty(M, ops.ld8(N), device="cuda", dtype=torch.bfloat16)
    add = torch.randn(M, N, device="cuda") if addend else None
    fn = lambda: ops.gemm(A, 0, W, 1, M, N, K, bias=bias, relu=relu, addend=add, out_f32=of, out_bf16=ob)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    byt = M * K * 2 + M * N * (4 if f32 else 2) + (M * N * 4 if addend else 0)
    print(f"N={N} K={K} f32={f32} addend={addend}: {ms*1e3:.1f} us, {byt/ms/1e6:.0f} GB/s algorithmic")
bench(192, 64)
bench(64, 64, f32=True)
bench(100, 64, relu=True)
bench(64, 100, f32=True)
bench(64, 192, f32=True, addend=True)
