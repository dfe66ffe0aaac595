import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
M, N, K = 8192, 8192, 1024
out = torch.empty(M, N, device="cuda")
for a_mn in (0, 1):
    for b_mn in (0, 1):
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
        fn = lambda: ops.gemm(A, a_mn, B, b_mn, M, N, K, out_f32=out)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"a_mn={a_mn} b_mn={b_mn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s")
