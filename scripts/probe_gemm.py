"""GPU probe for the tcgen05 GEMM building block: all operand-major combinations vs torch."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from bert4clickpath_b200 import _lib as L  # noqa: E402


def run(M, N, K, a_mn, b_mn, splits=1, bias=False, relu=False, bf16_out=False):
    torch.manual_seed(M * 7 + N * 3 + K)
    dev = "cuda"
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    A_st = A.t().contiguous() if a_mn else A.contiguous()
    B_st = B.t().contiguous() if b_mn else B.contiguous()
    lda = A_st.shape[1]
    ldb = B_st.shape[1]
    ep = L.GemmEpilogue()
    ep.alpha = 1.0
    bvec = None
    if bias:
        bvec = torch.randn(N, device=dev)
        ep.bias = bvec.data_ptr()
        ref = ref + bvec
    if relu:
        ep.relu = 1
        ref = ref.clamp_min(0)
    out = torch.full((max(splits, 1), M, N), float("nan"), device=dev)
    ep.out_f32 = out.data_ptr()
    ep.ld_f32 = N
    ep.split_stride = M * N
    outb = None
    if bf16_out:
        outb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        ep.out_bf16 = outb.data_ptr()
        ep.ld_bf16 = N
    L.call("b4cp_gemm_bf16", L.ptr(A_st), L.c_int(a_mn), L.c_long(lda), L.ptr(B_st),
           L.c_int(b_mn), L.c_long(ldb), L.c_int(M), L.c_int(N), L.c_int(K), L.c_int(splits),
           ctypes.byref(ep), L.stream_ptr())
    torch.cuda.synchronize()
    got = out.sum(0) if splits > 1 else out[0]
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= 2e-3 * max(scale, 1.0)
    msg = f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} splits={splits} bias={bias} relu={relu}: max_err={err:.3e} scale={scale:.2f} {'OK' if ok else 'FAIL'}"
    if bf16_out:
        e2 = (outb.float() - ref).abs().max().item()
        msg += f" bf16_err={e2:.3e}"
    print(msg, flush=True)
    return ok


if __name__ == "__main__":
    a_mn, b_mn = int(sys.argv[1]), int(sys.argv[2])
    L.call("b4cp_device_check")
    ok = True
    for (M, N, K) in [(128, 64, 64), (128, 256, 128), (256, 128, 192), (300, 200, 136),
                      (1000, 64, 64), (77, 104, 72), (512, 1024, 512)]:
        ok &= run(M, N, K, a_mn, b_mn)
    ok &= run(384, 256, 1024, a_mn, b_mn, splits=4)
    ok &= run(300, 200, 136, a_mn, b_mn, bias=True, relu=True, bf16_out=True)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)
