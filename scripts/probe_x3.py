"""fp32-class GEMM (bf16 x 3 split) against float64 on the shapes of the C1 head MLP."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bert4clickpath_b200 import ops
from bert4clickpath_b200.ops import F32, BF16

g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.randn(*s, device="cuda", generator=g)
def err(got, want): return float((got.double() - want).abs().max() / want.abs().max())

M = 448
for (K_in, N_out) in [(64, 1024), (1024, 512), (512, 256), (256, 128), (256, 64), (128, 256)]:
    x = rnd(M, K_in); W = rnd(K_in, N_out) * 0.1; b = rnd(N_out); dy = rnd(M, N_out)
    gate = (rnd(M, K_in) > 0).float() * rnd(M, K_in).abs()
    y = torch.zeros(M, N_out, device="cuda")
    ops.gemm(x, 0, W, 1, M, N_out, K_in, bias=b, relu=True, out_bf16=y)
    e_fwd = err(y, torch.relu(x.double() @ W.double() + b.double()))
    dx = torch.zeros(M, K_in, device="cuda")
    ops.gemm(dy, 0, W, 0, M, K_in, N_out, gate=gate, out_bf16=dx)
    want = (dy.double() @ W.double().t()) * (gate > 0)
    e_in = err(dx, want)
    dx2 = torch.zeros(M, K_in, device="cuda")
    ops.gemm(dy, 0, W, 0, M, K_in, N_out, out_f32=dx2)
    e_in_nogate = err(dx2, dy.double() @ W.double().t())
    dW = torch.zeros(K_in, N_out, device="cuda")
    ops.gemm_splitk(x, 1, dy, 1, K_in, N_out, M, dW)
    e_w = err(dW, x.double().t() @ dy.double())
    torch.cuda.synchronize()
    print(f"K_in={K_in} N_out={N_out}: fwd {e_fwd:.1e}  dx(gated) {e_in:.1e}  dx {e_in_nogate:.1e}  dW {e_w:.1e}", flush=True)
