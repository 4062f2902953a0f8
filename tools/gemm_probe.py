"""Cold-vs-warm timing of the wide-output GEMM epilogues (diagnostic)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops

dev = "cuda"
torch.manual_seed(0)
M, K = 2688, 2048
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).bfloat16()


def timeit(fn, n=24):
    for i in range(4):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (N, Kk) in ((8192, 2048), (2048, 8192), (6144, 2048)):
    NS = 6
    A = [rnd(M, Kk) for _ in range(NS)]
    W = rnd(N, Kk, sc=0.02)
    bias = rnd(N, sc=0.02)
    O1 = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NS)]
    O2 = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NS)]
    AUX = [rnd(M, N) for _ in range(NS)]
    flops = 2.0 * M * N * Kk
    variants = {
        "store": lambda i, s: ops.gemm(A[s], W, O1[s], M=M, N=N, K=Kk),
        "store+bias": lambda i, s: ops.gemm(A[s], W, O1[s], M=M, N=N, K=Kk, bias=bias),
        "gelu": lambda i, s: ops.gemm(A[s], W, O1[s], M=M, N=N, K=Kk, bias=bias, epi=ops.EPI_GELU),
        "gelu+out2": lambda i, s: ops.gemm(A[s], W, O1[s], M=M, N=N, K=Kk, bias=bias, epi=ops.EPI_GELU, out2=O2[s]),
        "res(no gate)": lambda i, s: ops.gemm(A[s], W, O1[s], M=M, N=N, K=Kk, bias=bias, epi=ops.EPI_GATE_RES, res=AUX[s]),
        "mul_dgelu": lambda i, s: ops.gemm(A[s], W, O1[s], M=M, N=N, K=Kk, epi=ops.EPI_MUL_DGELU, aux=AUX[s]),
    }
    for name, f in variants.items():
        warm = timeit(lambda i: f(i, 0))
        cold = timeit(lambda i: f(i, i % NS))
        print(f"N={N} K={Kk} {name:14s} warm {warm:7.1f} us ({flops/warm/1e6:6.0f} TF/s)   cold {cold:7.1f} us ({flops/cold/1e6:6.0f} TF/s)", flush=True)
    for bn in (128, 192, 256):
        cold = timeit(lambda i: ops.gemm(A[i % NS], W, O1[i % NS], M=M, N=N, K=Kk, block_n=bn))
        print(f"N={N} K={Kk} store bn={bn} cold {cold:7.1f} us ({flops/cold/1e6:6.0f} TF/s)", flush=True)
    del A, O1, O2, AUX
