"""Times fxn_gemm with different fused epilogues on the Gram / Decoder shapes (which epilogue option costs what)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flexynesis_b200 import _lib as L
from flexynesis_b200._lib import Planes

dev = torch.device("cuda", 0)
def planes(r, c):
    p = Planes.empty(r, c, dev)
    x = torch.randn(r, c, device=dev)
    L.split_planes(x, p)
    return p, x

def timeit(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:60s} {a.elapsed_time(b) / n * 1e3:8.1f} us")

for (M, N, K, tag) in ((4096, 4096, 128, "gram"), (4096, 5000, 512, "decoder")):
    A, _ = planes(M, K); B, _ = planes(N, K)
    out = Planes.empty(M, N, dev)
    C = torch.zeros(M, N, device=dev)
    cs = torch.zeros(N, device=dev)
    ra, rb = torch.rand(M, device=dev), torch.rand(N, device=dev)
    X = torch.rand(M, N, device=dev)
    acc = torch.zeros(1, device=dev)
    bias = torch.zeros(N, device=dev)
    timeit(f"{tag}: fp32 C only", lambda: L.gemm(M, N, K, A, 0, B, 0, C_ptr=C.data_ptr(), ldc=N))
    timeit(f"{tag}: planes only", lambda: L.gemm(M, N, K, A, 0, B, 0, out=out))
    timeit(f"{tag}: planes + colsums", lambda: L.gemm(M, N, K, A, 0, B, 0, out=out, colstats=cs.data_ptr(), stats_mode=3))
    timeit(f"{tag}: planes + colsums + gaussian", lambda: L.gemm(M, N, K, A, 0, B, 0, out=out, colstats=cs.data_ptr(), stats_mode=3,
                                                          epi_act=7, gauss_ra=ra.data_ptr(), gauss_rb=rb.data_ptr(), gauss_inv=0.01))
    timeit(f"{tag}: planes + colsums + sigmoid + mse", lambda: L.gemm(M, N, K, A, 0, B, 0, out=out, colstats=cs.data_ptr(), stats_mode=3,
                                                               epi_act=3, bias=bias.data_ptr(), mse_x=X.data_ptr(), ldx=N, mse_acc=acc.data_ptr()))
    timeit(f"{tag}: fp32 C + planes", lambda: L.gemm(M, N, K, A, 0, B, 0, C_ptr=C.data_ptr(), ldc=N, out=out))
