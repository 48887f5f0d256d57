"""Hand-written tcgen05 dW = X^T G kernel vs float64 / CUTLASS batched TN path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keras_geometric_b200 import _lib, ops
lib = _lib.load(); dev = torch.device("cuda", 0)
st = lambda: torch.cuda.current_stream().cuda_stream
def dw(X, G):
    M, Kx = X.shape; N = G.shape[1]
    npart = lib.kgb_linear_tc_dw_parts(0, M)
    parts = torch.empty((npart, Kx, N), device=dev)
    _lib.check(lib.kgb_linear_tc_dw(0, X.data_ptr(), X.stride(0), G.data_ptr(), G.stride(0), M, Kx, N, parts.data_ptr(), npart, st()), "dw")
    out = torch.empty((Kx, N), device=dev)
    _lib.check(lib.kgb_reduce_parts(0, parts.data_ptr(), npart, Kx * N, out.data_ptr(), st()), "reduce")
    return out
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
shapes = [(16, 32, 64), (1000, 100, 256), (5000, 256, 48), (40000, 132, 200), (2449029, 100, 256), (2449029, 256, 256), (2449029, 256, 48), (2449029, 48, 256)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for (M, Kx, N) in shapes:
    X = torch.randn(M, Kx, device=dev); G = torch.randn(M, N, device=dev)
    out = dw(X, G); torch.cuda.synchronize()
    ref = (X.double().t() @ G.double())
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    e32 = (((X.t() @ G).double() - ref).abs().max() / ref.abs().max()).item()
    msg = f"M={M} Kx={Kx} N={N}: err tc {err:.2e} torch32 {e32:.2e}"
    if M >= 100000:
        t_tc = t(lambda: dw(X, G))
        w = torch.randn(Kx, N, device=dev, requires_grad=True); xx = X
        def cut():
            S = 8192; L, rem = divmod(M, S)
            parts = torch.empty((L + (1 if rem else 0), Kx, N), device=dev)
            ops.dense_gemm(_lib.GEMM_TN, X, G, Kx, N, S, out=parts[:L], L=L, batch=(S * X.stride(0), S * G.stride(0), Kx * N))
            if rem: ops.dense_gemm(_lib.GEMM_TN, X[L*S:], G[L*S:], Kx, N, rem, out=parts[L])
        t_cut = t(cut)
        t_th = t(lambda: torch.matmul(X.t(), G))
        byt = 4.0 * (M * Kx + M * N); fl = 2.0 * M * Kx * N
        msg += f" | tc {t_tc:.3f} ms ({fl/t_tc/1e9:.0f} TF/s, {byt/t_tc/1e6:.0f} GB/s)  cutlass9x {t_cut:.3f} ms  torch {t_th:.3f} ms"
    print(msg, flush=True)
