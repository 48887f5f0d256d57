"""Hand-written tcgen05 3xTF32 GEMM (kgb_linear_tc) vs float64 / torch fp32 / the CUTLASS bf16x9 path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keras_geometric_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda", 0)
st = lambda: torch.cuda.current_stream().cuda_stream
def split(w, transpose):
    rows, cols = w.shape
    n_out, k = (cols, rows) if transpose else (rows, cols)
    bn = lib.kgb_linear_tc_rows(n_out)
    hi = torch.zeros((bn, k), device=dev); lo = torch.zeros((bn, k), device=dev)
    _lib.check(lib.kgb_split_tf32(0, w.data_ptr(), rows, cols, w.stride(0), int(transpose), hi.data_ptr(), lo.data_ptr(), st()), "split")
    return hi, lo
def tc(a, hi, lo, n, c=None, bias=None, relu=False, out=None):
    M, K = a.shape
    out = torch.empty((M, n), device=dev) if out is None else out
    _lib.check(lib.kgb_linear_tc(0, a.data_ptr(), a.stride(0), M, K, hi.data_ptr(), lo.data_ptr(), n,
                                 c.data_ptr() if c is not None else None, c.stride(0) if c is not None else 0,
                                 bias.data_ptr() if bias is not None else None, 1 if relu else 0, out.data_ptr(), out.stride(0), st()), "linear_tc")
    return out
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
shapes = [(128, 32, 64), (1000, 100, 256), (4096, 256, 256), (2449029, 100, 256), (2449029, 256, 256), (2449029, 256, 48), (2449029, 48, 256)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for (M, K, N) in shapes:
    A = torch.randn(M, K, device=dev); W = torch.randn(K, N, device=dev) * 0.1
    hi, lo = split(W, True)
    D = tc(A, hi, lo, N); torch.cuda.synchronize()
    sub = slice(0, min(M, 20000))
    ref64 = A[sub].double() @ W.double()
    e_tc = ((D[sub].double() - ref64).abs().max() / ref64.abs().max()).item()
    e_32 = (((A[sub] @ W).double() - ref64).abs().max() / ref64.abs().max()).item()
    tail = slice(max(0, M - 300), M)
    e_tail = ((D[tail].double() - A[tail].double() @ W.double()).abs().max() / ref64.abs().max()).item()
    # epilogue: + C + bias, relu
    C = torch.randn(M, N, device=dev); b = torch.randn(N, device=dev)
    D2 = tc(A, hi, lo, N, c=C, bias=b, relu=True)
    e_epi = ((D2[sub].double() - torch.relu(ref64 + C[sub].double() + b.double())).abs().max() / ref64.abs().max()).item()
    msg = f"M={M} K={K} N={N}: err tc {e_tc:.2e} (tail {e_tail:.2e}, epilogue {e_epi:.2e}) torch32 {e_32:.2e}"
    if M >= 100000:
        t_tc = t(lambda: tc(A, hi, lo, N, out=D))
        t_cut = t(lambda: ops.dense_gemm(_lib.GEMM_NN, A, W, M, N, K, out=D))
        t_th = t(lambda: torch.matmul(A, W, out=D))
        byt = 4.0 * (M * K + M * N); fl = 2.0 * M * K * N
        msg += f" | tc {t_tc:.3f} ms ({fl/t_tc/1e9:.0f} TF/s, {byt/t_tc/1e6:.0f} GB/s)  cutlass9x {t_cut:.3f} ms  torch {t_th:.3f} ms"
    print(msg, flush=True)
