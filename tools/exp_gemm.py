import ctypes, sys, os, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_scratch", sys.argv[1] if len(sys.argv) > 1 else "cut_t.so"))
lib.run_nn.restype = ctypes.c_int
lib.run_nn.argtypes = [ctypes.c_void_p]*4 + [ctypes.c_int]*3 + [ctypes.c_float]*2 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
dev = torch.device("cuda")
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for (M, K, N) in [(4096, 256, 256), (2449029, 100, 256), (2449029, 256, 256), (2449029, 256, 48), (2449029, 48, 256)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(K, N, device=dev) * 0.1
    D = torch.empty(M, N, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.run_nn(A.data_ptr(), B.data_ptr(), D.data_ptr(), D.data_ptr(), M, N, K, 1.0, 0.0, ws.data_ptr(), ws.numel(), st)
    torch.cuda.synchronize()
    print("shape", M, K, N, "rc", rc)
    if rc != 0: continue
    ref32 = A @ B
    sub = slice(0, 20000)
    ref64 = (A[sub].double() @ B.double())
    e_cut = ((D[sub].double() - ref64).abs().max() / ref64.abs().max()).item()
    e_t32 = ((ref32[sub].double() - ref64).abs().max() / ref64.abs().max()).item()
    t_cut = t(lambda: lib.run_nn(A.data_ptr(), B.data_ptr(), D.data_ptr(), D.data_ptr(), M, N, K, 1.0, 0.0, ws.data_ptr(), ws.numel(), st))
    t_t = t(lambda: torch.matmul(A, B, out=ref32))
    fl = 2.0 * M * K * N
    byt = 4.0 * (M * K + K * N + M * N)
    print(f"   err cutlass9xbf16 {e_cut:.2e}  torch fp32 {e_t32:.2e} | time cutlass {t_cut:.3f} ms ({fl/t_cut/1e9:.0f} TFLOP/s, {byt/t_cut/1e6:.0f} GB/s)  torch {t_t:.3f} ms ({fl/t_t/1e9:.0f} TFLOP/s)")
