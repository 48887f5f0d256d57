"""Timing + spot check of the tcgen05 GEMM kernels at C4's shapes for the library named by KGB200_LIB
(tuning variants from tools/build_variant.sh).  Prints one JSON line {shape: ms}."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keras_geometric_b200 import ops
dev = torch.device("cuda", 0)
M = 2_449_029
gen = torch.Generator(device=dev).manual_seed(0)
res = {"lib": os.path.basename(os.environ.get("KGB200_LIB", "default"))}


def t(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(statistics.median(ts), 3)


def check(out, want, what):
    err = float((out.double() - want).abs().max() / want.abs().max())
    assert err < 1e-5, (what, err)
    return err


worst = 0.0
for K, N in [(256, 256), (100, 256), (256, 48)]:
    a = torch.randn((M, K), device=dev, generator=gen)
    w = torch.randn((K, N), device=dev, generator=gen) * 0.1
    c = torch.randn((M, N), device=dev, generator=gen)
    b = torch.randn(N, device=dev, generator=gen)
    hi, lo = ops._split_weight(w, transpose=True)
    res[f"lin_{K}x{N}"] = t(lambda: ops.linear_tc(a, hi, lo, N))
    res[f"lin_{K}x{N}_c_bias_relu"] = t(lambda: ops.linear_tc(a, hi, lo, N, c=c, bias=b, relu=True))
    o = ops.linear_tc(a, hi, lo, N, c=c, bias=b, relu=True)
    s = slice(M - 5000, M)
    worst = max(worst, check(o[s], torch.relu(a[s].double() @ w.double() + c[s].double() + b.double()), (K, N)))
    if K == 256 and N == 256:
        a2 = torch.randn((M, K), device=dev, generator=gen)
        w2 = torch.randn((K, N), device=dev, generator=gen) * 0.1
        hi2, lo2 = ops._split_weight_pair(w, w2, transpose=True)
        res["lin2_512x256_bias_relu"] = t(lambda: ops.linear_tc2(a, a2, hi2, lo2, N, bias=b, relu=True))
        o = ops.linear_tc2(a, a2, hi2, lo2, N, bias=b, relu=True)
        worst = max(worst, check(o[s], torch.relu(a[s].double() @ w.double() + a2[s].double() @ w2.double() + b.double()), "tc2"))
        g = torch.randn((M, N), device=dev, generator=gen)
        res["dw_256x256"] = t(lambda: ops._dw_tc(a, g))
        del a2, g
    del a, c
res["max_rel_err"] = worst
print(json.dumps(res))
