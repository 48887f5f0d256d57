import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keras_geometric_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda", 0)
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
M, Kx, N = 16, 32, 64
X = torch.zeros(M, Kx, device=dev); G = torch.zeros(M, N, device=dev)
X[0, 0] = 1.0; X[1, 1] = 2.0; X[3, 5] = 3.0
G[0, 0] = 1.0; G[0, 7] = 5.0; G[1, 2] = 7.0; G[3, 40] = 11.0
npart = lib.kgb_linear_tc_dw_parts(0, M)
parts = torch.full((npart, Kx, N), -7.0, device=dev)
rc = lib.kgb_linear_tc_dw(0, X.data_ptr(), X.stride(0), G.data_ptr(), G.stride(0), M, Kx, N, parts.data_ptr(), npart, st())
torch.cuda.synchronize()
print("rc", rc, "npart", npart)
ref = X.t() @ G
print("ref nz", ref.nonzero().tolist(), ref[ref != 0].tolist())
p0 = parts[0]
print("out nz", p0.nonzero()[:20].tolist(), p0[p0 != 0][:20].tolist())
print("untouched (-7) count", int((p0 == -7).sum()))
