import ctypes, sys, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from seqdex_b200 import _lib
L = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K = 32768, 1024, 448
A = (torch.randn(M, K, device="cuda") * 0.1).bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
bias = torch.randn(N, device="cuda"); h = torch.randn(M, N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); out_t = torch.empty(N + 16, M, device="cuda", dtype=torch.bfloat16)
outf = torch.empty(M, N, device="cuda")
for mode, ot in [(0, out_t), (1, out_t)]:
    for _ in range(2):
        _lib.check(L.sdx_gemm_bf16_tn(mode, p(A), M, K, K, p(B), N, K, p(bias), p(h), N, p(out), N, p(ot), M, p(outf), N, 1, st()))
torch.cuda.synchronize()
