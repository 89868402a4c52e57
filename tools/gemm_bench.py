#!/usr/bin/env python
"""time the tcgen05 GEMM on the PPO shapes (CUDA events, L2-flushed between iterations) next to torch.matmul (cuBLAS)"""
import ctypes
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from seqdex_b200 import _lib  # noqa: E402

L = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ONCE = "--once" in sys.argv       # one launch per shape and library (for an ncu capture of every variant): no timing loop
if ONCE:
    def timeit(fn, iters=1):      # noqa: F811
        flush.zero_(); fn(); torch.cuda.synchronize()
        return 1.0
print(f"{'shape':34s} {'mode':>4s} {'ours us':>9s} {'TF/s':>7s} {'cublas us':>10s} {'TF/s':>7s}")
for (N, K, mode, name) in [(1024, 448, 0, "fwd L1"), (512, 1024, 0, "fwd L2"), (256, 512, 0, "fwd L3"), (1024, 512, 1, "dX L2"), (512, 256, 1, "dX L3"),
                           (1024, 448, 3, "plain")]:
    A = (torch.randn(M, K, device="cuda") * 0.1).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    h = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out_t = torch.empty(N + 16, M, device="cuda", dtype=torch.bfloat16)
    outf = torch.empty(M, N, device="cuda")
    fn = lambda: _lib.check(L.sdx_gemm_bf16_tn(mode, p(A), M, K, K, p(B), N, K, p(bias), p(h), N, p(out), N, p(out_t), M, p(outf), N, 1, st()))
    t = timeit(fn)
    tc = timeit(lambda: torch.matmul(A, B.T))
    fl = 2.0 * M * N * K
    print(f"{name + f' [{M}x{N}x{K}]':34s} {mode:4d} {t * 1e3:9.1f} {fl / t / 1e9:7.1f} {tc * 1e3:10.1f} {fl / tc / 1e9:7.1f}")
# dW shapes: [N_l, K_l+16] = dZt[N_l, M] . At[K_l+16, M]^T, split-K
for (Nl, Kl, name) in [(1024, 448, "dW L1"), (512, 1024, "dW L2"), (256, 512, "dW L3")]:
    A = (torch.randn(Nl, M, device="cuda") * 0.1).bfloat16()
    B = (torch.randn(Kl + 16, M, device="cuda") * 0.1).bfloat16()
    outf = torch.zeros(Nl, Kl + 16, device="cuda")
    tiles = ((Nl + 127) // 128) * ((Kl + 16 + 127) // 128)
    t2 = ((Nl + 255) // 256) * ((Kl + 16 + 255) // 256)
    ref = torch.matmul(A.float(), B.float().T)
    for splits in sorted({max(1, (148 + tiles - 1) // tiles), max(1, (296 + tiles - 1) // tiles), max(1, 74 // t2), max(1, 148 // t2)}):
        fn = lambda: _lib.check(L.sdx_gemm_bf16_tn(2, p(A), Nl, M, M, p(B), Kl + 16, M, None, None, 0, None, 0, None, 0, p(outf), Kl + 16, splits, st()))
        outf.zero_(); fn(); torch.cuda.synchronize()
        err = float((outf - ref).abs().max() / ref.abs().max())
        assert err < 2e-3, (name, splits, err)               # split-K fp32 accumulation of bf16 products against the fp32 matmul
        t = timeit(fn)
        tc = timeit(lambda: torch.matmul(A, B.T))
        fl = 2.0 * M * Nl * (Kl + 16)
        print(f"{name + f' [{Nl}x{Kl + 16}x{M}] s={splits}':34s} {2:4d} {t * 1e3:9.1f} {fl / t / 1e9:7.1f} {tc * 1e3:10.1f} {fl / tc / 1e9:7.1f}")
