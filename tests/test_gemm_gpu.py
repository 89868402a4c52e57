"""tcgen05 GEMM (csrc/sdx_gemm.cuh) against a plain PyTorch fp32 reference of the same contraction.
Operands are bf16 (exactly representable in fp32), accumulation fp32 -> tolerance is the bf16 rounding of
the OUTPUT for bf16 modes (rel 2^-8) and fp32 summation-order noise for fp32 modes (rel 1e-5 * sqrt(K))."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(mode, A, B, bias=None, h=None, out=None, out_t=None, outf=None, splits=1):
    from seqdex_b200 import _lib
    L = _lib.load()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    M, K = A.shape
    N = B.shape[0]
    _lib.check(L.sdx_gemm_bf16_tn(mode, p(A), M, K, A.stride(0), p(B), N, B.stride(0), p(bias), p(h), h.stride(0) if h is not None else 0,
                                  p(out), out.stride(0) if out is not None else 0, p(out_t), out_t.stride(0) if out_t is not None else 0,
                                  p(outf), outf.stride(0) if outf is not None else 0, splits,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (1024, 1024, 448), (200, 464, 192), (2048, 512, 1024)])
def test_plain_fp32(M, N, K):
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    _gemm(3, A, B, outf=out)
    ref = A.float() @ B.float().T
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-3)


def test_forward_bias_elu_and_transposed_copy():
    torch.manual_seed(1)
    M, N, K = 512, 384, 448
    A = (torch.randn(M, K, device="cuda") * 0.2).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.2).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    out_t = torch.zeros(N + 16, M, device="cuda", dtype=torch.bfloat16)
    _gemm(0, A, B, bias=bias, out=out, out_t=out_t)
    ref = torch.nn.functional.elu(A.float() @ B.float().T + bias)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(out_t[:N].float(), ref.T.contiguous(), rtol=1e-2, atol=1e-2)
    assert (out_t[N:] == 0).all()


def test_backward_dx_elu_grad():
    torch.manual_seed(2)
    M, N, K = 384, 256, 512
    dZ = (torch.randn(M, K, device="cuda") * 0.3).bfloat16()      # A: [M, N_l]
    Wt = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()      # B: W^T stored [K_l, N_l]
    h = torch.nn.functional.elu(torch.randn(M, N, device="cuda")).bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    out_t = torch.zeros(N, M, device="cuda", dtype=torch.bfloat16)
    _gemm(1, dZ, Wt, h=h, out=out, out_t=out_t)
    hf = h.float()
    ref = (dZ.float() @ Wt.float().T) * torch.where(hf > 0, torch.ones_like(hf), hf + 1)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(out_t.float(), ref.T.contiguous(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("splits", [1, 4, 7])
def test_backward_dw_split_k(splits):
    torch.manual_seed(3)
    M, N, K = 256, 272, 4096      # dW[N_l, K_l+16] = dZt[N_l, batch] . Ht[K_l+16, batch]^T
    A = (torch.randn(M, K, device="cuda") * 0.1).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    out = torch.zeros(M, N, device="cuda")
    _gemm(2, A, B, outf=out, splits=splits)
    ref = A.float() @ B.float().T
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("M,N,K", [(32768, 512, 256), (20000 + 8, 328, 192)])
def test_wide_tiles_forward_and_dx(M, N, K):
    """large outputs take the 128 x 256 tile path (sdx_ppo.cu: `wide`); ragged M and N exercise the clipped TMA boxes"""
    torch.manual_seed(4)
    A = (torch.randn(M, K, device="cuda") * 0.2).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.2).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    out_t = torch.zeros(N + 16, M, device="cuda", dtype=torch.bfloat16)
    _gemm(0, A, B, bias=bias, out=out, out_t=out_t)
    acc = A.float() @ B.float().T
    ref = torch.nn.functional.elu(acc + bias)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(out_t[:N].float(), ref.T.contiguous(), rtol=1e-2, atol=1e-2)
    assert (out_t[N:] == 0).all()
    h = torch.nn.functional.elu(torch.randn(M, N, device="cuda")).bfloat16()
    out2 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    out2_t = torch.zeros(N, M, device="cuda", dtype=torch.bfloat16)
    _gemm(1, A, B, h=h, out=out2, out_t=out2_t)
    hf = h.float()
    ref2 = acc * torch.where(hf > 0, torch.ones_like(hf), hf + 1)
    torch.testing.assert_close(out2.float(), ref2, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(out2_t.float(), ref2.T.contiguous(), rtol=1e-2, atol=2e-2)
