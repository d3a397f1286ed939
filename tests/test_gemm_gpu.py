"""GPU: the hand-written tcgen05/TMEM/TMA bf16 GEMM (csrc/gemm_tcgen05.cu) against a plain
PyTorch fp32 matmul of the same bf16-rounded operands.  Output is bf16, so the tolerance is
one bf16 ulp of the result plus fp32 accumulation-order noise: |err| <= 2^-7 * |ref| + 1e-3."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, B, bias=None, obs=None, nodes=1, relu=False, m_dev=None):
    from melissa_b200 import _lib
    L = _lib.lib()
    L.mls_test_gemm_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    M, K = A.shape
    N = B.shape[0]
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(L.mls_test_gemm_bf16(A.data_ptr(), B.data_ptr(), _lib.ptr(bias), _lib.ptr(obs),
                                    obs.shape[1] * 8 if obs is not None else 0, nodes, out.data_ptr(), M, N, K, int(relu),
                                    _lib.ptr(m_dev), _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    return out


def _check(out, ref, rows=None):
    out, ref = out.float().cpu(), ref.float().cpu()
    if rows is not None:
        out, ref = out[:rows], ref[:rows]
    err = (out - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + 1e-3
    bad = (err > tol).sum().item()
    assert bad == 0, f"{bad} / {err.numel()} entries off; max err {err.max().item():.4g}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 128, 128), (1000, 256, 128), (18900, 1024, 512),
                                   (5, 256, 1152), (777, 1536, 128), (40000, 1024, 128), (129, 256, 256)])
def test_plain_gemm(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    out = _gemm(A, B)
    _check(out, A.float() @ B.float().T)


def test_fused_epilogue_bias_rowscale_relu_and_device_row_count():
    nodes, graphs, N, K = 50, 40, 1024, 512
    M = nodes * graphs
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    obs = torch.rand(graphs, nodes, 8, device="cuda", generator=g)
    obs[:, :, 7] = (torch.rand(graphs, nodes, device="cuda", generator=g) < 0.6).float()
    scale = obs[:, :, 7].reshape(M, 1)
    ref = scale * (A.float() @ B.float().T) + bias
    _check(_gemm(A, B, bias=bias, obs=obs, nodes=nodes), ref)
    _check(_gemm(A, B, bias=bias, obs=obs, nodes=nodes, relu=True), ref.clamp_min(0))
    m_dev = torch.tensor([333], dtype=torch.int32, device="cuda")
    out = _gemm(A, B, bias=bias, m_dev=m_dev)
    _check(out, A.float() @ B.float().T + bias, rows=333)
    assert torch.isnan(out[384:].float()).all()      # tiles past the device-side row count are never written
