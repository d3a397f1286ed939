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


@pytest.mark.parametrize("rows_a,kt", [(128, 64), (50, 64), (50, 48), (60, 16)])
def test_umma_mn_major_operand_layout(rows_a, kt):
    """The MN-major, 128B-swizzled operand image the table-mode attention kernel builds by hand (attn_table.cu:
    rows of 128 B per K index, 8-row groups 1024 B apart = SBO, the next 64 columns 8 KiB further = LBO, K-step
    2048 B): one CTA, D[128 x 128] = A[rows_a x 64] (K-major) x B[64 x 128] (MN-major) over the first kt of K."""
    from melissa_b200 import _lib
    L = _lib.lib()
    L.mls_test_umma_mn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
    L.mls_test_umma_mn.restype = C.c_int
    torch.manual_seed(rows_a + kt)
    A = torch.randn(rows_a, 64, device="cuda").to(torch.bfloat16)
    B = torch.randn(64, 128, device="cuda").to(torch.bfloat16)
    D = torch.full((128, 128), float("nan"), device="cuda")
    _lib.check(L.mls_test_umma_mn(A.data_ptr(), B.data_ptr(), D.data_ptr(), rows_a, kt, 8192 >> 4, 1024 >> 4, 2048 >> 4, None))
    torch.cuda.synchronize()
    want = A.float()[:, :kt] @ B.float()[:kt]
    assert float((D[:rows_a] - want).abs().max()) < 1e-3


@pytest.mark.parametrize("with_c,dot_relu,two,compact", [(True, False, False, False), (True, True, True, True), (False, True, True, False)])
def test_epilogue_dots_row_index_and_dots_only(with_c, dot_relu, two, compact):
    """Epilogue extras: per-128-column dot products with one or two vectors (linear parts of the GATv2 logits; the
    dueling heads' output layer), optionally on max(value, 0); row scale through a row index (compacted
    controlling-node rows); C = NULL (dots only); device-side row count."""
    from melissa_b200 import _lib
    L = _lib.lib()
    vp = C.c_void_p
    L.mls_test_gemm_bf16_ex.argtypes = [vp, vp, vp, vp, C.c_longlong, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp,
                                        C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.mls_test_gemm_bf16_ex.restype = C.c_int
    torch.manual_seed(3)
    M, N, K, nodes = 700, 256, 192, 10
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") / 8).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    n_src = 130                                                 # node rows = 130 graphs x 10 nodes
    obs = torch.zeros(n_src, nodes, 8, device="cuda")
    obs[:, :, 7] = (torch.rand(n_src, nodes, device="cuda") > 0.3).float()
    row_index = torch.randint(0, n_src * nodes, (M,), device="cuda", dtype=torch.int32) if compact else None
    dv = torch.randn(N, device="cuda")
    dv2 = torch.randn(N, device="cuda") if two else None
    dots = torch.full((M, N // 128), float("nan"), device="cuda")
    dots2 = torch.full((M, N // 128), float("nan"), device="cuda") if two else None
    Cout = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda") if with_c else None
    m_dev = torch.tensor([650], dtype=torch.int32, device="cuda")
    _lib.check(L.mls_test_gemm_bf16_ex(A.data_ptr(), B.data_ptr(), bias.data_ptr(), obs.data_ptr(), nodes * 8, nodes, _lib.ptr(row_index),
                                       dv.data_ptr(), _lib.ptr(dv2), dots.data_ptr(), _lib.ptr(dots2), int(dot_relu), _lib.ptr(Cout),
                                       M, N, K, 0, m_dev.data_ptr(), _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    rows = row_index.long() if compact else torch.arange(M, device="cuda")
    scale = obs.reshape(-1, 8)[rows, 7]
    x = (A.float() @ B.float().T) * scale[:, None] + bias[None, :]
    xd = x.clamp_min(0) if dot_relu else x
    live = slice(0, 650)
    want = (xd * dv[None, :]).reshape(M, N // 128, 128).sum(2)
    assert torch.allclose(dots[live], want[live], rtol=2e-3, atol=2e-2)
    assert (dots[650:] == 0).all()                              # rows beyond the device-side count: cleared by the launcher, never accumulated
    if two:
        want2 = (xd * dv2[None, :]).reshape(M, N // 128, 128).sum(2)
        assert torch.allclose(dots2[live], want2[live], rtol=2e-3, atol=2e-2)
    if with_c:
        _check(Cout[live], x[live])
        assert torch.isnan(Cout[650:].float()).all()
