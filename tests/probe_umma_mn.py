"""Descriptor probe for the MN-major B operand (run on the GPU box; prints which (LBO, SBO, k-advance) is right)."""
import ctypes as C
import itertools
import numpy as np
import torch
from melissa_b200 import _lib

L = _lib.lib()
L.mls_test_umma_mn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
L.mls_test_umma_mn.restype = C.c_int
torch.manual_seed(0)
for rows_a, kt in ((128, 64), (50, 64), (50, 48), (60, 16)):
    A = torch.randn(rows_a, 64, device="cuda").to(torch.bfloat16)
    A[:, kt:] = 0 if kt < 64 else A[:, kt:]
    B = torch.randn(64, 128, device="cuda").to(torch.bfloat16)
    want = (A.float()[:, :kt] @ B.float()[:kt]).cpu().numpy()
    for lbo, sbo, kadv in itertools.product((512, 64, 1), (64, 512), (128, 2, 64)):
        D = torch.full((128, 128), float("nan"), device="cuda")
        _lib.check(L.mls_test_umma_mn(A.data_ptr(), B.data_ptr(), D.data_ptr(), rows_a, kt, lbo, sbo, kadv, None))
        torch.cuda.synchronize()
        got = D.cpu().numpy()[:rows_a]
        err = float(np.nanmax(np.abs(got - want))) if np.isfinite(got).all() else float("inf")
        print(f"rows_a={rows_a} kt={kt} lbo={lbo} sbo={sbo} kadv={kadv}: max err {err:.4g} {'OK' if err < 1e-3 else ''}", flush=True)
