"""GPU: tcgen05 building blocks (shared-memory descriptors, K-major / MN-major operands, TMEM read-back)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 128), (256, 64), (32, 128), (128, 32), (48, 16)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_umma_gemm(N, K, a_mn, b_mn):
    from keypointfusion_b200 import _lib, ops
    torch.manual_seed(N * 7 + K + a_mn * 2 + b_mn)
    A = torch.randn(128, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    D = torch.zeros(128, N, device="cuda")
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    ops._call("kpf_umma_selftest", ops._p(Ain), ops._p(Bin), ops._p(D), N, K, a_mn, b_mn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert torch.allclose(D, ref, rtol=1e-4, atol=1e-3), (D - ref).abs().max()
