"""tcgen05/TMEM GEMM (prompt-prefill path) vs a plain fp32 PyTorch reference of the same op."""
import pytest
import torch

from sesameai import _native
from sesameai import synthetic as syn

pytestmark = pytest.mark.gpu


def _run(x, w, epi=0, resid=None):
    N, K = x.shape
    M = w.shape[0]
    y = torch.empty(N, M // 2 if epi == 2 else M, dtype=torch.bfloat16, device="cuda")
    _native.check(_native.lib().csm_k_gemm_tc(x.data_ptr(), w.data_ptr(), N, K, M, y.data_ptr(), epi,
                                              resid.data_ptr() if resid is not None else None,
                                              torch.cuda.current_stream().cuda_stream))
    return y


def _mk(N, K, M, seed):
    x = torch.empty(N, K, device="cuda")
    w = torch.empty(M, K, device="cuda")
    syn.hash_uniform_(x, seed, 1, 1.0)
    syn.hash_uniform_(w, seed, 2, K ** -0.5)
    return x.to(torch.bfloat16), w.to(torch.bfloat16)


@pytest.mark.parametrize("N,K,M", [(128, 2048, 2048), (300, 1024, 1536), (50, 8192, 1024), (1000, 2048, 2051),
                                   (1, 64, 128), (129, 256, 384), (1568, 2048, 3072)])
def test_gemm_matches_fp32_reference(N, K, M):
    x, w = _mk(N, K, M, N + K)
    y = _run(x, w)
    ref = x.float() @ w.float().t()
    tol = 2.0 ** -8 * max(1.0, ref.abs().max().item()) * 1.01  # one bf16 rounding of an fp32-accumulated dot
    assert (y.float() - ref).abs().max().item() <= tol
    # and against the small-row CUDA-core kernel (same rounding point, different summation order)
    y2 = torch.empty_like(y)
    _native.check(_native.lib().csm_k_linear(x.data_ptr(), w.data_ptr(), N, K, M, y2.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)) if K % 256 == 0 else None
    if K % 256 == 0:
        assert (y.float() - y2.float()).abs().max().item() <= 2 * tol


def test_gemm_residual_epilogue():
    x, w = _mk(200, 2048, 2048, 5)
    h = torch.empty(200, 2048, device="cuda")
    syn.hash_uniform_(h, 9, 9, 2.0)
    h = h.to(torch.bfloat16)
    y = _run(x, w, epi=1, resid=h)
    ref = ((x.float() @ w.float().t()).to(torch.bfloat16).float() + h.float()).to(torch.bfloat16)
    assert (y.float() - ref.float()).abs().max().item() <= 2.0 ** -7 * 4


def test_gemm_swiglu_pairs_epilogue():
    x, w = _mk(130, 1024, 512, 6)  # 256 (gate, up) pairs interleaved
    y = _run(x, w, epi=2)
    lin = (x.float() @ w.float().t()).to(torch.bfloat16)
    gate, up = lin[:, 0::2], lin[:, 1::2]
    ref = (torch.nn.functional.silu(gate) * up)
    assert y.shape == (130, 256)
    assert (y.float() - ref.float()).abs().max().item() <= 2.0 ** -7 * max(1.0, ref.float().abs().max().item())
