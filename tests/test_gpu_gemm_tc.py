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


# (decode-sized row counts run cluster split-K with 2 / 4 / 8 slices; the wide shapes -- a 256-stream gate/up -- the
#  persistent kernel with partly filled last row tiles)
@pytest.mark.parametrize("N,K,M", [(128, 2048, 2048), (300, 1024, 1536), (50, 8192, 1024), (1000, 2048, 2051),
                                   (1, 64, 128), (129, 256, 384), (1568, 2048, 3072), (256, 1024, 16384), (200, 512, 9600),
                                   (384, 256, 6400)])
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


@pytest.mark.parametrize("N,M", [(130, 512), (256, 16384)])  # cluster split-K; the persistent kernel on a 256-stream gate/up
def test_gemm_swiglu_pairs_epilogue(N, M):
    x, w = _mk(N, 1024, M, 6)  # M / 2 (gate, up) pairs interleaved
    y = _run(x, w, epi=2)
    lin = (x.float() @ w.float().t()).to(torch.bfloat16)
    gate, up = lin[:, 0::2], lin[:, 1::2]
    ref = (torch.nn.functional.silu(gate) * up)
    assert y.shape == (N, M // 2)
    assert (y.float() - ref.float()).abs().max().item() <= 2.0 ** -7 * max(1.0, ref.float().abs().max().item())


@pytest.mark.parametrize("N,K,M,epi", [(256, 8192, 1024, 1), (64, 1024, 1024, 0), (128, 1024, 1536, 0), (512, 2048, 2048, 1),
                                       (256, 1024, 2051, 0), (200, 8192, 2048, 1), (96, 1024, 512, 2)])
def test_split_k_matches_the_single_pass_gemm(N, K, M, epi):
    """Decode-sized row counts with few output tiles run split-K (partials in fp32, added in slice order by the
    last CTA of a tile): same result as the single-pass kernel up to the fp32 summation order, deterministic,
    and the arrival counters are left at zero."""
    x, w = _mk(N, K, M, N + K + M)
    resid = None
    if epi == 1:
        resid = torch.empty(N, M, device="cuda")
        syn.hash_uniform_(resid, 3, 3, 1.0)
        resid = resid.to(torch.bfloat16)
    part = torch.empty(148 * 128 * 128, dtype=torch.float32, device="cuda")
    counters = torch.zeros(148, dtype=torch.int32, device="cuda")

    def run_split():
        y = torch.empty(N, M // 2 if epi == 2 else M, dtype=torch.bfloat16, device="cuda")
        _native.check(_native.lib().csm_k_gemm_tc_splitk(x.data_ptr(), w.data_ptr(), N, K, M, y.data_ptr(), epi,
                                                         resid.data_ptr() if resid is not None else None, part.data_ptr(),
                                                         counters.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return y

    a, b = run_split(), run_split()
    assert torch.equal(a, b)
    assert int(counters.abs().sum()) == 0
    ref = _run(x, w, epi=epi, resid=resid)
    scale = max(1.0, ref.float().abs().max().item())
    assert (a.float() - ref.float()).abs().max().item() <= 2.0 ** -7 * scale
    assert (a == ref).float().mean().item() >= 0.99  # only the last bit of a few sums moves with the order
