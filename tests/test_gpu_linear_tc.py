"""tcgen05 fp16x2-split GEMM (csrc/linear_tc.cu) against an fp64 torch reference of the same op.
Stated tolerance: max |err| <= 4e-6 * max |ref| (measured ~1e-6: the tensor core's fp32 accumulation
truncates; a plain fp32 SGEMM measures ~6e-7 on these shapes)."""
import math

import pytest
import torch

from adsorbdiff_b200 import _cabi
from adsorbdiff_b200._cabi import call, ptr

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
A_SCALE, W_SCALE = 16.0, 1024.0


def _split(t, scale, rows):
    m, k = t.shape
    buf = torch.zeros(2 * rows * k, dtype=torch.float16, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_split_f16", DEV, ptr(t), k, m, k, scale, ptr(buf), rows, ptr(status))
    assert int(status.item()) == 0
    return buf


def _ssilu(x):
    return torch.nn.functional.silu(x) / 0.6


@pytest.mark.parametrize("M,N,K", [(7, 256, 256), (82, 512, 512), (300, 1536, 512), (1000, 512, 1024),
                                   (128, 256, 64), (129, 1024, 512), (20992, 1536, 512)])
@pytest.mark.parametrize("act", [_cabi.ACT_NONE, _cabi.ACT_SSILU])
def test_linear_tc_matches_fp64(M, N, K, act):
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, generator=g) * 1.7).to(DEV)
    W = ((torch.rand(N, K, generator=g) * 2 - 1) * math.sqrt(6.0 / (N + K))).to(DEV)
    bias = (torch.rand(N, generator=g) * 0.2 - 0.1).to(DEV)
    rows = (M + 127) // 128 * 128
    a_sp, w_sp = _split(A, A_SCALE, rows), _split(W, W_SCALE, N)
    out = torch.full((M, N), float("nan"), device=DEV)
    out_sp = torch.zeros(2 * rows * N, dtype=torch.float16, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_linear_tc", DEV, ptr(a_sp), rows, M, ptr(w_sp), N, K, ptr(bias), 1.0 / (A_SCALE * W_SCALE), act,
         ptr(out), N, ptr(out_sp), rows, A_SCALE, ptr(status))
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    ref = A.double() @ W.double().T + bias.double()
    if act == _cabi.ACT_SSILU:
        ref = _ssilu(ref)
    scale = float(ref.abs().max())
    err = float((out.double() - ref).abs().max())
    print(f"linear_tc M={M} N={N} K={K} act={act}: err/max = {err / scale:.2e}")
    assert err <= 4e-6 * scale, (err / scale)
    planes = out_sp.view(2, rows, N)[:, :M].double()
    rec = (planes[0] + planes[1]) / A_SCALE
    assert float((rec - ref).abs().max()) <= 5e-6 * scale
    # rows beyond M of the split output are never written
    assert float(out_sp.view(2, rows, N)[:, M:].abs().sum()) == 0.0


def test_split_overflow_sets_status():
    t = torch.full((4, 64), 1.0e5, device=DEV)
    buf = torch.zeros(2 * 128 * 64, dtype=torch.float16, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_split_f16", DEV, ptr(t), 64, 4, 64, 16.0, ptr(buf), 128, ptr(status))
    assert int(status.item()) & _cabi.STATUS_F16_OVERFLOW


def _run_tc(M, N, K, act, seed):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 1.7).to(DEV)
    W = ((torch.rand(N, K, generator=g) * 2 - 1) * math.sqrt(6.0 / (N + K))).to(DEV)
    bias = (torch.rand(N, generator=g) * 0.2 - 0.1).to(DEV)
    rows = (M + 127) // 128 * 128
    a_sp, w_sp = _split(A, A_SCALE, rows), _split(W, W_SCALE, N)
    out = torch.full((M, N), float("nan"), device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_linear_tc", DEV, ptr(a_sp), rows, M, ptr(w_sp), N, K, ptr(bias), 1.0 / (A_SCALE * W_SCALE), act,
         ptr(out), N, None, 0, A_SCALE, ptr(status))
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    return out


def test_cta_pair_variant_matches_single_cta():
    """cta_group::2 path (two CTAs, one MMA over 256 rows, half of the weight tile staged per CTA): bit-identical to
    the single-CTA kernel -- same products, same accumulation order -- including an odd number of row tiles."""
    lib = _cabi.load()
    outs = []
    try:
        for pair in (0, 1):
            lib.adk_set_tc_pair(pair)
            outs.append([_run_tc(M, 1536, 512, _cabi.ACT_SSILU, seed=5) for M in (20992, 19333)])
    finally:
        lib.adk_set_tc_pair(1)   # the default
    for a, b in zip(*outs):
        assert torch.isfinite(a).all() and torch.equal(a, b)
