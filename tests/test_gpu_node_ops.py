"""Unit parity of the node-wise kernels (csrc/node_ops.cu) against plain PyTorch fp32 formulas of the reference
lines they replace, through the C ABI.  Tolerance: 2e-6 of the tensor's max magnitude (same arithmetic, different
association); fp16x2 operand planes must reconstruct the fp32 output to 2^-21 relative."""
import math

import pytest
import torch

from adsorbdiff_b200 import _cabi
from adsorbdiff_b200._cabi import call, ptr

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TOL = 2e-6
SCALE = 16.0


def _planes(buf, rows, K, M, scale):
    p = buf.view(2, rows, K)[:, :M].double()
    return (p[0] + p[1]) / scale


def _close(a, b, tol=TOL):
    return float((a.double() - b.double()).abs().max()) <= tol * max(float(b.double().abs().max()), 1e-30)


@pytest.mark.parametrize("N,F", [(5, 64), (83, 512), (300, 192)])
def test_layernorm_matches_torch(N, F):
    g = torch.Generator().manual_seed(N + F)
    x = (torch.randn(N, F, generator=g) * 2 + 0.3).to(DEV)
    gamma, beta = (torch.rand(F, generator=g) + 0.5).to(DEV), (torch.randn(F, generator=g) * 0.1).to(DEV)
    rows = (N + 127) // 128 * 128
    y = torch.empty(N, F, device=DEV)
    sp = torch.zeros(2 * rows * F, dtype=torch.float16, device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_layernorm", DEV, ptr(x), ptr(gamma), ptr(beta), N, F, 1e-5, ptr(y), ptr(sp), rows, SCALE, ptr(st))
    ref = torch.nn.functional.layer_norm(x, (F,), gamma, beta, 1e-5)   # painn_denoising.py:531
    assert _close(y, ref) and int(st.item()) == 0
    assert _close(_planes(sp, rows, F, N, SCALE), y, 2.0 ** -20)


@pytest.mark.parametrize("N,F", [(7, 64), (83, 512)])
def test_update_prep_and_gate_match_torch(N, F):
    g = torch.Generator().manual_seed(3 * N + F)
    x = torch.randn(N, F, generator=g).to(DEV)
    vec = (torch.randn(N, 3, F, generator=g) * 0.1).to(DEV)
    vp = (torch.randn(N, 3, 2 * F, generator=g) * 0.3).to(DEV)
    h = torch.randn(N, 3 * F, generator=g).to(DEV)
    scale = torch.tensor(1.37, device=DEV)
    rows, rows3 = (N + 127) // 128 * 128, (3 * N + 127) // 128 * 128
    dot, cat = torch.empty(N, F, device=DEV), torch.empty(N, 2 * F, device=DEV)
    sp = torch.zeros(2 * rows * 2 * F, dtype=torch.float16, device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_update_prep", DEV, ptr(x), ptr(vp), N, F, ptr(dot), ptr(cat), ptr(sp), rows, SCALE, ptr(st))
    vec1, vec2 = vp[..., :F], vp[..., F:]
    dot_ref = (vec1 * vec2).sum(dim=1) / math.sqrt(F)                       # painn_denoising.py:602-606
    cat_ref = torch.cat([x, torch.sqrt((vec2 ** 2).sum(dim=1) + 1e-8)], dim=-1)  # :608-613
    assert _close(dot, dot_ref) and _close(cat, cat_ref)
    assert _close(_planes(sp, rows, 2 * F, N, SCALE), cat, 2.0 ** -20)

    x2, v2 = x.clone(), vec.clone()
    vsp = torch.zeros(2 * rows3 * F, dtype=torch.float16, device=DEV)
    call("adk_update_gate", DEV, ptr(h), ptr(dot), ptr(vp), ptr(scale), N, F, ptr(x2), ptr(v2), ptr(vsp), rows3, 1024.0, ptr(st))
    a, b, c = h[:, :F], h[:, F:2 * F], h[:, 2 * F:]
    x_ref = (x + (a + b * dot) / math.sqrt(2.0)) * scale                    # :614-623, 449-451, scale_factor.py:157-172
    v_ref = vec + c.unsqueeze(1) * vec1
    assert _close(x2, x_ref) and _close(v2, v_ref) and int(st.item()) == 0
    assert _close(_planes(vsp, rows3, F, 3 * N, 1024.0), v2.view(3 * N, F), 2.0 ** -20)
    # an unfitted ScaleFactor (0) means "no multiply"
    x3, v3 = x.clone(), vec.clone()
    zero = torch.zeros((), device=DEV)
    call("adk_update_gate", DEV, ptr(h), ptr(dot), ptr(vp), ptr(zero), N, F, ptr(x3), ptr(v3), None, 0, 0.0, ptr(st))
    assert _close(x3, x + (a + b * dot) / math.sqrt(2.0))


@pytest.mark.parametrize("N,C,Co", [(9, 64, 32), (83, 512, 256), (40, 256, 1)])
def test_head_prep_and_gate_match_torch(N, C, Co):
    g = torch.Generator().manual_seed(N + C + Co)
    x = torch.randn(N, C, generator=g).to(DEV)
    v1p = (torch.randn(N, 3, C, generator=g) * 0.2).to(DEV)
    rows, rows3 = (N + 127) // 128 * 128, (3 * N + 127) // 128 * 128
    cat = torch.empty(N, 2 * C, device=DEV)
    sp = torch.zeros(2 * rows * 2 * C, dtype=torch.float16, device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("adk_head_prep", DEV, ptr(x), ptr(v1p), N, C, ptr(cat), ptr(sp), rows, SCALE, ptr(st))
    cat_ref = torch.cat([x, torch.norm(v1p, dim=-2)], dim=-1)               # painn_denoising.py:688-692
    assert _close(cat, cat_ref) and _close(_planes(sp, rows, 2 * C, N, SCALE), cat, 2.0 ** -20)

    u = torch.randn(N, 2 * Co, generator=g).to(DEV)
    v2p = (torch.randn(N, 3, Co, generator=g) * 0.2).to(DEV)
    xo, vo = torch.empty(N, Co, device=DEV), torch.empty(N, 3, Co, device=DEV)
    planes = Co % 4 == 0
    vsp = torch.zeros(2 * rows3 * Co, dtype=torch.float16, device=DEV)
    call("adk_head_gate", DEV, ptr(u), ptr(v2p), N, Co, ptr(xo), ptr(vo), ptr(vsp) if planes else None, rows3, 1024.0, ptr(st))
    s, gate = u[:, :Co], u[:, Co:]
    assert _close(xo, torch.nn.functional.silu(s) / 0.6)                    # ScaledSiLU, base_layers.py:65-72
    assert _close(vo, gate.unsqueeze(1) * v2p) and int(st.item()) == 0      # :693-697
    if planes:
        assert _close(_planes(vsp, rows3, Co, 3 * N, 1024.0), vo.view(3 * N, Co), 2.0 ** -20)


def test_embed_matches_torch():
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(83, 128, generator=g).to(DEV)
    z = torch.randint(1, 84, (50,), generator=g).to(DEV)
    x = torch.empty(50, 128, device=DEV)
    call("adk_embed", DEV, ptr(z), ptr(emb), 83, 50, 128, ptr(x), None, None)
    assert torch.equal(x, emb[z - 1])                                       # embedding_block.py:42
