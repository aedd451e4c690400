"""Activation epilogues of the GEMM kernel at the edges of their range.

The reference asserts gelu(-100) == 0, gelu(-1) != 0, gelu(0) == 0, gelu(20) == 20 for every activation it ships
(ppdiffusers/tests/models/test_activations.py:30-63: swish / silu / mish / gelu). The erf-GELU epilogue is a fitted
polynomial-sigmoid, so these tests push large-magnitude pre-activations through `ops.linear(..., act=...)` and the
GEGLU / SwiGLU epilogues and compare with torch's exact functions. Tolerance: bf16 output rounding (2^-8 relative)
plus 1e-3 absolute.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16

EDGES = [-100.0, -20.0, -12.0, -11.5, -11.0, -8.0, -7.0, -3.0, -1.0, 0.0, 1.0, 3.0, 7.0, 8.0, 11.0, 11.5, 12.0, 20.0, 100.0]
REF = {1: F.silu, 2: F.gelu, 3: lambda z: F.gelu(z, approximate="tanh"), 4: lambda z: z * torch.sigmoid(1.702 * z)}


@pytest.fixture(scope="module")
def ops():
    from paddlemix_b200 import ops as _ops
    _ops.init(0)
    return _ops


def _edge_problem(n_cols=128, K=64):
    """A[M,K] @ W[N,K]^T whose pre-activation z[m, n] is EXACTLY EDGES[m % len] for every column: A rows are one-hot
    in column 0 scaled by the edge value (bf16-exact), W[:, 0] = 1."""
    M = 128
    a = torch.zeros(M, K)
    vals = torch.tensor([EDGES[m % len(EDGES)] for m in range(M)])
    a[:, 0] = vals
    w = torch.zeros(n_cols, K)
    w[:, 0] = 1.0
    return a.to(bf16).cuda(), w.to(bf16).cuda(), vals.to(bf16).float()


@pytest.mark.parametrize("act", [1, 2, 3, 4])
def test_activation_epilogue_edges(ops, act):
    a, w, vals = _edge_problem()
    out = ops.linear(a, w, None, act=act).float().cpu()
    ref = REF[act](vals)[:, None].expand_as(out)
    err = (out - ref).abs()
    assert (err <= 1e-3 + 2 ** -8 * ref.abs()).all(), (act, err.max().item(), vals[err.max(1).values.argmax()].item())
    # the reference's own assertions (test_activations.py:55-63)
    i_m100, i_0, i_20, i_m1 = EDGES.index(-100.0), EDGES.index(0.0), EDGES.index(20.0), EDGES.index(-1.0)
    assert out[i_m100, 0].item() == 0.0
    assert out[i_0, 0].item() == 0.0
    assert out[i_m1, 0].item() != 0.0
    if act != 4:  # quick-GELU(20) = 20 * sigmoid(34.04) rounds to 20 as well, asserted below for all
        assert out[i_20, 0].item() == 20.0
    assert out[i_20, 0].item() == 20.0


@pytest.mark.parametrize("glu", [1, 2])
def test_glu_epilogue_edges(ops, glu):
    """value * act(gate) with gates at the edge values: a negative outlier gate must zero the product (the round-1
    polynomial returned gate itself for gate <= -12)."""
    a, w1, vals = _edge_problem(n_cols=64)
    K = a.shape[1]
    # interleaved columns: 2j = value (constant 3.0 through A's column 1), 2j+1 = gate (edge value)
    a[:, 1] = 1.0
    w = torch.zeros(128, K, dtype=bf16, device="cuda")
    w[0::2, 1] = 3.0
    w[1::2, 0] = 1.0
    out = ops.linear(a, w, None, glu=glu).float().cpu()
    gate = F.gelu(vals) if glu == 1 else F.silu(vals)
    ref = (3.0 * gate)[:, None].expand_as(out)
    err = (out - ref).abs()
    assert out.shape == (128, 64)
    assert (err <= 3e-3 + 2 ** -8 * ref.abs()).all(), (glu, err.max().item())
    assert out[EDGES.index(-100.0), 0].item() == 0.0


def test_geglu_wide_gates(ops):
    """GEGLU with N(0, 6^2) gates (real SDXL feed-forward gates have outliers well beyond |x| = 11)."""
    M, N2, K = 512, 256, 128
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).to(bf16).cuda()
    wv = (torch.randn(N2, K, generator=g) * K ** -0.5).to(bf16).cuda()
    wg = (torch.randn(N2, K, generator=g) * 6.0 * K ** -0.5).to(bf16).cuda()
    w = torch.stack([wv, wg], 1).reshape(2 * N2, K).contiguous()
    out = ops.linear(a, w, None, glu=1).float()
    val, gate = a.float() @ wv.float().t(), a.float() @ wg.float().t()
    assert gate.abs().max().item() > 15.0  # the test actually reaches the old failure range
    ref = val * F.gelu(gate)
    err = (out - ref).abs()
    assert (err <= 2e-2 + 1.5e-2 * ref.abs()).all(), err.max().item()
