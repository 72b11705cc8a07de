"""GPU parity tests for the raw C-ABI kernels (GEMM, in-batch CE fwd/bwd, helpers).

All calls go through libtt_b200.so via two_tower_models_b200.ops; references are the CPU oracle /
plain fp32 torch on the CPU.  Integer-grid inputs make the tensor-core results exact in fp32, so layout
or descriptor mistakes show up as hard mismatches rather than as tolerance noise.
"""
import pytest
import torch

import oracle
from helpers import assert_close_fro, bf16_round, rel_fro

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _report(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    bad = (got != ref)
    msg = f"{name}: max|diff|={float((got - ref).abs().max()):.4g}, mismatched {int(bad.sum())}/{bad.numel()}"
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        msg += " first bad " + ", ".join(f"{tuple(i)}: got {float(got[tuple(i)]):.4g} ref {float(ref[tuple(i)]):.4g}" for i in idx)
    return msg


def _grid(shape, g, lo=-3, hi=4):
    return torch.randint(lo, hi, shape, generator=g).float()


GEMM_CASES = [
    # M, N, K, a_mn, b_mn
    (128, 128, 64, False, False),
    (128, 64, 128, False, False),
    (256, 256, 256, False, False),
    (300, 200, 136, False, False),
    (1000, 384, 128, False, False),
    (128, 128, 128, False, True),
    (128, 128, 128, True, False),
    (128, 128, 128, True, True),
    (256, 128, 1000, True, True),
    (50, 40, 333, True, True),
    (333, 72, 40, False, True),
    (4096, 256, 128, False, False),
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", GEMM_CASES)
def test_gemm_exact_grid(M, N, K, a_mn, b_mn):
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = _grid((M, K), g)
    B = _grid((N, K), g)
    ref = A @ B.t()
    dev = _dev()
    A16 = (A.t().contiguous() if a_mn else A).to(dev)
    B16 = (B.t().contiguous() if b_mn else B).to(dev)
    A16 = ops.cast_rows_bf16(A16)
    B16 = ops.cast_rows_bf16(B16)
    out32 = torch.full((M, N), float("nan"), device=dev)
    out16 = torch.empty((M, ops._r8(N)), dtype=torch.bfloat16, device=dev)
    ops.gemm(A16, B16, M, N, K, a_mn=a_mn, b_mn=b_mn, out32=out32, out16=out16)
    torch.cuda.synchronize()
    assert torch.equal(out32.cpu(), ref), _report("gemm f32", out32, ref)
    assert torch.equal(out16[:, :N].float().cpu(), bf16_round(ref)), _report("gemm bf16", out16[:, :N], bf16_round(ref))


def test_gemm_epilogue_bias_relu_mask_alpha():
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(11)
    M, N, K = 200, 96, 72
    A, B = _grid((M, K), g), _grid((N, K), g)
    bias = _grid((N,), g)
    mask = _grid((M, N), g, -1, 2)
    dev = _dev()
    A16, B16 = ops.cast_rows_bf16(A.to(dev)), ops.cast_rows_bf16(B.to(dev))
    mask16 = ops.cast_rows_bf16(mask.to(dev))
    out = torch.empty((M, N), device=dev)
    ops.gemm(A16, B16, M, N, K, bias=bias.to(dev), relu=True, out32=out)
    assert torch.equal(out.cpu(), torch.relu(A @ B.t() + bias)), _report("bias+relu", out, torch.relu(A @ B.t() + bias))
    ops.gemm(A16, B16, M, N, K, relu_mask=mask16, alpha=0.5, out32=out)
    ref = 0.5 * (A @ B.t()) * (mask > 0)
    assert torch.equal(out.cpu(), ref), _report("mask+alpha", out, ref)


def test_gemm_fused_column_sums():
    """colsum += sum over rows of the final (masked) outputs, taken from the fp32 accumulators."""
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(21)
    M, N, K = 777, 200, 96
    A, B = _grid((M, K), g), _grid((N, K), g)
    mask = _grid((M, N), g, -1, 2)
    dev = _dev()
    A16, B16, mask16 = ops.cast_rows_bf16(A.to(dev)), ops.cast_rows_bf16(B.to(dev)), ops.cast_rows_bf16(mask.to(dev))
    out16 = torch.empty((M, ops._r8(N)), dtype=torch.bfloat16, device=dev)
    cs = torch.zeros(N, device=dev)
    ops.gemm(A16, B16, M, N, K, relu_mask=mask16, out16=out16, colsum=cs)
    ref = (A @ B.t()) * (mask > 0)
    assert torch.equal(cs.cpu(), ref.sum(0)), _report("colsum", cs, ref.sum(0))
    assert torch.equal(out16[:, :N].float().cpu(), bf16_round(ref))


def test_gemm_split_k_accumulate():
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(12)
    M, N, K = 256, 128, 4096  # weight-gradient shape: tiny output, long reduction over the batch
    A, B = _grid((K, M), g, -2, 3), _grid((K, N), g, -2, 3)
    dev = _dev()
    A16, B16 = ops.cast_rows_bf16(A.to(dev)), ops.cast_rows_bf16(B.to(dev))
    out = torch.zeros((M, N), device=dev)
    ops.gemm(A16, B16, M, N, K, a_mn=True, b_mn=True, out32=out, accumulate=True)
    ref = A.t() @ B
    assert torch.equal(out.cpu(), ref), _report("split-k", out, ref)


def test_gemm_randn_tolerance():
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(13)
    M, N, K = 512, 256, 256
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    dev = _dev()
    out = torch.empty((M, N), device=dev)
    ops.gemm(ops.cast_rows_bf16(A.to(dev)), ops.cast_rows_bf16(B.to(dev)), M, N, K, out32=out)
    ref = bf16_round(A).double() @ bf16_round(B).double().t()
    assert rel_fro(out, ref) < 1e-5  # fp32 accumulation of exact bf16 products


def test_gather_scatter_colsum_cast():
    from two_tower_models_b200 import ops

    g = torch.Generator().manual_seed(3)
    dev = _dev()
    table = torch.randn(97, 50, generator=g)
    ids = torch.randint(0, 97, (333,), generator=g)
    out = torch.zeros((333, 56), dtype=torch.bfloat16, device=dev)
    ops.gather_rows(table.to(dev), ids.to(dev), out)
    assert torch.equal(out[:, :50].float().cpu(), bf16_round(table[ids]))
    assert float(out[:, 50:].float().abs().sum()) == 0.0
    out32 = torch.empty((333, 50), device=dev)
    ops.gather_rows(table.to(dev), ids.to(dev), out32)
    assert torch.equal(out32.cpu(), table[ids])
    src = torch.randn(333, 50, generator=g)
    grad = ops.scatter_add_rows(src.to(dev), ids.to(dev), 50, 97)
    ref = torch.zeros(97, 50).index_add_(0, ids, src)
    assert_close_fro(grad, ref, rtol=1e-6, what="scatter_add")
    cs = ops.colsum(src.to(dev), 50)
    assert_close_fro(cs, src.sum(0), rtol=1e-5, what="colsum f32")
    s16 = ops.cast_rows_bf16(src.to(dev))
    assert s16.shape == (333, 56)
    assert torch.equal(s16[:, :50].float().cpu(), bf16_round(src)) and float(s16[:, 50:].float().abs().sum()) == 0
    cs16 = ops.colsum(s16, 50)
    assert_close_fro(cs16, bf16_round(src).sum(0), rtol=1e-5, what="colsum bf16")


CE_CASES = [
    # B, N, d, target_offset
    (128, 128, 64, 0),
    (512, 512, 64, 0),      # config 1 shape
    (32, 32, 40, 0),        # reference unit-test shape (DI=40)
    (300, 300, 128, 0),
    (256, 1024, 128, 512),  # rank 2 of 4, all-gathered items
    (200, 456, 256, 101),
    (1024, 1024, 256, 0),
    (2048, 2048, 128, 0),
    (2048, 2048, 256, 0),     # d = 256: the shared-memory-operand variant of the backward, several tiles per CTA
    (128, 32768, 128, 4096),  # few users against many (all-gathered) items: the dV pass walks several row tiles per CTA
    (4096, 160, 64, 0) if False else (96, 20000, 64, 77),
]


def _ce_inputs(B, N, d, seed, scale=0.4):
    g = torch.Generator().manual_seed(seed)
    U = bf16_round(torch.randn(B, d, generator=g) * scale)
    V = bf16_round(torch.randn(N, d, generator=g) * scale)
    return U, V


@pytest.mark.parametrize("B,N,d,off", CE_CASES)
def test_inbatch_ce_forward(B, N, d, off):
    from two_tower_models_b200 import ops

    U, V = _ce_inputs(B, N, d, B + N + d)
    ce_ref, lse_ref = oracle.inbatch_ce(U.double(), V.double(), off)
    dev = _dev()
    ce, lse = ops.inbatch_ce_forward_raw(ops.cast_rows_bf16(U.to(dev)), ops.cast_rows_bf16(V.to(dev)), B, N, d, off)
    torch.cuda.synchronize()
    # tolerance: inputs are bf16-exact, accumulation fp32, ex2.approx ~2 ulp -> abs 2e-5 on O(10) values
    err_lse = float((lse.cpu().double() - lse_ref).abs().max())
    err_ce = float((ce.cpu().double() - ce_ref).abs().max())
    assert err_lse < 5e-5 and err_ce < 5e-5, (err_lse, err_ce, _report("lse", lse, lse_ref.float()))


@pytest.mark.parametrize("B,N,d,off", CE_CASES)
def test_inbatch_ce_backward(B, N, d, off):
    from two_tower_models_b200 import ops

    U, V = _ce_inputs(B, N, d, 2 * B + N + d)
    g = torch.rand(B, generator=torch.Generator().manual_seed(5)) / B
    _, lse_ref = oracle.inbatch_ce(U.double(), V.double(), off)
    dU_ref, dV_ref = oracle.inbatch_ce_backward(U.double(), V.double(), lse_ref, g.double(), off)
    dev = _dev()
    dU, dV, dU16, dV16 = ops.inbatch_ce_backward_raw(
        ops.cast_rows_bf16(U.to(dev)), ops.cast_rows_bf16(V.to(dev)), B, N, d, off,
        lse_ref.float().to(dev), g.to(dev))
    torch.cuda.synchronize()
    # dS is rounded to bf16 before the second tensor-core GEMM: relative error ~2^-9 per term, averaging down
    assert_close_fro(dU, dU_ref, rtol=4e-3, what="dU " + _report("dU", dU, dU_ref.float()))
    assert_close_fro(dV, dV_ref, rtol=4e-3, what="dV " + _report("dV", dV, dV_ref.float()))
    assert_close_fro(dU16[:, :d].float(), dU_ref, rtol=8e-3, what="dU bf16")
    assert_close_fro(dV16[:, :d].float(), dV_ref, rtol=8e-3, what="dV bf16")


@pytest.mark.parametrize("B,N,d,off", [(300, 300, 128, 0), (256, 1024, 128, 512), (512, 512, 64, 0), (200, 456, 256, 101)])
def test_inbatch_ce_backward_signed_and_scaled_g(B, N, d, off):
    """Upstream g of either sign, exact zeros, and the two device scalars (incoming d loss, weight normalisation)
    multiplied in by the kernels: the statistics-in-the-MMA backward carries |g| in the exponent and the signs
    separately (row sign in the dU pass, per-user sign words in the dV pass)."""
    from two_tower_models_b200 import ops

    U, V = _ce_inputs(B, N, d, 3 * B + N + d)
    gen = torch.Generator().manual_seed(11)
    g = (torch.rand(B, generator=gen) - 0.4) / B
    g[::7] = 0.0
    s1, s2 = torch.tensor([-1.7]), torch.tensor([0.35])
    _, lse_ref = oracle.inbatch_ce(U.double(), V.double(), off)
    dU_ref, dV_ref = oracle.inbatch_ce_backward(U.double(), V.double(), lse_ref, (g * s1 * s2).double(), off)
    dev = _dev()
    dU, dV, dU16, dV16 = ops.inbatch_ce_backward_raw(
        ops.cast_rows_bf16(U.to(dev)), ops.cast_rows_bf16(V.to(dev)), B, N, d, off,
        lse_ref.float().to(dev), g.to(dev), g_scale=s1.to(dev), g_scale2=s2.to(dev))
    torch.cuda.synchronize()
    assert_close_fro(dU, dU_ref, rtol=4e-3, what="dU " + _report("dU", dU, dU_ref.float()))
    assert_close_fro(dV, dV_ref, rtol=4e-3, what="dV " + _report("dV", dV, dV_ref.float()))
    zero_rows = dU[::7].abs().max()
    assert float(zero_rows) == 0.0, "rows with g = 0 must get an exactly zero dU"


def test_inbatch_ce_autograd_function_matches_oracle_autograd():
    from two_tower_models_b200 import ops

    B, d = 384, 128
    U, V = _ce_inputs(B, B, d, 77)
    w = torch.rand(B, generator=torch.Generator().manual_seed(6))
    Uc, Vc = U.clone().requires_grad_(True), V.clone().requires_grad_(True)
    ce_ref, _ = oracle.inbatch_ce(Uc, Vc)
    (ce_ref * w).mean().backward()
    dev = _dev()
    Ug, Vg = U.to(dev).requires_grad_(True), V.to(dev).requires_grad_(True)
    ce = ops.inbatch_cross_entropy(Ug, Vg)
    (ce * w.to(dev)).mean().backward()
    assert float((ce.detach().cpu() - ce_ref.detach()).abs().max()) < 1e-4
    assert_close_fro(Ug.grad, Uc.grad, rtol=4e-3, what="dU")
    assert_close_fro(Vg.grad, Vc.grad, rtol=4e-3, what="dV")


def test_inbatch_ce_full_size_properties():
    """Config-2 size (B=8192, d=128): size-independent properties instead of a CPU oracle pass.
    (1) sum_j dS_ij = 0  =>  sum over items of dV equals 0-weighted ... we check sum_i dU_i.V-identity:
        sum_j dV_j = sum_i (sum_j dS_ij) U_i = 0;  (2) lse >= max_j S_ij >= S_ii  => ce >= 0;
    (3) a 1024-row slice agrees with the oracle."""
    from two_tower_models_b200 import ops

    B, d = 8192, 128
    U, V = _ce_inputs(B, B, d, 99)
    dev = _dev()
    U16, V16 = ops.cast_rows_bf16(U.to(dev)), ops.cast_rows_bf16(V.to(dev))
    ce, lse = ops.inbatch_ce_forward_raw(U16, V16, B, B, d, 0)
    assert bool((ce >= -1e-4).all())
    sl = slice(3000, 4024)
    ce_ref, lse_ref = oracle.inbatch_ce(U[sl].double(), V.double(), 3000)
    assert float((lse[sl].cpu().double() - lse_ref).abs().max()) < 1e-4
    assert float((ce[sl].cpu().double() - ce_ref).abs().max()) < 1e-4
    g = torch.full((B,), 1.0 / B, device=dev)
    dU, dV, _, _ = ops.inbatch_ce_backward_raw(U16, V16, B, B, d, 0, lse, g)
    col_sum = dV.double().sum(0).abs().max()
    scale = dV.double().abs().sum(0).max()
    assert float(col_sum) < 2e-3 * float(scale), (float(col_sum), float(scale))
    dU_ref, _ = oracle.inbatch_ce_backward(U[sl].double(), V.double(), lse_ref, torch.full((1024,), 1.0 / B, dtype=torch.float64), 3000)
    assert_close_fro(dU[sl], dU_ref, rtol=4e-3, what="dU slice")
    # (4) a 512-item slice of dV against the definition  dV[sl] = dS[:, sl]^T U  (all 8192 users, fp64 on the host);
    # the row statistics are the device's own lse, already pinned by (3)
    csl = slice(5000, 5512)
    lse64 = lse.cpu().double()
    P = torch.exp(U.double() @ V[csl].double().t() - lse64[:, None])
    P[torch.arange(5000, 5512), torch.arange(512)] -= 1.0
    dV_ref = (P / B).t() @ U.double()
    assert_close_fro(dV[csl], dV_ref, rtol=4e-3, what="dV slice")


def test_inbatch_ce_config5_shape_slices():
    """BASELINE configs[4] per-rank shape: 8192 local users against 65 536 all-gathered items at d = 256, positives at
    column row + 3 * 8192 (rank 3 of 8).  Checked on slices against the fp64 definition: ce / lse and dU of 512 users
    (full rows of S), dV of 256 items (all 8192 users; the row statistics are the device's own lse, pinned by the first check)."""
    from two_tower_models_b200 import ops

    B, N, d, off = 8192, 65536, 256, 3 * 8192
    U, V = _ce_inputs(B, N, d, 123, scale=0.3)
    dev = _dev()
    U16, V16 = ops.cast_rows_bf16(U.to(dev)), ops.cast_rows_bf16(V.to(dev))
    ce, lse = ops.inbatch_ce_forward_raw(U16, V16, B, N, d, off)
    sl = slice(2048, 2560)
    ce_ref, lse_ref = oracle.inbatch_ce(U[sl].double(), V.double(), off + 2048)
    assert float((lse[sl].cpu().double() - lse_ref).abs().max()) < 1e-4
    assert float((ce[sl].cpu().double() - ce_ref).abs().max()) < 1e-4
    g = (torch.rand(B, generator=torch.Generator().manual_seed(9)) + 0.5) / N
    dU, dV, _, _ = ops.inbatch_ce_backward_raw(U16, V16, B, N, d, off, lse, g.to(dev))
    torch.cuda.synchronize()
    dU_ref, _ = oracle.inbatch_ce_backward(U[sl].double(), V.double(), lse_ref, g[sl].double(), off + 2048)
    assert_close_fro(dU[sl], dU_ref, rtol=4e-3, what="dU slice")
    csl = slice(off + 5000, off + 5256)  # items whose positives are users 5000..5255
    lse64 = lse.cpu().double()
    P = torch.exp(U.double() @ V[csl].double().t() - lse64[:, None])
    P[torch.arange(5000, 5256), torch.arange(256)] -= 1.0
    dV_ref = (P * g.double()[:, None]).t() @ U.double()
    assert_close_fro(dV[csl], dV_ref, rtol=4e-3, what="dV slice (items with positives)")
    csl2 = slice(100, 356)  # items of another rank: no positives among the local users
    P2 = torch.exp(U.double() @ V[csl2].double().t() - lse64[:, None])
    assert_close_fro(dV[csl2], (P2 * g.double()[:, None]).t() @ U.double(), rtol=4e-3, what="dV slice (foreign items)")
