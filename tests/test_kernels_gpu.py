"""Kernel-level parity: every C-ABI entry point against a plain PyTorch fp32 reference of the same op.

Tolerances (stated per test): inputs are bf16, accumulation is fp32 on both sides, so GEMM-type outputs must agree
to bf16 output rounding (rel 2^-8) plus summation-order noise; fp32 outputs to ~1e-5 relative.
"""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _lib():
    from vit_plasticity_b200 import _lib

    return _lib


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _report(name, got, ref, atol, rtol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    if bad.any():
        idx = torch.nonzero(bad)
        rows = idx[:, 0]
        cols = idx[:, 1] if idx.shape[1] > 1 else idx[:, 0]
        msg = (
            f"{name}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {float(err.max()):.4g} "
            f"(ref max {float(ref.abs().max()):.4g}); first bad at {idx[0].tolist()} got {float(got[tuple(idx[0])]):.5g} "
            f"ref {float(ref[tuple(idx[0])]):.5g}; bad rows range [{int(rows.min())},{int(rows.max())}] "
            f"cols range [{int(cols.min())},{int(cols.max())}]"
        )
        if got.dim() == 2:
            # coarse 32x32 block map of the failures (first 16x16 blocks)
            bm = bad[: 32 * 16, : 32 * 16]
            pr, pc = (-bm.shape[0]) % 32, (-bm.shape[1]) % 32
            bm = torch.nn.functional.pad(bm, (0, pc, 0, pr))
            blk = bm.reshape(bm.shape[0] // 32, 32, bm.shape[1] // 32, 32).any(3).any(1)
            msg += "\nblock map (32x32):\n" + "\n".join("".join("X" if v else "." for v in r) for r in blk.tolist())
        pytest.fail(msg)


# ---------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------
GEMM_MODES = ["pair", "single", "pair+steal", "single+steal"]


def pytest_generate_tests(metafunc):
    """Every GEMM test runs under both tile mappings — CTA pairs (256 x 256, tcgen05 cta_group::2; the default) and single
    CTAs (128 x 256) — and under both tile orders (static stride, work stealing). Other tests run once."""
    if metafunc.function.__name__.startswith("test_gemm"):
        metafunc.fixturenames.append("_gemm_mode")
        metafunc.parametrize("_gemm_mode", GEMM_MODES, indirect=True)


@pytest.fixture
def _gemm_mode(request):
    lib = _lib().lib()
    before, before_s = lib.vb_get_gemm_cta_pair(), lib.vb_get_gemm_scheduler()
    lib.vb_set_gemm_cta_pair(1 if request.param.startswith("pair") else 0)
    lib.vb_set_gemm_scheduler(1 if request.param.endswith("+steal") else 0)
    yield request.param
    lib.vb_set_gemm_cta_pair(before)
    lib.vb_set_gemm_scheduler(before_s)


def test_gemm_pair_and_single_mappings_agree_bitwise():
    """Same k order per output element in both mappings: bf16 / fp32 outputs must be identical, not just close."""
    L = _lib()
    lib = L.lib()
    m, n, k = 1000, 768, 1536
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    bias = _rand(n, seed=3)
    outs = []
    for pair in (1, 0):
        lib.vb_set_gemm_cta_pair(pair)
        o16 = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        o32 = torch.empty(m, n, device=DEV, dtype=torch.float32)
        L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16, bias=bias, out=o16)
        L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_F32, out=o32)
        torch.cuda.synchronize()
        outs.append((o16, o32))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("m", [1000, 12608 // 4, 300])
def test_gemm_tile_192_agrees_bitwise_with_256(m):
    """The 192-column tile variants (N = 768 GEMMs at few token rows: proj / fc2 forward with the residual epilogue, K-major
    B staged as a 96-row box per CTA; fc1 / qkv dgrad, MN-major B staged as two 64-column boxes) keep the k order of every
    output element: identical bits to the 256-column tiles, ragged last row block included; and both match fp32 torch."""
    L = _lib()
    lib = L.lib()
    before, before_pair = lib.vb_get_gemm_tile_n(), lib.vb_get_gemm_cta_pair()
    lib.vb_set_gemm_cta_pair(1)
    try:
        for n, k in ((768, 768), (768, 3072), (384, 320)):
            a = _rand(m, k, seed=1).bfloat16()
            w = _rand(n, k, seed=2, scale=0.05).bfloat16()  # forward weight [n_out = n, k_in = k]
            bias = _rand(n, seed=3)
            res = _rand(m, n, seed=4).bfloat16()
            dy = _rand(m, k, seed=5).bfloat16()              # dgrad: dx [m, n] = dy [m, k] @ W [k, n], W stored [k, n]
            wd = _rand(k, n, seed=6, scale=0.05).bfloat16()
            outs = []
            for tile in (256, 192):
                lib.vb_set_gemm_tile_n(tile)
                o_res = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
                o_dg = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
                L.gemm(a, w, m=m, n=n, k=k, epilogue=L.EPI_BF16_RESID, bias=bias, aux=res, out=o_res)
                L.gemm(dy, wd, m=m, n=n, k=k, b_layout=1, epilogue=L.EPI_BF16, out=o_dg)
                torch.cuda.synchronize()
                outs.append((o_res, o_dg))
            assert torch.equal(outs[0][0], outs[1][0]), (n, k)
            assert torch.equal(outs[0][1], outs[1][1]), (n, k)
            _report("gemm_tile192_resid", outs[1][0], a.float() @ w.float().T + bias + res.float(), atol=3e-2, rtol=1e-2)
            _report("gemm_tile192_dgrad", outs[1][1], dy.float() @ wd.float(), atol=3e-2, rtol=1e-2)
    finally:
        lib.vb_set_gemm_tile_n(before)
        lib.vb_set_gemm_cta_pair(before_pair)


GEMM_SHAPES = [(128, 256, 64), (256, 256, 128), (384, 512, 256), (1000, 768, 768), (197 * 3, 2304, 768), (130, 136, 72)]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_kmajor_bias(m, n, k):
    L = _lib()
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    bias = _rand(n, seed=3)
    out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16, bias=bias, out=out)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T + bias
    _report("gemm_bias", out, ref, atol=2e-2, rtol=1e-2)


def test_gemm_nobias_f32_out():
    L = _lib()
    m, n, k = 300, 512, 320
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float32)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_F32, out=out)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T
    _report("gemm_f32", out, ref, atol=1e-4, rtol=1e-4)


def test_gemm_residual():
    L = _lib()
    m, n, k = 788, 768, 3072
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.02).bfloat16()
    bias = _rand(n, seed=3)
    res = _rand(m, n, seed=4).bfloat16()
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16_RESID, bias=bias, aux=res, out=out)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T + bias + res.float()
    _report("gemm_resid", out, ref, atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("m,n,k", [(300, 136, 192), (257, 768, 768), (1000, 520, 64)])
def test_gemm_residual_short_k_ragged(m, n, k):
    """K < 2048 takes the TMA-fed residual path of the pair kernel: ragged M (odd CTA of the last pair partly / wholly
    beyond M) and N that ends inside a 32-column chunk."""
    L = _lib()
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    bias = _rand(n, seed=3)
    res = _rand(m, n, seed=4).bfloat16()
    out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16_RESID, bias=bias, aux=res, out=out)
    torch.cuda.synchronize()
    _report("gemm_resid_ragged", out, a.float() @ b.float().T + bias + res.float(), atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("tokens,n_out,k_in", [(333, 136, 200), (130, 64, 72)])
def test_gemm_dgrad_mulaux_ragged(tokens, n_out, k_in):
    L = _lib()
    dy = _rand(tokens, n_out, seed=1).bfloat16()
    w = _rand(n_out, k_in, seed=2, scale=0.05).bfloat16()
    z = _rand(tokens, k_in, seed=3).bfloat16()
    out = torch.full((tokens, k_in), float("nan"), device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(k_in, device=DEV, dtype=torch.float32)
    L.gemm(dy, w, m=tokens, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16_MULAUX, aux=z, out=out, out_colsum=cs)
    torch.cuda.synchronize()
    _report("dgrad_mulaux_ragged", out, (dy.float() @ w.float()) * z.float(), atol=3e-2, rtol=1e-2)
    _report("dgrad_mulaux_ragged_colsum", cs, out.float().sum(0), atol=2e-3, rtol=1e-4)


def test_gemm_gelu_two_outputs():
    L = _lib()
    m, n, k = 394, 3072, 768
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    bias = _rand(n, seed=3)
    act = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    z = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16_GELU, bias=bias, out=act, out2=z)
    torch.cuda.synchronize()
    zref = a.float() @ b.float().T + bias
    _report("gemm_gelu.z", z, zref, atol=2e-2, rtol=1e-2)
    # activation must be gelu of the *stored* (bf16-rounded) pre-activation
    _report("gemm_gelu.a", act, torch.nn.functional.gelu(z.float()), atol=1e-3, rtol=1e-2)
    # training variant: out2 = gelu'(z) (exact-erf derivative), out = gelu(z) of the fp32 pre-activation
    gp = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16_GELU_GRAD, bias=bias, out=act, out2=gp)
    torch.cuda.synchronize()
    zf = zref.clone().requires_grad_(True)
    g = torch.nn.functional.gelu(zf)
    g.sum().backward()
    _report("gemm_gelu_grad.a", act, g.detach(), atol=2e-2, rtol=1e-2)
    _report("gemm_gelu_grad.dg", gp, zf.grad, atol=1e-2, rtol=1e-2)
    # inference variant: no second output
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_BF16_GELU, bias=bias, out=act)
    torch.cuda.synchronize()
    _report("gemm_gelu_single.a", act, torch.nn.functional.gelu(zref), atol=2e-2, rtol=1e-2)


def test_gemm_dgrad_mn_major_b_with_dgelu():
    L = _lib()
    tokens, n_out, k_in = 394, 768, 3072  # dx[tokens, k_in] = dy[tokens, n_out] @ W[n_out, k_in]
    dy = _rand(tokens, n_out, seed=1).bfloat16()
    w = _rand(n_out, k_in, seed=2, scale=0.05).bfloat16()
    z = _rand(tokens, k_in, seed=3).bfloat16()
    out = torch.empty(tokens, k_in, device=DEV, dtype=torch.bfloat16)
    L.gemm(dy, w, m=tokens, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16, out=out)
    torch.cuda.synchronize()
    ref = dy.float() @ w.float()
    _report("dgrad", out, ref, atol=3e-2, rtol=1e-2)
    L.gemm(dy, w, m=tokens, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16_DGELU, aux=z, out=out)
    torch.cuda.synchronize()
    zf = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).backward(ref)
    _report("dgrad_dgelu", out, zf.grad, atol=3e-2, rtol=1e-2)
    # saved-derivative variant: out = acc * aux
    L.gemm(dy, w, m=tokens, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16_MULAUX, aux=z, out=out)
    torch.cuda.synchronize()
    _report("dgrad_mulaux", out, ref * z.float(), atol=3e-2, rtol=1e-2)
    # fused column sums of the bf16 output (= fc1's bias gradient when out = dz)
    cs = torch.zeros(k_in, device=DEV, dtype=torch.float32)
    out2 = torch.empty_like(out)
    L.gemm(dy, w, m=tokens, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16_MULAUX, aux=z, out=out2, out_colsum=cs)
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    _report("dgrad_mulaux_colsum", cs, out.float().sum(0), atol=2e-3, rtol=1e-4)


def test_gemm_dgrad_rowdot_is_attention_delta():
    """ROWDOT epilogue: dx = dy W stored as bf16 AND per-(row, 64-column group) dot products of the stored dx with a second
    operand, laid out [sample, group, row-in-sample] — the attention backward's delta when the operand is O."""
    L = _lib()
    batch, seq, heads = 3, 197, 12
    tokens, e = batch * seq, heads * 64
    dy = _rand(tokens, e, seed=1).bfloat16()
    w = _rand(e, e, seed=2, scale=0.05).bfloat16()
    o = _rand(tokens, e, seed=3).bfloat16()
    dx = torch.empty(tokens, e, device=DEV, dtype=torch.bfloat16)
    delta = torch.full((batch, heads, seq), float("nan"), device=DEV, dtype=torch.float32)
    L.gemm(dy, w, m=tokens, n=e, k=e, b_layout=1, epilogue=L.EPI_BF16_ROWDOT, aux=o, out=dx, sumsq=delta, rows_per_sample=seq, cols_per_group=64, n_groups=heads)
    torch.cuda.synchronize()
    ref = dy.float() @ w.float()
    _report("rowdot.dx", dx, ref, atol=3e-2, rtol=1e-2)
    # delta is defined on the STORED (bf16) dx, like the stand-alone delta kernel that reads dO and O from memory
    dref = (dx.float() * o.float()).view(batch, seq, heads, 64).sum(-1).permute(0, 2, 1)
    _report("rowdot.delta", delta.reshape(batch * heads, seq), dref.reshape(batch * heads, seq), atol=2e-3, rtol=1e-4)


@pytest.mark.parametrize("tokens,n_out,k_in,split_k", [(256, 128, 256, 1), (1000, 768, 768, 3), (197 * 8, 2304, 768, 4), (333, 136, 200, 2), (197 * 16, 768, 3072, 0)])
def test_gemm_wgrad_mn_major_both_splitk(tokens, n_out, k_in, split_k):
    L = _lib()
    dy = _rand(tokens, n_out, seed=1).bfloat16()
    x = _rand(tokens, k_in, seed=2).bfloat16()
    dw = torch.zeros(n_out, k_in, device=DEV, dtype=torch.float32)
    L.gemm(dy, x, m=n_out, n=k_in, k=tokens, a_layout=1, b_layout=1, epilogue=L.EPI_F32_ADD, out=dw, split_k=split_k)
    torch.cuda.synchronize()
    ref = dy.float().T @ x.float()
    _report("wgrad", dw, ref, atol=2e-3, rtol=1e-4)
    # accumulate semantics: a second call doubles the result
    L.gemm(dy, x, m=n_out, n=k_in, k=tokens, a_layout=1, b_layout=1, epilogue=L.EPI_F32_ADD, out=dw, split_k=split_k)
    torch.cuda.synchronize()
    _report("wgrad_acc", dw, 2 * ref, atol=4e-3, rtol=1e-4)


def test_gemm_sumsq_per_sample_group():
    L = _lib()
    samples, rows, k, n, cpg = 6, 197, 768, 1536, 768
    m = samples * rows
    a = _rand(m, k, seed=1).bfloat16()
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    ss = torch.zeros(samples, n // cpg, device=DEV, dtype=torch.float32)
    L.gemm(a, b, m=m, n=n, k=k, epilogue=L.EPI_SUMSQ, sumsq=ss, rows_per_sample=rows, cols_per_group=cpg, n_groups=n // cpg)
    torch.cuda.synchronize()
    c = (a.float() @ b.float().T).reshape(samples, rows, n // cpg, cpg)
    ref = (c**2).sum(dim=(1, 3))
    _report("sumsq", ss, ref, atol=1e-2, rtol=1e-4)


def test_gemm_strided_a_view():
    L = _lib()
    m, n, k = 256, 256, 128
    big = _rand(m, 3 * k, seed=1).bfloat16()
    a = big[:, k : 2 * k]
    b = _rand(n, k, seed=2, scale=0.05).bfloat16()
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    L.gemm(a, b, m=m, n=n, k=k, out=out)
    torch.cuda.synchronize()
    _report("gemm_strided", out, a.float() @ b.float().T, atol=2e-2, rtol=1e-2)


# ---------------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(197 * 4, 768), (100, 1024), (37, 128), (64, 1280)])
def test_layernorm_fwd_bwd(rows, cols):
    L = _lib()
    eps = 1e-12
    x = (_rand(rows, cols, seed=1) * 2 + 0.5).bfloat16()
    g = _rand(cols, seed=2) * 0.5 + 1.0
    b = _rand(cols, seed=3) * 0.1
    y, mean, rstd = L.layernorm_fwd(x, g, b, eps)
    torch.cuda.synchronize()
    xf = x.float().requires_grad_(True)
    gf = g.clone().requires_grad_(True)
    bf = b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (cols,), gf, bf, eps)
    _report("ln_fwd", y, ref, atol=2e-2, rtol=1e-2)
    _report("ln_mean", mean[:, None], xf.mean(1, keepdim=True), atol=1e-5, rtol=1e-5)
    dy = _rand(rows, cols, seed=4).bfloat16()
    dres = _rand(rows, cols, seed=5).bfloat16()
    ref.backward(dy.float())
    dg = torch.zeros(cols, device=DEV)
    db = torch.zeros(cols, device=DEV)
    dx = L.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, dgamma=dg, dbeta=db)
    torch.cuda.synchronize()
    _report("ln_dx", dx, xf.grad + dres.float(), atol=3e-2, rtol=1e-2)
    _report("ln_dgamma", dg[None], gf.grad[None], atol=1e-2, rtol=1e-3)
    _report("ln_dbeta", db[None], bf.grad[None], atol=1e-2, rtol=1e-3)
    # the same call also sums the columns of dres (the bias gradient of the Linear that wrote the residual stream): accumulates
    # into its buffer, leaves every other output bit-identical, also with the parameter gradients switched off (frozen norm)
    cs = torch.full((cols,), 0.5, device=DEV)
    dg2, db2 = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
    dx2 = L.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, dgamma=dg2, dbeta=db2, dres_colsum=cs)
    cs_frozen = torch.zeros(cols, device=DEV)
    dx3 = L.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, dres_colsum=cs_frozen)
    torch.cuda.synchronize()
    assert torch.equal(dx2, dx) and torch.equal(dx3, dx)
    _report("ln_dres_colsum", cs[None], (dres.float().sum(0) + 0.5)[None], atol=1e-3, rtol=1e-4)
    _report("ln_dres_colsum_frozen", cs_frozen[None], dres.float().sum(0)[None], atol=1e-3, rtol=1e-4)
    _report("ln_dgamma2", dg2[None], gf.grad[None], atol=1e-2, rtol=1e-3)


def test_layernorm_bwd_colsum_many_rows():
    """More rows than one pass of the persistent grid (every warp accumulates several rows into its shared-memory slice)."""
    L = _lib()
    rows, cols = 197 * 64, 768
    x = _rand(rows, cols, seed=1).bfloat16()
    g = _rand(cols, seed=2) * 0.5 + 1.0
    _, mean, rstd = L.layernorm_fwd(x, g, torch.zeros(cols, device=DEV), 1e-12)
    dy, dres = _rand(rows, cols, seed=4).bfloat16(), _rand(rows, cols, seed=5).bfloat16()
    cs = torch.zeros(cols, device=DEV)
    L.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, dres_colsum=cs)
    torch.cuda.synchronize()
    _report("ln_dres_colsum_rows", cs[None], dres.double().sum(0).float()[None], atol=2e-2, rtol=1e-4)


def test_layernorm_constant_row_is_finite():
    L = _lib()
    x = torch.full((8, 768), 3.0, device=DEV, dtype=torch.bfloat16)
    g = torch.ones(768, device=DEV)
    b = torch.full((768,), 0.25, device=DEV)
    y, _, rstd = L.layernorm_fwd(x, g, b, 1e-12)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all() and torch.isfinite(rstd).all()
    assert torch.allclose(y.float(), torch.full_like(y.float(), 0.25))


# ---------------------------------------------------------------------------------------------------
# Attention
# ---------------------------------------------------------------------------------------------------
def _attn_ref(qkv, batch, seq, heads, hd):
    e = heads * hd
    q, k, v = qkv.float().reshape(batch, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(batch * seq, e)
    return o, torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("batch,seq,heads", [(2, 197, 12), (3, 5, 2), (1, 17, 4), (2, 208, 3), (1, 257, 2)])
def test_attention_fwd_bwd(batch, seq, heads):
    L = _lib()
    hd = 64
    e = heads * hd
    qkv = _rand(batch * seq, 3 * e, seed=1).bfloat16()
    out, lse = L.attention_fwd(qkv, batch, seq, heads, hd)
    torch.cuda.synchronize()
    qf = qkv.float().requires_grad_(True)
    oref, lref = _attn_ref(qf, batch, seq, heads, hd)
    _report("attn_out", out, oref, atol=2e-2, rtol=2e-2)
    _report("attn_lse", lse.reshape(-1, seq), lref.reshape(-1, seq), atol=2e-3, rtol=1e-3)
    dout = _rand(batch * seq, e, seed=2).bfloat16()
    oref.backward(dout.float())
    dqkv = L.attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd)
    torch.cuda.synchronize()
    _report("attn_dqkv", dqkv, qf.grad, atol=3e-2, rtol=3e-2)
    # fused qkv-bias gradient: column sums of dqkv from the same kernel (or the column-sum pass on the long-sequence path)
    dbias = torch.zeros(3 * e, device=DEV, dtype=torch.float32)
    dqkv2 = L.attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd, dbias=dbias)
    torch.cuda.synchronize()
    _report("attn_dqkv_with_bias", dqkv2, qf.grad, atol=3e-2, rtol=3e-2)
    _report("attn_dbias", dbias, qf.grad.sum(0), atol=5e-2, rtol=2e-2)


def test_attention_bwd_with_precomputed_delta_matches():
    """vb_attention_bwd_with_delta (delta supplied by the caller) == vb_attention_bwd (own delta pass), bit for bit."""
    L = _lib()
    batch, seq, heads = 4, 197, 12
    qkv = _rand(batch * seq, 3 * heads * 64, seed=1).bfloat16()
    out, lse = L.attention_fwd(qkv, batch, seq, heads, 64)
    do = _rand(batch * seq, heads * 64, seed=2).bfloat16()
    ref = L.attention_bwd(qkv, out, do, lse, batch, seq, heads, 64)
    delta = (do.float() * out.float()).view(batch, seq, heads, 64).sum(-1).permute(0, 2, 1).contiguous()
    got = L.attention_bwd(qkv, None, do, lse, batch, seq, heads, 64, delta=delta)
    torch.cuda.synchronize()
    _report("attn_bwd_with_delta", got, ref, atol=2e-3, rtol=2e-2)


@pytest.mark.parametrize("batch,seq,heads", [(4, 197, 12), (3, 50, 2), (2, 208, 16)])
def test_attention_bwd_query_only_bias(batch, seq, heads):
    """vb_attention_bwd_with_delta_qbias: same dqkv bits as the all-thirds entry; the query third of the bias gradient is the
    column sum of dQ; the key and value thirds of the buffer are left alone. The identities the caller relies on for those
    two hold on the kernel's own output: column sums of dK ~ 0 and column sums of dV ~ column sums of dO (against the
    magnitude of the query third)."""
    L = _lib()
    e = heads * 64
    qkv = _rand(batch * seq, 3 * e, seed=1).bfloat16()
    out, lse = L.attention_fwd(qkv, batch, seq, heads, 64)
    do = _rand(batch * seq, e, seed=2).bfloat16()
    delta = (do.float() * out.float()).view(batch, seq, heads, 64).sum(-1).permute(0, 2, 1).contiguous()
    db_all = torch.zeros(3 * e, device=DEV)
    ref = L.attention_bwd(qkv, None, do, lse, batch, seq, heads, 64, dbias=db_all, delta=delta)
    db_q = torch.full((3 * e,), 7.0, device=DEV)
    got = L.attention_bwd(qkv, None, do, lse, batch, seq, heads, 64, dbias=db_q, delta=delta, q_bias_only=True)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    assert torch.equal(db_q[e:], torch.full((2 * e,), 7.0, device=DEV))
    _report("attn_qbias", (db_q[:e] - 7.0)[None], db_all[None, :e], atol=2e-3 * float(db_all[:e].abs().max()), rtol=1e-4)
    scale = float(db_all[:e].abs().max())
    assert float(db_all[e:2 * e].abs().max()) <= 2e-2 * scale, "column sums of dK should vanish"
    _report("attn_vbias_identity", db_all[None, 2 * e:], do.float().sum(0)[None], atol=2e-2 * scale, rtol=2e-2)


def test_attention_pair_delta():
    L = _lib()
    batch, seq, heads, hd = 2, 197, 12, 64
    e = heads * hd
    qa = _rand(batch * seq, 3 * e, seed=1).bfloat16()
    qb = _rand(batch * seq, 3 * e, seed=2).bfloat16()
    delta = torch.empty(batch * seq, e, device=DEV, dtype=torch.bfloat16)
    L.attention_pair_delta(qa, qb, delta, batch, seq, heads, hd)
    torch.cuda.synchronize()
    ref = _attn_ref(qa, batch, seq, heads, hd)[0] - _attn_ref(qb, batch, seq, heads, hd)[0]
    _report("attn_pair", delta, ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("layers,batch,seq,heads", [(3, 5, 197, 12), (2, 3, 5, 2), (1, 150, 197, 4)])
def test_attention_pair_delta_all_layers(layers, batch, seq, heads):
    """One launch over every layer's slice of the concatenated qkv projections (the estimator's layout)."""
    L = _lib()
    hd = 64
    e = heads * hd
    qa = _rand(batch * seq, layers * 3 * e, seed=1).bfloat16()
    qb = _rand(batch * seq, layers * 3 * e, seed=2).bfloat16()
    delta = L.attention_pair_delta_layers(qa, qb, layers, batch, seq, heads, hd)
    torch.cuda.synchronize()
    assert delta.shape == (layers, batch * seq, e)
    for i in range(layers):
        sl = slice(i * 3 * e, (i + 1) * 3 * e)
        ref = _attn_ref(qa[:, sl], batch, seq, heads, hd)[0] - _attn_ref(qb[:, sl], batch, seq, heads, hd)[0]
        _report(f"attn_pair_layers[{i}]", delta[i], ref, atol=2e-2, rtol=2e-2)
    # identical inputs cancel exactly (fp32 subtraction of identical accumulators)
    zero = L.attention_pair_delta_layers(qa, qa, layers, batch, seq, heads, hd)
    torch.cuda.synchronize()
    assert float(zero.float().abs().max()) == 0.0


def _attn_core64(qkv, batch, seq, heads, hd):
    """fp64 attention core of a [batch*seq, 3E] projection tensor -> [batch*seq, E]"""
    e = heads * hd
    q, k, v = (t.reshape(batch, seq, heads, hd).transpose(1, 2) for t in qkv.double().split(e, dim=1))
    o = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1) @ v
    return o.transpose(1, 2).reshape(batch * seq, e)


@pytest.mark.parametrize("layers,batch,seq,heads", [(3, 5, 197, 12), (2, 3, 5, 2), (1, 150, 197, 4), (2, 2, 208, 3), (1, 4, 130, 2)])
@pytest.mark.parametrize("scale", [10.0, 1.0, 1e-2, 1e-4])
def test_attention_perturb_delta_all_layers(layers, batch, seq, heads, scale):
    """Perturbation-form paired attention: attn(a + d) - attn(a) from (qkv_a, dqkv), against fp64 on the SAME bf16
    operands. Stated tolerance: relative L2 error of every layer's [tokens, E] difference <= 1e-2 at every
    perturbation size (bf16 roundings of p, g and the output only; no cancellation), exactly 0 for d = 0."""
    L = _lib()
    hd = 64
    e = heads * hd
    qa = _rand(batch * seq, layers * 3 * e, seed=1).bfloat16()
    dq = (_rand(batch * seq, layers * 3 * e, seed=2) * scale).bfloat16()
    delta = L.attention_perturb_delta_layers(qa, dq, layers, batch, seq, heads, hd)
    torch.cuda.synchronize()
    assert delta.shape == (layers, batch * seq, e)
    assert torch.isfinite(delta.float()).all()
    for i in range(layers):
        sl = slice(i * 3 * e, (i + 1) * 3 * e)
        ref = _attn_core64(qa[:, sl].double() + dq[:, sl].double(), batch, seq, heads, hd) - _attn_core64(qa[:, sl], batch, seq, heads, hd)
        err = float((delta[i].double() - ref).norm() / ref.norm())
        assert err <= 1e-2, f"layer {i} scale {scale}: rel L2 {err:.3e}"
        # row-wise too: no token may be off by more than 3 % of its own difference norm (bf16 output rounding is 0.4 %),
        # including scale 10 where the perturbed softmax is one-hot on a key that had negligible weight before
        rerr = (delta[i].double() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30)
        row_tol = 3e-2
        assert float(rerr.max()) <= row_tol, f"layer {i} scale {scale}: worst row rel {float(rerr.max()):.3e} at row {int(rerr.argmax())}"
    zero = L.attention_perturb_delta_layers(qa, torch.zeros_like(dq), layers, batch, seq, heads, hd)
    torch.cuda.synchronize()
    assert float(zero.float().abs().max()) == 0.0


def test_attention_perturb_delta_large_score_shift_is_finite():
    """A difference whose scores exceed the fp32 exp range (|dS| / 8 >> 88) must not produce inf / nan: the row max of dS
    is removed before the exponential."""
    L = _lib()
    layers, batch, seq, heads, hd = 1, 2, 197, 2, 64
    e = heads * hd
    qa = _rand(batch * seq, 3 * e, seed=3).bfloat16()
    dq = (_rand(batch * seq, 3 * e, seed=4) * 40.0).bfloat16()
    delta = L.attention_perturb_delta_layers(qa, dq, layers, batch, seq, heads, hd)
    torch.cuda.synchronize()
    assert torch.isfinite(delta.float()).all()
    ref = _attn_core64(qa.double() + dq.double(), batch, seq, heads, hd) - _attn_core64(qa, batch, seq, heads, hd)
    assert float((delta[0].double() - ref).norm() / ref.norm()) <= 2e-2


# ---------------------------------------------------------------------------------------------------
# Element-wise helpers
# ---------------------------------------------------------------------------------------------------
def test_casts_roundtrip():
    L = _lib()
    x = _rand(1000003 // 1, seed=1)[: 1000000 - 3].contiguous()
    xb = L.cast_f32_to_bf16(x)
    assert torch.equal(xb, x.bfloat16())
    assert torch.equal(L.cast_bf16_to_f32(xb), xb.float())


def test_im2col_matches_conv_unfold():
    L = _lib()
    n, c, h, w, p = 3, 3, 64, 48, 16
    img = _rand(n, c, h, w, seed=1)
    img2 = _rand(n, c, h, w, seed=2)
    ref = torch.nn.functional.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(-1, c * p * p)
    assert torch.equal(L.im2col_patches(img, p), ref.bfloat16())
    ref2 = torch.nn.functional.unfold(img - img2, kernel_size=p, stride=p).transpose(1, 2).reshape(-1, c * p * p)
    assert torch.equal(L.im2col_patches(img, p, img2), ref2.bfloat16())


def test_assemble_tokens_fwd_bwd():
    L = _lib()
    batch, np_, e = 5, 196, 768
    po = _rand(batch * np_, e, seed=1).bfloat16()
    cls = _rand(e, seed=2)
    pos = _rand(np_ + 1, e, seed=3)
    tok, tok32 = L.assemble_tokens(po, None, cls, pos, batch, np_, e, want_f32=True)
    ref = torch.cat([cls.expand(batch, 1, e), po.float().reshape(batch, np_, e)], 1) + pos
    assert torch.allclose(tok32.reshape(batch, np_ + 1, e), ref, atol=1e-6)
    assert torch.equal(tok, ref.reshape(-1, e).bfloat16())
    dtok = _rand(batch * (np_ + 1), e, seed=4).bfloat16()
    dcls = torch.zeros(e, device=DEV)
    dpos = torch.zeros(np_ + 1, e, device=DEV)
    dpatch = L.assemble_tokens_bwd(dtok, dcls, dpos, batch, np_, e)
    d3 = dtok.float().reshape(batch, np_ + 1, e)
    assert torch.equal(dpatch.reshape(batch, np_, e), dtok.reshape(batch, np_ + 1, e)[:, 1:])
    assert torch.allclose(dpos, d3.sum(0), atol=1e-4)
    assert torch.allclose(dcls, d3[:, 0].sum(0), atol=1e-4)


def test_colsum_and_add():
    L = _lib()
    x = _rand(1234, 2304, seed=1).bfloat16()
    out = torch.ones(2304, device=DEV)
    L.colsum_bf16(x, out)
    assert torch.allclose(out, 1 + x.float().sum(0), atol=2e-2, rtol=1e-4)
    y = _rand(1234, 2304, seed=2).bfloat16()
    assert torch.equal(L.add_bf16(x, y), (x.float() + y.float()).bfloat16())


def test_rowsumsq_and_ln_pair():
    L = _lib()
    s, rows, cols = 4, 197, 768
    a = _rand(s * rows, cols, seed=1)
    b = _rand(s * rows, cols, seed=2)
    out = torch.zeros(s, device=DEV)
    L.rowsumsq_diff_f32(a, b, out, s, rows, cols)
    ref = ((a - b) ** 2).reshape(s, -1).sum(1)
    assert torch.allclose(out, ref, rtol=1e-5)
    u = torch.zeros(s, cols, device=DEV)
    L.layernorm_pair_sqdiff(a, b, u, s, rows, cols, 1e-12)
    za = torch.nn.functional.layer_norm(a, (cols,), eps=1e-12)
    zb = torch.nn.functional.layer_norm(b, (cols,), eps=1e-12)
    refu = ((za - zb) ** 2).reshape(s, rows, cols).sum(1)
    assert torch.allclose(u, refu, rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------------
# fused optimizers over the flat gradient arena
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_fused_arena_optimizers_match_torch(opt_name):
    """clip_grad_norm_(1.0) + torch.optim.{SGD(momentum 0.9), AdamW} against FusedSGD / FusedAdamW (one norm launch + one
    update launch over the arena) on identical parameters and gradients, four steps; fp32 round-off tolerance."""
    from vit_plasticity_b200.finetune import FusedAdamW, FusedSGD

    shapes = [(257, 64), (64,), (1000, 33), (7,), (128, 128)]
    g = torch.Generator().manual_seed(3)
    ref = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    got = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    if opt_name == "sgd":
        o_ref, o_got = torch.optim.SGD(ref, lr=1e-2, momentum=0.9, weight_decay=1e-3), FusedSGD(got, lr=1e-2, momentum=0.9, weight_decay=1e-3)
    else:
        o_ref, o_got = torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-2), FusedAdamW(got, lr=1e-3, weight_decay=1e-2)
    for step in range(4):
        grads = [torch.randn(*s, generator=g).to(DEV) * (0.1 + step) for s in shapes]
        for p, q, gr in zip(ref, got, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr)  # FusedSGD keeps .grad as views of its arena
        n_ref = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
        n_got = o_got.step(max_norm=1.0)
        o_got.zero_grad()
        torch.cuda.synchronize()
        assert abs(float(n_ref) - float(n_got)) <= 1e-5 * float(n_ref)
        for p, q in zip(ref, got):
            _report(f"{opt_name} step {step}", q.detach().reshape(1, -1), p.detach().reshape(1, -1), atol=2e-6, rtol=2e-5)


@pytest.mark.parametrize("scale", [1.0, 1e-3, 1e-6, 30.0])
@pytest.mark.parametrize("cols", [128, 768, 1024])
def test_layernorm_delta_sqdiff(scale, cols):
    """u[s, c] = sum_rows (zhat(a + scale d) - zhat(a))^2 in perturbation form against fp64; stated tolerance 2e-5
    relative at every scale (two fp32 LayerNorms subtracted lose ~1e-7 / scale), exactly 0 for d = 0."""
    L = _lib()
    s, rows = 3, 197
    a = _rand(s * rows, cols, seed=1)
    d = _rand(s * rows, cols, seed=2)
    u = torch.zeros(s, cols, device=DEV)
    L.layernorm_delta_sqdiff(a, d, scale, u, s, rows, cols, 1e-12)
    za = torch.nn.functional.layer_norm(a.double(), (cols,), eps=1e-12)
    zb = torch.nn.functional.layer_norm(a.double() + scale * d.double(), (cols,), eps=1e-12)
    ref = ((zb - za) ** 2).reshape(s, rows, cols).sum(1)
    assert float(((u.double() - ref).abs() / ref).max()) <= 2e-5
    u0 = torch.zeros(s, cols, device=DEV)
    L.layernorm_delta_sqdiff(a, torch.zeros_like(d), 1.0, u0, s, rows, cols, 1e-12)
    assert float(u0.abs().max()) == 0.0


def test_scale_bf16():
    L = _lib()
    x = _rand(1000, 1237, seed=5).bfloat16().reshape(-1)[: 1000 * 1237 - 3].contiguous()  # ragged tail
    y = L.scale_bf16(x, 1e-3)
    assert torch.equal(y, (x.float() * 1e-3).bfloat16())


def test_philox_normal_matches_oracle_and_is_batching_invariant():
    """Device noise generator of the sweep against oracle/philox_oracle.py (Philox4x32-10 pinned on Random123's known
    answers; Box-Muller in fp64 there, fp32 here: abs tolerance 2e-5), and image i's noise is the same whatever batch or
    shard it is drawn in."""
    from oracle.philox_oracle import normal_images

    L = _lib()
    shape, seed = (3, 32, 32), 5
    got = L.philox_normal(6, shape, seed, 7, DEV)
    ref = torch.from_numpy(normal_images(6, 3 * 32 * 32, seed, 7)).reshape(6, *shape)
    assert float((got.cpu() - ref).abs().max()) <= 2e-5
    again = torch.cat([L.philox_normal(2, shape, seed, 7, DEV), L.philox_normal(4, shape, seed, 9, DEV)])
    assert torch.equal(got, again)
    big = L.philox_normal(64, (3, 224, 224), 1, 2**33, DEV)  # image indices beyond 32 bits reach the counter's high word
    assert abs(float(big.mean())) < 2e-3 and abs(float(big.std()) - 1.0) < 2e-3 and torch.isfinite(big).all()
    assert not torch.equal(big[0], L.philox_normal(1, (3, 224, 224), 1, 0, DEV)[0])
