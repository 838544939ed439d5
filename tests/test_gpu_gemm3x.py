"""tcgen05 3xTF32 GEMM kernels (csrc/gemm3x.cu) against fp64 products of the same inputs.

Bars: 2e-6 max-normalised error (plain cuBLAS fp32 reaches ~1e-6 on the same inputs); the layout probes must be
bit-exact (every value is a small integer, exactly representable in TF32 hi/lo parts)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("m,n,k,transpose,bias,relu", [
    (128, 16, 8, False, False, False),          # smallest supported shape, one K step
    (130, 300, 44, False, False, False),        # row tail, K tail inside a chunk
    (777, 304, 128, False, False, False),       # two N halves of 160 + 144
    (1000, 256, 256, False, True, False),       # two halves of 128
    (4096, 300, 300, False, True, True),        # bench width with bias + ReLU epilogue
    (18269, 300, 300, False, False, False),     # bench shape (config #2 node count)
    (18269, 300, 300, True, False, False),      # dX = dY . W (transposed weight image)
    (5, 64, 64, False, True, False),            # fewer rows than one tile
])
def test_gemm3x_matches_fp64(cuda, m, n, k, transpose, bias, relu):
    from graph_hscn_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    a = torch.randn(m, k, device=cuda, generator=g)
    w = torch.randn((k, n) if transpose else (n, k), device=cuda, generator=g) / k ** 0.5
    b = torch.randn(n, device=cuda, generator=g) if bias else None
    assert gemm.gemm3x_supported(m, n, k)
    c = gemm.gemm3x(a, gemm.gemm3x_prep(w, transpose), n, b, relu)
    ref = a.double() @ (w.double() if transpose else w.double().t())
    if bias:
        ref = ref + b.double()
    if relu:
        ref = ref.relu()
    assert c.shape == (m, n)
    assert _rel(c, ref) < 2e-6


@pytest.mark.parametrize("rows,m,n", [(64, 128, 32), (100, 44, 64), (1000, 300, 300), (18269, 300, 300),
                                      (18269, 256, 256), (5000, 300, 48), (16, 4, 16)])
def test_gemm3x_tn_matches_fp64(cuda, rows, m, n):
    from graph_hscn_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(rows + m)
    p = torch.randn(rows, m, device=cuda, generator=g)
    q = torch.randn(rows, n, device=cuda, generator=g)
    assert gemm.gemm3x_tn_supported(rows, m, n)
    out = gemm.gemm3x_tn(p, q)
    assert out.shape == (m, n)
    assert _rel(out, p.double().t() @ q.double()) < 2e-6


def test_gemm3x_same_sign_data_has_no_truncation_bias(cuda):
    """Post-ReLU features are all >= 0: a single TMEM accumulator truncates coherently (3e-6 .. 1e-4 after the weight
    gradient); the three-accumulator scheme must stay at fp32 level (cuBLAS fp32 itself is at ~1e-6 here)."""
    from graph_hscn_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(9000, 300, device=cuda, generator=g) * 5            # all positive
    w = torch.rand(300, 300, device=cuda, generator=g) / 300           # all positive
    dy = torch.rand(9000, 300, device=cuda, generator=g)
    y = gemm.gemm3x(x, gemm.gemm3x_prep(w), 300)
    dw = gemm.gemm3x_tn(dy, x)
    ey, edw = _rel(y, x.double() @ w.double().t()), _rel(dw, dy.double().t() @ x.double())
    assert ey < 2e-6 and edw < 2e-6, f"same-sign inputs: y {ey:.2e}, dW {edw:.2e}"


def test_gemm3x_layout_probes_are_bit_exact(cuda):
    """Unit impulses expose a wrong swizzle / descriptor immediately: A = e_(r,kk) gives C[r, :] = W[:, kk]."""
    from graph_hscn_b200 import gemm
    m, n, k = 128, 32, 32
    w = torch.arange(n, device=cuda).float()[:, None] * 100 + torch.arange(k, device=cuda).float()[None, :]
    img = gemm.gemm3x_prep(w)
    for r, kk in [(0, 0), (1, 0), (0, 1), (0, 4), (0, 8), (5, 13), (9, 31), (77, 20), (127, 7)]:
        a = torch.zeros(m, k, device=cuda)
        a[r, kk] = 1.0
        c = gemm.gemm3x(a, img, n)
        assert torch.equal(c[r], w[:, kk])
        c[r] = 0
        assert not bool(c.any())
    rows, mm_, nn_ = 64, 128, 32
    q = torch.arange(rows, device=cuda).float()[:, None] * 100 + torch.arange(nn_, device=cuda).float()[None, :]
    for r, mm in [(0, 0), (1, 0), (0, 1), (3, 37), (9, 5), (17, 100), (40, 127)]:
        p = torch.zeros(rows, mm_, device=cuda)
        p[r, mm] = 1.0
        out = gemm.gemm3x_tn(p, q)
        assert torch.equal(out[mm], q[r])
        out[mm] = 0
        assert not bool(out.any())


def test_linear_uses_the_tcgen05_kernels_and_falls_back_for_unsupported_shapes(cuda):
    from graph_hscn_b200 import gemm
    from graph_hscn_b200._lib import lib
    assert gemm.USE_TCGEN05
    x = torch.randn(5000, 300, device=cuda, requires_grad=True)
    w = torch.randn(300, 300, device=cuda, requires_grad=True)
    before = lib().launches
    y = gemm.linear(x, w, None)
    y.sum().backward()
    # prep + gemm3x (fwd), prep + gemm3x (dX), gemm3x_tn (dW): five C-ABI calls, no split_cat
    assert lib().launches - before == 5
    # 151 tiles = one wave of 148 + 3: the trailing rows go to a plain fp32 GEMM, results stay at fp32 level
    assert gemm.wave_rows(19300) == 148 * 128 and gemm.wave_rows(18269) == 18269 and gemm.wave_rows(150000) == 150000
    xb = torch.randn(19300, 300, device=cuda, requires_grad=True)
    yb = gemm.linear(xb, w, None)
    gy = torch.randn_like(yb)
    dxb, dwb = torch.autograd.grad(yb, (xb, w), gy)
    assert _rel(yb, xb.detach().double() @ w.detach().double().t()) < 2e-6
    assert _rel(dxb, gy.double() @ w.detach().double()) < 2e-6
    assert _rel(dwb, gy.double().t() @ xb.detach().double()) < 2e-6
    wide = torch.randn(512, 300, device=cuda)                        # n_out = 512 > 320: library 3xTF32 path
    assert not gemm.gemm3x_supported(5000, 512, 300)
    ref = x.detach().double() @ wide.double().t()
    assert _rel(gemm.linear(x.detach(), wide, None), ref) < 2e-6
