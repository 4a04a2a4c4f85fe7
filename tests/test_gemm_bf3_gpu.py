"""tcgen05 bf16x3 GEMM on pre-split operands (hg_gemm_bf16x3) against an fp64 torch reference: every operand layout and
epilogue the split-precision MLP path uses, ragged extents included.  Two error budgets are checked separately:
  * exactness of the kernel on the operands it is given (reference = fp64 product of the SAME hi + lo values): < 5e-6,
    i.e. only the dropped lo*lo term (2^-18 relative per element at worst) and fp32 accumulation;
  * accuracy against the original fp32 operands (what the gradient bar sees): ~5e-6 per product."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(x, ld=None):
    """(R, C) fp32 -> (2, R, ld) int16 planes through the library's own splitter + the values they represent."""
    from humanoid import _native as nat
    R, Cc = x.shape
    ld = ld or (Cc + 7) // 8 * 8
    out = torch.zeros(2, R, ld, dtype=torch.int16, device="cuda")
    xc = x.contiguous()
    nat.check(nat.lib.hg_split_bf16(xc.data_ptr(), xc.stride(0), nat.Split.of(out), R, Cc, 0), "hg_split_bf16")
    val = torch.empty(R, Cc, device="cuda")
    nat.check(nat.lib.hg_unsplit_bf16(nat.Split.of(out), val.data_ptr(), Cc, R, Cc, 0), "hg_unsplit_bf16")
    torch.cuda.synchronize()
    return out, val


def _unsplit(planes, cols):
    from humanoid import _native as nat
    R = planes.shape[1]
    val = torch.empty(R, cols, device="cuda")
    nat.check(nat.lib.hg_unsplit_bf16(nat.Split.of(planes), val.data_ptr(), cols, R, cols, 0), "hg_unsplit_bf16")
    torch.cuda.synchronize()
    return val


def _gemm(As, Bs, M, N, K, a_mn, b_mn, epilogue, bias=None, Hs=None, split_k=1, colsum=None):
    from humanoid import _native as nat
    d = nat.GemmSplit()
    d.A, d.B = nat.Split.of(As), nat.Split.of(Bs)
    C = Cs = None
    if epilogue in (0, 1, 4):
        C = torch.zeros(M, N, device="cuda")
        d.C, d.ldc = C.data_ptr(), N
    else:
        Cs = torch.zeros(2, M, (N + 7) // 8 * 8, dtype=torch.int16, device="cuda")
        d.Cs = nat.Split.of(Cs)
    d.bias = bias.data_ptr() if bias is not None else None
    if Hs is not None:
        d.Hs = nat.Split.of(Hs)
    d.colsum = colsum.data_ptr() if colsum is not None else None
    d.M, d.N, d.K = M, N, K
    d.a_mn_major, d.b_mn_major, d.epilogue, d.split_k = a_mn, b_mn, epilogue, split_k
    nat.check(nat.lib.hg_gemm_bf16x3(d, torch.cuda.current_stream().cuda_stream), "hg_gemm_bf16x3")
    torch.cuda.synchronize()
    return C if C is not None else _unsplit(Cs, N)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_split_roundtrip():
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(300, 705, device="cuda", generator=g) * torch.logspace(-6, 3, 705, device="cuda")
    planes, val = _split(x)
    assert planes.shape == (2, 300, 712)
    rel = ((val - x).abs() / x.abs().clamp_min(1e-30)).max()
    assert float(rel) < 2.0 ** -16, float(rel)             # hi + lo carries >= 16 significant bits
    assert (planes[:, :, 705:] == 0).all()


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (384, 512, 705), (300, 768, 219), (4096, 256, 512), (128, 64, 96), (61440, 128, 256)])
def test_forward_layout_k_major(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    Xs, Xv = _split(X)
    Ws, Wv = _split(W)
    C = _gemm(Xs, Ws, M, N, K, 0, 0, 0)
    assert _rel(C, Xv.double() @ Wv.double().t()) < 5e-6, _rel(C, Xv.double() @ Wv.double().t())
    ref = X.double() @ W.double().t()
    assert _rel(C, ref) < 1e-5, _rel(C, ref)
    C1 = _gemm(Xs, Ws, M, N, K, 0, 0, 1, bias=b)
    assert _rel(C1, ref + b.double()) < 1e-5
    H = _gemm(Xs, Ws, M, N, K, 0, 0, 2, bias=b)              # split store of ELU(acc + b)
    assert _rel(H, torch.nn.functional.elu(ref + b.double())) < 1.5e-5, _rel(H, torch.nn.functional.elu(ref + b.double()))
    P = _gemm(Xs, Ws, M, N, K, 0, 0, 5)                      # plain split store
    assert _rel(P, ref) < 1.5e-5


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (512, 512, 256), (384, 768, 256), (200, 256, 128), (61440, 256, 128)])
def test_dgrad_layout(M, N, K):
    """dX (M x N) = (dZ (M x K) W (K x N)) * ELU'(h): A K-major, B MN-major; split store + column sums (bias gradient)."""
    g = torch.Generator(device="cuda").manual_seed(7)
    dZ = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(K, N, device="cuda", generator=g)
    Hh = torch.nn.functional.elu(torch.randn(M, N, device="cuda", generator=g))
    dZs, _ = _split(dZ)
    Ws, _ = _split(W)
    Hs, Hv = _split(Hh)
    ref = (dZ.double() @ W.double()) * torch.where(Hv > 0, torch.ones_like(Hv), Hv + 1).double()
    cs = torch.zeros(N, device="cuda")
    C = _gemm(dZs, Ws, M, N, K, 0, 1, 3, Hs=Hs, colsum=cs)
    assert _rel(C, ref) < 1.5e-5, _rel(C, ref)
    want = ref.sum(0)
    assert float((cs.double() - want).abs().max()) < 2e-5 * float(ref.abs().sum(0).max()), (cs[:4], want[:4])


@pytest.mark.parametrize("Nout,Kin,batch,split", [(128, 128, 256, 1), (512, 705, 4096, 4), (768, 219, 1024, 2), (256, 512, 61440, 16),
                                                  (64, 40, 300, 2)])
def test_wgrad_layout(Nout, Kin, batch, split):
    """dW (Nout x Kin) = dZ^T X: both operands MN-major, split-K with fp32 atomics; ragged batch / widths."""
    g = torch.Generator(device="cuda").manual_seed(9)
    dZ = torch.randn(batch, Nout, device="cuda", generator=g)
    X = torch.randn(batch, Kin, device="cuda", generator=g)
    dZs, dv = _split(dZ)
    Xs, xv = _split(X)
    C = _gemm(dZs, Xs, Nout, Kin, batch, 1, 1, 4, split_k=split)
    exact = dv.double().t() @ xv.double()
    assert _rel(C, exact) < max(5e-6, 2e-7 * batch ** 0.5), _rel(C, exact)
    ref = dZ.double().t() @ X.double()
    assert _rel(C, ref) < max(1e-5, 2e-7 * batch ** 0.5), _rel(C, ref)


def test_argument_errors():
    from humanoid import _native as nat
    Xs = torch.zeros(2, 64, 705, dtype=torch.int16, device="cuda")       # pitch 705: not a multiple of 8
    Ws = torch.zeros(2, 32, 712, dtype=torch.int16, device="cuda")
    with pytest.raises(nat.NativeError, match="16-byte"):
        _gemm(Xs, Ws, 64, 32, 705, 0, 0, 0)
    Xs = torch.zeros(2, 64, 712, dtype=torch.int16, device="cuda")
    with pytest.raises(nat.NativeError, match="split_k"):
        _gemm(Xs, Ws, 64, 32, 705, 0, 0, 0, split_k=2)


@pytest.mark.parametrize("env", [{"HG_BF3_PAIR": "3"}, {"HG_BF3_PAIR": "2"}, {"HG_BF3_PAIR": "0"}, {"HG_BF3_PAIR": "3", "HG_BF3_TMA_STORE": "0"},
                                 {"HG_BF3_TMA_H": "0"}],
                         ids=["pairs-everywhere", "pairs-cta_group2-tma", "single-cta-only", "direct-stores", "dgrad-h-by-lane-loads"])
def test_kernel_variants(env):
    """The library picks the CTA-pair (cta_group::2) or single-CTA kernel per launch and reads its knobs once per process:
    rerun this module's layout tests in a child process with each form pinned, so every variant sees every layout."""
    import os
    import subprocess
    import sys
    if os.environ.get("HG_BF3_VARIANT_CHILD"):
        pytest.skip("child run")
    child_env = dict(os.environ, HG_BF3_VARIANT_CHILD="1", **env)
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-x", "-q", "-m", "gpu", "-k", "layout or roundtrip", "-p", "no:cacheprovider"],
                       env=child_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
