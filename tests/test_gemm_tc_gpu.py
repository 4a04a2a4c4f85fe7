"""tcgen05 TF32 / 3xTF32 GEMM (hg_gemm_tf32) against an fp64 torch reference, all operand layouts used by the MLP."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(A, B, M, N, K, a_mn, b_mn, passes, epilogue=0, bias=None, H=None, split_k=1, trust=0, C=None, B_lo=None):
    from humanoid import _native as nat
    if C is None:
        C = torch.zeros(M, N, device="cuda")
    d = nat.Gemm()
    d.A, d.B, d.C = A.data_ptr(), B.data_ptr(), C.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.H = H.data_ptr() if H is not None else None
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb, d.ldc = A.stride(0), B.stride(0), C.stride(0)
    d.ldh = H.stride(0) if H is not None else 0
    d.B_lo = B_lo.data_ptr() if B_lo is not None else None
    d.a_mn_major, d.b_mn_major, d.epilogue, d.passes, d.split_k, d.trust_hw_truncation = a_mn, b_mn, epilogue, passes, split_k, trust
    nat.check(nat.lib.hg_gemm_tf32(d, torch.cuda.current_stream().cuda_stream), "hg_gemm_tf32")
    torch.cuda.synchronize()
    return C


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _pad4(t):
    """(R, C) tensor re-laid out with a row pitch that is a multiple of 4 floats (TMA requirement)."""
    R, Cc = t.shape
    ld = (Cc + 3) // 4 * 4
    buf = torch.zeros(R, ld, device=t.device)
    buf[:, :Cc] = t
    return buf[:, :Cc]


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (384, 512, 705), (300, 12, 128), (4096, 768, 219), (128, 256, 512)])
@pytest.mark.parametrize("passes", [1, 3])
def test_forward_layout_k_major(M, N, K, passes):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X = _pad4(torch.randn(M, K, device="cuda", generator=g))
    W = _pad4(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    b = torch.randn(N, device="cuda", generator=g)
    ref = X.double() @ W.double().t()
    C = _run(X, W, M, N, K, 0, 0, passes)
    tol = 1e-5 if passes == 3 else 2e-3
    assert _rel(C, ref) < tol, _rel(C, ref)
    C2 = _run(X, W, M, N, K, 0, 0, passes, epilogue=2, bias=b)
    ref2 = torch.nn.functional.elu(ref + b.double())
    assert _rel(C2, ref2) < tol * 2, _rel(C2, ref2)


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (512, 512, 256), (384, 768, 256), (200, 128, 12)])
def test_dgrad_layout(M, N, K):
    """dX (M x N) = dZ (M x K) W (K x N): A K-major, B MN-major; epilogue multiplies by ELU'."""
    g = torch.Generator(device="cuda").manual_seed(7)
    dZ = _pad4(torch.randn(M, K, device="cuda", generator=g))
    W = _pad4(torch.randn(K, N, device="cuda", generator=g))
    Hh = torch.nn.functional.elu(torch.randn(M, N, device="cuda", generator=g))
    ref = (dZ.double() @ W.double()) * torch.where(Hh > 0, torch.ones_like(Hh), Hh + 1).double()
    C = _run(dZ, W, M, N, K, 0, 1, 3, epilogue=3, H=Hh)
    assert _rel(C, ref) < 1e-5, _rel(C, ref)


@pytest.mark.parametrize("Nout,Kin,batch,split", [(128, 128, 256, 1), (512, 705, 4096, 4), (12, 128, 2048, 3), (768, 219, 1024, 2),
                                                  (256, 512, 61440, 16)])
def test_wgrad_layout(Nout, Kin, batch, split):
    """dW (Nout x Kin) = dZ^T X: both operands MN-major, split-K with atomics."""
    g = torch.Generator(device="cuda").manual_seed(9)
    dZ = _pad4(torch.randn(batch, Nout, device="cuda", generator=g))
    X = _pad4(torch.randn(batch, Kin, device="cuda", generator=g))
    ref = dZ.double().t() @ X.double()
    C = _run(dZ, X, Nout, Kin, batch, 1, 1, 3, epilogue=4, split_k=split)
    # fp32 accumulation over `batch` terms: the error floor grows like sqrt(batch) * 2^-24
    assert _rel(C, ref) < max(1e-5, 2e-7 * batch ** 0.5), _rel(C, ref)


def test_hardware_truncates_tf32_operands():
    """Does the tensor core ignore the low 13 mantissa bits (truncate) of fp32 operands?  If yes, the raw
    tile can serve as the 'hi' operand without an explicit mask (trust_hw_truncation=1)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 256, 128, 256
    X = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    mask = torch.tensor(-8192, dtype=torch.int32, device="cuda")       # 0xFFFFE000
    Xt = (X.view(torch.int32) & mask).view(torch.float32)
    Wt = (W.view(torch.int32) & mask).view(torch.float32)
    raw = _run(X, W, M, N, K, 0, 0, 1)
    trunc = _run(Xt, Wt, M, N, K, 0, 0, 1)
    same = torch.equal(raw, trunc)
    print("tensor core truncates tf32 operands:", same, " max diff", float((raw - trunc).abs().max()))
    a = _run(X, W, M, N, K, 0, 0, 3, trust=1)
    b = _run(X, W, M, N, K, 0, 0, 3, trust=0)
    ref = X.double() @ W.double().t()
    print("3xTF32 rel err trust=1:", _rel(a, ref), " trust=0:", _rel(b, ref))
    assert _rel(b, ref) < 1e-5


def test_alignment_errors():
    from humanoid import _native as nat
    X = torch.randn(64, 705, device="cuda")       # pitch 705 floats: not a multiple of 16 bytes
    W = torch.randn(32, 705, device="cuda")
    with pytest.raises(nat.NativeError, match="16-byte"):
        _run(X, W, 64, 32, 705, 0, 0, 3)


@pytest.mark.parametrize("M,N,K", [(4096, 512, 705), (4096, 256, 512), (4096, 128, 256), (300, 768, 219), (4096, 12, 128)])
def test_presplit_weight_residuals(M, N, K):
    """B_lo = hg_tf32_residual(B) loaded by TMA (the rollout path: splitter handles A only) must reproduce the in-kernel
    split bit for bit -- same residual formula, same MMA order -- for every tile width the skinny-batch heuristic picks."""
    from humanoid import _native as nat
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X = _pad4(torch.randn(M, K, device="cuda", generator=g))
    W = _pad4(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    b = torch.randn(N, device="cuda", generator=g)
    Wbuf = torch.zeros(N, W.stride(0), device="cuda")
    Wbuf[:, :K] = W
    lo = torch.empty_like(Wbuf)
    nat.check(nat.lib.hg_tf32_residual(Wbuf.data_ptr(), lo.data_ptr(), Wbuf.numel(), 0), "hg_tf32_residual")
    Wv = Wbuf[:, :K]
    a = _run(X, Wv, M, N, K, 0, 0, 3, epilogue=2, bias=b, trust=1)
    c = _run(X, Wv, M, N, K, 0, 0, 3, epilogue=2, bias=b, trust=1, B_lo=lo[:, :K])
    assert torch.equal(a, c), float((a - c).abs().max())
    ref = torch.nn.functional.elu(X.double() @ W.double().t() + b.double())
    assert _rel(c, ref) < 2e-5, _rel(c, ref)
