"""GPU: each kernel through the C-ABI against numpy on the same seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand(rng, *shape):
    return (rng.randn(*shape) + 1j * rng.randn(*shape)).astype(np.complex64)


def _graded(rng, m, n, decay):
    a = rng.randn(m, n) + 1j * rng.randn(m, n)
    u, s, vh = np.linalg.svd(a, full_matrices=False)
    s = s * np.exp(-np.arange(len(s)) / len(s) * decay)
    return ((u * s) @ vh).astype(np.complex64)


def _svd(mats, k, lc):
    import torch
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    mats = np.ascontiguousarray(mats, dtype=np.complex64)
    nj, m, n = mats.shape
    dev = torch.device("cuda")
    x = torch.from_numpy(mats).to(dev)
    left = torch.zeros((nj, m, k), dtype=torch.complex64, device=dev)
    right = torch.zeros((nj, k, n), dtype=torch.complex64, device=dev)
    sv = torch.zeros((nj, min(m, n)), dtype=torch.float32, device=dev)
    info = torch.full((nj, 2), -1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(lib.mpsb_svd_workspace_bytes(nj, m, n), 256), dtype=torch.uint8, device=dev)
    _lib.check(lib.mpsb_svd(x.data_ptr(), nj, m, n, k, int(lc), left.data_ptr(), right.data_ptr(), sv.data_ptr(),
                            info.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mpsb_svd")
    torch.cuda.synchronize()
    return left.cpu().numpy(), right.cpu().numpy(), sv.cpu().numpy(), info.cpu().numpy()


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 64), (130, 67, 33), (256, 128, 512)])
def test_cgemm(M, N, K):
    import torch
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    rng = np.random.RandomState(M * 1000 + N)
    nb = 3
    a, b = _rand(rng, nb, M, K), _rand(rng, nb, K, N)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.zeros((nb, M, N), dtype=torch.complex64, device="cuda")
    _lib.check(lib.mpsb_cgemm(da.data_ptr(), K, 1, 0, M * K, db.data_ptr(), N, 1, 0, K * N, dc.data_ptr(), N, M * N,
                              M, N, K, nb, _lib.stream_ptr()))
    ref = a.astype(np.complex128) @ b.astype(np.complex128)
    np.testing.assert_allclose(dc.cpu().numpy(), ref, atol=2e-5 * np.sqrt(K) * 4)
    # A^H B with strides: A stored [K][M]
    at = np.ascontiguousarray(np.conj(np.transpose(a, (0, 2, 1))))
    dat = torch.from_numpy(at).cuda()
    _lib.check(lib.mpsb_cgemm(dat.data_ptr(), 1, M, 1, M * K, db.data_ptr(), N, 1, 0, K * N, dc.data_ptr(), N, M * N,
                              M, N, K, nb, _lib.stream_ptr()))
    np.testing.assert_allclose(dc.cpu().numpy(), ref, atol=2e-5 * np.sqrt(K) * 4)


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (128, 128, 64), (256, 384, 100), (130, 70, 33), (512, 512, 256)])
def test_cgemm_tc(M, N, K):
    """tcgen05 3xTF32 complex GEMM (tc_gemm.cu) against numpy complex128: fp32-level accuracy."""
    import torch
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    rng = np.random.RandomState(M + 3 * N + 7 * K)
    nb = 3
    a, b = _rand(rng, nb, M, K), _rand(rng, nb, K, N)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.zeros((nb, M, N), dtype=torch.complex64, device="cuda")
    ws = torch.empty(lib.mpsb_cgemm_tc_workspace_bytes(M, N, K, nb) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mpsb_cgemm_tc(da.data_ptr(), M * K, db.data_ptr(), K * N, dc.data_ptr(), N, M * N, M, N, K, nb,
                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = a.astype(np.complex128) @ b.astype(np.complex128)
    err = np.abs(dc.cpu().numpy() - ref).max()
    assert err <= 2e-5 * np.sqrt(K), err          # a small multiple of the fp32 FFMA kernel's error (scripts/tc_accuracy.py)


@pytest.mark.parametrize("chi", [(32, 16, 32), (64, 64, 64), (33, 17, 40), (128, 128, 128), (256, 200, 192)])
def test_theta_tensor_core(chi):
    """theta through the tcgen05 kernel (workspace given) against the einsum, incl. ragged shapes."""
    import torch
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    cl, cm, cr = chi
    d = 2
    rng = np.random.RandomState(cl * 7 + cm * 3 + cr)
    nb = 2
    A, B = _rand(rng, nb, cl, d, cm), _rand(rng, nb, cm, d, cr)
    G = _rand(rng, nb, d, d, d, d)
    dA, dB, dG = (torch.from_numpy(x).cuda() for x in (A, B, G))
    desc = np.zeros(1, dtype=_lib.GATE2_DESC)
    desc[0] = (dA.data_ptr(), dB.data_ptr(), 0, 0, dG.data_ptr(), 0, cl * d * cm, cm * d * cr, 0, 0, d ** 4, 0)
    ddesc = _lib.to_device_bytes(desc, "cuda")
    out = torch.zeros((nb, d * cl, d * cr), dtype=torch.complex64, device="cuda")
    nbytes = lib.mpsb_theta_workspace_bytes(1, nb, d, cl, cm, cr)
    assert nbytes > 0
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mpsb_theta(ddesc.data_ptr(), 1, nb, d, cl, cm, cr, out.data_ptr(), ws.data_ptr(), ws.numel(),
                              _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = np.einsum("bxypq,blpm,bmqr->blxyr", G.astype(np.complex128), A.astype(np.complex128), B.astype(np.complex128))
    ref = ref.reshape(nb, cl * d, d * cr)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=3e-5 * np.sqrt(cm) * 4)


@pytest.mark.parametrize("shape", [(2, 2), (4, 2), (2, 8), (8, 8), (16, 32), (32, 16), (64, 64), (128, 64),
                                   (64, 128), (100, 36), (128, 128)])
@pytest.mark.parametrize("lc", [1, 0])
def test_svd_small_matches_lapack(shape, lc):
    m, n = shape
    rng = np.random.RandomState(m * 131 + n * 7 + lc)
    mats = np.stack([_graded(rng, m, n, 0.0), _graded(rng, m, n, 10.0), _graded(rng, m, n, 30.0)])
    k = min(m, n)
    left, right, sv, info = _svd(mats, k, lc)
    assert (info[:, 0] == 0).all(), info
    assert (info[:, 1] <= 14).all(), info
    for j in range(len(mats)):
        sref = np.linalg.svd(mats[j].astype(np.complex128), compute_uv=False)
        assert np.abs(sv[j] - sref).max() <= 1e-5 * sref[0]            # north_star: 1e-5 relative
        np.testing.assert_allclose(left[j] @ right[j], mats[j], atol=3e-6 * sref[0])
        iso = left[j] if lc else right[j].conj().T
        np.testing.assert_allclose(iso.conj().T @ iso, np.eye(k), atol=2e-5)
    # truncated: exact projection onto the leading singular subspace
    kk = max(1, k // 2)
    left, right, sv, info = _svd(mats, kk, lc)
    for j in range(len(mats)):
        u, s, vh = np.linalg.svd(mats[j].astype(np.complex128), full_matrices=False)
        best = (u[:, :kk] * s[:kk]) @ vh[:kk]
        gap = s[kk - 1] - s[kk] if kk < len(s) else s[kk - 1]
        tol = 1e-5 * s[0] * (1 + s[0] / max(gap, 1e-3 * s[0]) * 0.05)
        np.testing.assert_allclose(left[j] @ right[j], best, atol=10 * tol)


def test_svd_ties_and_rank_deficiency():
    # Bell + maxsvals=1 (README.md:48-53): diagonal with a tie keeps the FIRST
    m = np.diag([2 ** -0.5, 2 ** -0.5]).astype(np.complex64)[None]
    for lc in (1, 0):
        left, right, sv, info = _svd(m, 1, lc)
        np.testing.assert_allclose(left[0] @ right[0], np.diag([2 ** -0.5, 0]), atol=1e-7)
    # zero singular values are kept with an orthonormal isometry (core_test.py:932-944)
    z = np.zeros((1, 8, 8), np.complex64)
    z[0, 0, 0] = 1
    z[0, 3, 5] = 0.5
    for lc in (1, 0):
        left, right, sv, info = _svd(z, 8, lc)
        np.testing.assert_allclose(sv[0], [1, 0.5, 0, 0, 0, 0, 0, 0], atol=1e-7)
        iso = left[0] if lc else right[0].conj().T
        np.testing.assert_allclose(iso.conj().T @ iso, np.eye(8), atol=1e-6)
        np.testing.assert_allclose(left[0] @ right[0], z[0], atol=1e-7)
    # all-zero matrix
    left, right, sv, info = _svd(np.zeros((1, 4, 4), np.complex64), 4, 1)
    assert np.all(sv == 0) and np.all(np.isfinite(left)) and np.all(np.isfinite(right))


def test_svd_batch_of_128x128():
    rng = np.random.RandomState(9)
    mats = np.stack([_graded(rng, 128, 128, d) for d in (0.0, 5.0, 20.0, 40.0, 0.0, 10.0)])
    left, right, sv, info = _svd(mats, 64, 1)
    assert (info[:, 0] == 0).all()
    for j in range(len(mats)):
        sref = np.linalg.svd(mats[j].astype(np.complex128), compute_uv=False)
        assert np.abs(sv[j] - sref).max() <= 1e-5 * sref[0]
        np.testing.assert_allclose(left[j].conj().T @ left[j], np.eye(64), atol=2e-5)


@pytest.mark.parametrize("chi", [(1, 1, 1), (2, 1, 2), (4, 8, 2), (16, 16, 16), (33, 17, 40), (64, 64, 64)])
@pytest.mark.parametrize("d", [2, 3, 4])
def test_theta(chi, d):
    import torch
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    cl, cm, cr = chi
    rng = np.random.RandomState(cl * 7 + cm * 3 + cr + d)
    nb = 3
    A, B = _rand(rng, nb, cl, d, cm), _rand(rng, nb, cm, d, cr)
    G = _rand(rng, nb, d, d, d, d)
    dA, dB, dG = (torch.from_numpy(x).cuda() for x in (A, B, G))
    desc = np.zeros(1, dtype=_lib.GATE2_DESC)
    desc[0] = (dA.data_ptr(), dB.data_ptr(), 0, 0, dG.data_ptr(), 0, cl * d * cm, cm * d * cr, 0, 0, d ** 4, 0)
    ddesc = _lib.to_device_bytes(desc, "cuda")
    out = torch.zeros((nb, d * cl, d * cr), dtype=torch.complex64, device="cuda")
    _lib.check(lib.mpsb_theta(ddesc.data_ptr(), 1, nb, d, cl, cm, cr, out.data_ptr(), None, 0, _lib.stream_ptr()))
    ref = np.einsum("bxypq,blpm,bmqr->blxyr", G.astype(np.complex128), A.astype(np.complex128), B.astype(np.complex128))
    ref = ref.reshape(nb, cl * d, d * cr)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=3e-5 * np.sqrt(cm) * 4)


@pytest.mark.parametrize("scale", [1e-18, 1e-9, 1e6])
def test_svd_is_scale_robust(scale):
    """Tiny / huge entries (products of numerical zeros appear in structured circuits) must not
    overflow the Householder or rotation formulas."""
    rng = np.random.RandomState(3)
    mats = np.stack([_graded(rng, 16, 16, 5.0), _graded(rng, 16, 16, 60.0)]) * np.float32(scale)
    mats[1, :, 3] = 0
    mats[1, 5, :] *= np.float32(1e-12)
    for lc in (1, 0):
        left, right, sv, info = _svd(mats.astype(np.complex64), 16, lc)
        assert np.all(np.isfinite(left)) and np.all(np.isfinite(right)) and np.all(np.isfinite(sv))
        for j in range(2):
            sref = np.linalg.svd(mats[j].astype(np.complex128), compute_uv=False)
            assert np.abs(sv[j] - sref).max() <= 1e-5 * sref[0]
            np.testing.assert_allclose(left[j] @ right[j], mats[j], atol=3e-6 * sref[0])


def test_svd_denormal_diagonal_regression():
    import os
    M = np.load(os.path.join(os.path.dirname(__file__), "golden", "ghz_theta_16x8.npy"))
    for lc in (1, 0):
        left, right, sv, info = _svd(M[None], 8, lc)
        np.testing.assert_allclose(left[0] @ right[0], M, atol=1e-6)
        iso = left[0] if lc else right[0].conj().T
        np.testing.assert_allclose(iso.conj().T @ iso, np.eye(8), atol=2e-6)


@pytest.mark.parametrize("shape", [(256, 256), (200, 136), (130, 300), (512, 512), (160, 128)])
@pytest.mark.parametrize("lc", [1, 0])
def test_svd_large_matches_lapack(shape, lc):
    """d*chi > 128: block Jacobi (Gram + Hermitian EVD + apply) in global memory."""
    m, n = shape
    rng = np.random.RandomState(m * 17 + n + lc)
    mats = np.stack([_graded(rng, m, n, 0.0), _graded(rng, m, n, 12.0)])
    k = min(m, n)
    left, right, sv, info = _svd(mats, k, lc)
    assert (info[:, 0] == 0).all(), info
    for j in range(len(mats)):
        sref = np.linalg.svd(mats[j].astype(np.complex128), compute_uv=False)
        assert np.abs(sv[j] - sref).max() <= 1e-5 * sref[0], (info, np.abs(sv[j] - sref).max() / sref[0])
        np.testing.assert_allclose(left[j] @ right[j], mats[j], atol=5e-6 * sref[0])
        iso = left[j] if lc else right[j].conj().T
        np.testing.assert_allclose(iso.conj().T @ iso, np.eye(k), atol=3e-5)
    kk = k // 2
    left, right, sv, info = _svd(mats, kk, lc)
    for j in range(len(mats)):
        u, s, vh = np.linalg.svd(mats[j].astype(np.complex128), full_matrices=False)
        best = (u[:, :kk] * s[:kk]) @ vh[:kk]
        # projection error is second order in the residual non-orthogonality, first order near the cut
        err = np.linalg.norm(left[j] @ right[j] - best) / np.linalg.norm(best)
        assert err < 2e-3, err
        assert abs(np.linalg.norm(left[j] @ right[j]) - np.linalg.norm(best)) < 1e-5 * np.linalg.norm(best)


@pytest.mark.parametrize("tc", [0, 1, 2])
@pytest.mark.parametrize("shape", [(512, 512), (200, 136), (130, 300)])
def test_svd_large_all_apply_kernels(monkeypatch, shape, tc):
    """The block rotation is applied by bj_apply_kernel (FFMA, tc = 0), by bj_apply_tma_kernel (tcgen05
    3xTF32 fed by TMA, tc = 1: the default of full launches) or by bj_apply_tc_kernel (the same MMAs fed
    by loader warps, tc = 2: rows that are not 16-byte aligned); two matrices would always take the
    first.  Force each in turn (the library reads the switches at call time) and hold all three to the
    LAPACK bounds above, in both orientations (ragged row lengths 136 / 300 / 130 / 200)."""
    monkeypatch.setenv("MPSB_LARGE_TC_APPLY", "0" if tc == 0 else "1")
    monkeypatch.setenv("MPSB_LARGE_TMA_APPLY", "1" if tc == 1 else "0")
    test_svd_large_matches_lapack(shape, 1)
    test_svd_large_matches_lapack(shape, 0)


@pytest.mark.parametrize("tg", [0, 1])
@pytest.mark.parametrize("shape", [(512, 512), (160, 128), (130, 300), (200, 136)])
def test_svd_large_both_gram_kernels(monkeypatch, shape, tg):
    """The pair Grams come from the fused FFMA kernel or, for rows of >= 1024 entries in full launches,
    from bj_gram_tc_kernel (tcgen05 3xTF32, two pairs per MMA tile).  Force each in turn on shapes that
    would not choose it: an odd number of pairs (160 rows: the second pair of the last tile is missing),
    ragged row lengths (300, 136: partial chunks) and 512^2."""
    monkeypatch.setenv("MPSB_LARGE_TC_GRAM", str(tg))
    test_svd_large_matches_lapack(shape, 1)
    test_svd_large_matches_lapack(shape, 0)
