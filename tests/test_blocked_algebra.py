"""CPU statement (numpy, fp64) of the tile algebra the large-n GP path runs on the device (csrc/gp_big.cu), checked against
dense linear algebra: left-looking blocked Cholesky with the fused panel solve, U = L^-T by block distance, the lower block
triangle of Khat^-1 = U U^T contracted on the fly with row sums AND column sums, the analytic output-scale identity, the
shift-invariance of the feature gradient, and the joint-likelihood identity the predictive path uses (csrc/gp_post.cu).
No kernels are involved: this pins the derivations in DESIGN.md sections 5a / 5b independently of the GPU tests."""
import numpy as np

NB = 4          # tile edge of this restatement (128 on the device): small, so that several tiles fit a quick test


def _setup(n, seed=0, F=2):
    rs = np.random.RandomState(seed)
    u = rs.normal(size=(n, F))                       # scaled features (z / lengthscale)
    r = rs.normal(size=n)                            # residuals y - m
    s, sig2 = 0.7, 0.3                               # output scale, noise variance
    tot, rho = s + sig2, s / (s + sig2)
    d2 = ((u[:, None, :] - u[None, :, :]) ** 2).sum(-1)
    k = np.exp(-0.5 * d2)                            # unnormalised kernel, k_aa = 1
    Khat = rho * k + (1 - rho) * np.eye(n)           # = (s k + sig2 I) / tot
    return u, r, s, sig2, tot, rho, k, Khat


def _tiles(n):
    nb = (n + NB - 1) // NB
    return nb, nb * NB


def _pad(M, npad):
    out = np.eye(npad)
    out[:M.shape[0], :M.shape[1]] = M
    return out


def _blocked_factor(Kp, nb):
    """Left-looking: C_ik = K_ik - sum_{j<k} L_ij L_kj^T; diagonal tile potrf + inverse; panel L_ik = C_ik Linv_kk^T."""
    T = lambda M, i, j: M[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]          # noqa: E731
    L = np.zeros_like(Kp)
    Linv = [None] * nb
    for kk in range(nb):
        C = T(Kp, kk, kk) - sum(T(L, kk, j) @ T(L, kk, j).T for j in range(kk))
        Lkk = np.linalg.cholesky(C)
        T(L, kk, kk)[:] = Lkk
        Linv[kk] = np.linalg.inv(Lkk)
        for i in range(kk + 1, nb):
            C = T(Kp, i, kk) - sum(T(L, i, j) @ T(L, kk, j).T for j in range(kk))
            T(L, i, kk)[:] = C @ Linv[kk].T
    return L, Linv


def _blocked_u(L, Linv, nb):
    """U = L^-T: U_aa = Linv_aa^T; U_ab = -(sum_{j=a}^{b-1} U_aj L_bj^T) Linv_bb^T, by block distance b - a."""
    T = lambda M, i, j: M[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]          # noqa: E731
    U = np.zeros_like(L)
    for a in range(nb):
        T(U, a, a)[:] = Linv[a].T
    for dist in range(1, nb):
        for a in range(nb - dist):
            b = a + dist
            S = sum(T(U, a, j) @ T(L, b, j).T for j in range(a, b))
            T(U, a, b)[:] = -S @ Linv[b].T
    return U


def test_blocked_cholesky_inverse_and_gradient_contraction_match_dense_algebra():
    n = 14                                           # 4 tiles, the last one half empty
    u, r, s, sig2, tot, rho, k, Khat = _setup(n)
    nb, npad = _tiles(n)
    Kp = _pad(Khat, npad)                            # identity padding rows, as the kernels use
    L, Linv = _blocked_factor(Kp, nb)
    assert np.allclose(L @ L.T, Kp, atol=1e-12)
    U = _blocked_u(L, Linv, nb)
    assert np.allclose(U, np.linalg.inv(L).T, atol=1e-10)
    rp = np.zeros(npad); rp[:n] = r
    v = np.linalg.solve(L, rp)
    alpha = U @ v                                    # alphahat = Khat^-1 r
    assert np.allclose(alpha[:n], np.linalg.solve(Khat, r), atol=1e-10)

    # Khat^-1 tiles of the LOWER block triangle only; row sums for block a, column sums for block b < a
    up = np.full((npad, u.shape[1]), 1e9); up[:n] = u            # padding rows "far away": k = 0 against everything
    T = lambda M, i, j: M[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]          # noqa: E731
    S1 = np.zeros((npad, u.shape[1]))
    trace = 0.0
    for a in range(nb):
        for b in range(a + 1):
            Kinv = sum(T(U, a, m) @ T(U, b, m).T for m in range(a, nb))
            ra, cb = slice(a * NB, (a + 1) * NB), slice(b * NB, (b + 1) * NB)
            du = up[ra, None, :] - up[None, cb, :]
            kk = np.exp(-0.5 * (du ** 2).sum(-1))
            w = (np.outer(alpha[ra], alpha[cb]) / tot - Kinv) * kk
            S1[ra] += (w[:, :, None] * du).sum(1)                                # rows of block a
            if b < a:
                S1[cb] += -(w[:, :, None] * du).sum(0)                           # rows of block b: column sums, du antisymmetric
            else:
                trace += np.sum(alpha[ra] ** 2 / tot - np.diag(Kinv))
    # dense reference: w_ab = (beta_a alphahat_b - Khat^-1_ab) k_ab, S1_a = sum_b w_ab (u_a - u_b)
    Kinv_d = np.linalg.inv(Khat)
    a_d = Kinv_d @ r
    w_d = (np.outer(a_d, a_d) / tot - Kinv_d) * k
    S1_d = (w_d[:, :, None] * (u[:, None, :] - u[None, :, :])).sum(1)
    assert np.allclose(S1[:n], S1_d, atol=1e-9)
    assert np.allclose(S1[:n].sum(0), 0.0, atol=1e-9)                            # shift invariance: the projection removes only rounding
    valid = np.arange(npad) < n
    trace_valid = np.sum((alpha ** 2 / tot - np.diag(U @ U.T))[valid])
    assert np.isclose(trace_valid, np.sum(a_d ** 2) / tot - np.trace(Kinv_d))

    # gradient of L = log N(r | 0, s k + sig2 I) w.r.t. the features, through dL/dK = 1/2 (alpha alpha^T - K^-1):
    # dL/du_a = - sum_b 2 G_ab s k_ab (u_a - u_b) = - rho S1_a   (G = (beta alphahat^T - Khat^-1) / (2 tot) in normalised terms)
    Kt = s * k + sig2 * np.eye(n)
    Kt_inv = np.linalg.inv(Kt)
    al = Kt_inv @ r
    G = 0.5 * (np.outer(al, al) - Kt_inv)
    dLdu = -(2 * (G * s * k)[:, :, None] * (u[:, None, :] - u[None, :, :])).sum(1)
    assert np.allclose(dLdu, -rho * S1[:n], atol=1e-9)

    # output scale without element-wise cancellation: sum_ab (beta_a alphahat_b - Khat^-1_ab) k_ab = (quad - n - (1 - rho) tr) / rho
    quad = r @ a_d / tot
    lhs = np.sum(w_d)
    assert np.isclose(lhs, (quad - n - (1 - rho) * (np.sum(a_d ** 2) / tot - np.trace(Kinv_d))) / rho)
    assert np.isclose(0.5 * lhs / tot, np.sum(G * k))                            # = dL/ds
    # value: quad and log det from the factor
    logdet = n * np.log(tot) + 2 * np.sum(np.log(np.diag(L)[:n]))
    assert np.isclose(logdet, np.linalg.slogdet(Kt)[1]) and np.isclose(quad, r @ al)


def test_joint_likelihood_identity_of_the_predictive_path():
    """log N(y* | mu*, Sigma*) = log N([y_c; y*] | prior) - log N(y_c | prior): what pacoh_gp_posterior computes with two
    marginal-likelihood calls instead of an n* x n* factorisation; and mu / var through w = L^-1 khat."""
    rs = np.random.RandomState(3)
    nc, ns = 5, 9
    x = rs.normal(size=(nc + ns, 2))
    y = rs.normal(size=nc + ns)
    s, sig2 = 0.8, 0.2
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    K = s * np.exp(-0.5 * d2) + sig2 * np.eye(nc + ns)

    def logn(v, C):
        return -0.5 * v @ np.linalg.solve(C, v) - 0.5 * np.linalg.slogdet(C)[1] - 0.5 * len(v) * np.log(2 * np.pi)

    Kcc, Kcs, Kss = K[:nc, :nc], K[:nc, nc:], K[nc:, nc:]
    mu = Kcs.T @ np.linalg.solve(Kcc, y[:nc])
    Sigma = Kss - Kcs.T @ np.linalg.solve(Kcc, Kcs)
    assert np.isclose(logn(y[nc:] - mu, Sigma), logn(y, K) - logn(y[:nc], Kcc))
    tot = s + sig2
    Lc = np.linalg.cholesky(Kcc / tot)
    W = np.linalg.solve(Lc, Kcs / tot)               # w_j = L^-1 khat_j
    v = np.linalg.solve(Lc, y[:nc])
    assert np.allclose(W.T @ v, mu) and np.allclose(tot * (1 - (W ** 2).sum(0)), np.diag(Sigma))
    assert np.allclose(tot * (K[nc:, nc:] / tot - W.T @ W), Sigma)
