"""numpy model of the DEFLATED k_cg (csrc/ba_kernels.cuh), statement for statement, used to validate the recurrences before they
went into the cooperative kernel (which cannot run in the CPU container):

  * one fused reduction per iteration: gamma = (r,r), delta = (S~ r, r), nu = (S~ W)^T r   (k + 2 values)
  * the owner of a row updates   p = r + beta p - W mu,  s = w + beta s - (AW) mu,  x += alpha p,  r' = r - alpha s
  * a neighbour's r' is recomputed lazily from its OLD (r, w, s) only; the missing deflation term is added by the owner through
    Z = S~ (S~ W), precomputed once per solve:      w'_i = sum_j B_ij [r_j - alpha (w_j + beta s_j)] + alpha (Z mu)_i
  * (p, S~ p) = delta - beta gamma / alpha_prev - mu^T nu
  * basis: Ritz vectors harvested from the residual history of the FIRST (undeflated) solve of a run (Lanczos tridiagonal from
    the CG coefficients), kept in the UNSCALED unknowns (W_y = Linv^T W~) and re-scaled with every solve's own block-Jacobi
    factors (W~ = L^T W_y).

    python tests/scripts/deflated_cg_kernel_model.py [V] [P]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["ANALYTIC"] = "none"
from deflated_cg_prototype import reduced_system_big  # noqa: E402
from ptz_calib_b200 import synth  # noqa: E402


def kernel_cg(A, b, tol, W=None, max_iter=2000, harvest=0):
    """the iteration of k_cg.  Returns x, iterations, history (residuals, alpha, beta, gamma) when harvest > 0"""
    n = len(b)
    k = 0 if W is None else W.shape[1]
    x = np.zeros(n)
    r = b.copy()
    if k:
        AW = A @ W
        Z = A @ AW
        E = W.T @ AW
        Ec = np.linalg.cholesky(E)
        Einv = np.linalg.inv(E)
        c0 = np.linalg.solve(Ec.T, np.linalg.solve(Ec, W.T @ b))
        x = W @ c0
        r = b - AW @ c0
    w = np.zeros(n); s = np.zeros(n); p = np.zeros(n)
    alpha = beta = 0.0
    mu = np.zeros(k)
    gamma_old = gamma0 = 0.0
    hist_r, hist_a, hist_b, hist_g = [], [], [], []
    it = 0
    b2 = b @ b
    status = 1
    while True:
        # owner update from the OLD state
        pn = r + beta * p - (W @ mu if k else 0.0)
        sn = w + beta * s - (AW @ mu if k else 0.0)
        x = x + alpha * pn
        rn = r - alpha * sn
        # lazy neighbours: old state only, plus the owner's Z term
        lazy = r - alpha * (w + beta * s)
        wn = A @ lazy + (alpha * (Z @ mu) if k else 0.0)
        p, s, r, w = pn, sn, rn, wn
        if harvest and it < harvest:
            hist_r.append(rn.copy())
        g, d = rn @ rn, wn @ rn
        nu = AW.T @ rn if k else None
        if it == 0:
            gamma0 = b2  # (the kernel measures against |b~|, also when deflation starts from x0 != 0)
            if not g > 0:
                status = 0
                break
            mu = Einv @ nu if k else mu
            den = d - (mu @ nu if k else 0.0)
            if not den > 0:
                status = 2
                break
            beta = 0.0
            alpha = g / den
        else:
            if np.sqrt(g) <= tol * np.sqrt(gamma0):
                status = 0
                break
            if it >= max_iter:
                break
            beta = g / gamma_old
            mu = Einv @ nu if k else mu
            den = d - beta * g / alpha - (mu @ nu if k else 0.0)
            if not den > 0:
                status = 2
                break
            alpha = g / den
        hist_a.append(alpha); hist_b.append(beta); hist_g.append(g)
        gamma_old = g
        it += 1
    return x, it, status, (hist_r, hist_a, hist_b, hist_g)


def ritz_from_history(hist, k):
    """what the host does after the first solve: tridiagonal from (alpha_j, beta_j), lowest k Ritz vectors in the residual basis.
    hist_a[j] is the step length used from residual j to j+1, hist_b[j] the beta that produced direction j (beta_0 = 0)."""
    R, al, be, ga = hist
    m = min(len(R), len(al))
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1.0 / al[j] + (be[j] / al[j - 1] if j > 0 else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = -np.sqrt(be[j + 1]) / al[j]
    th, Y = np.linalg.eigh(T)
    Rn = np.stack([R[j] / np.sqrt(ga[j]) for j in range(m)], 1)
    W = Rn @ Y[:, :k]
    return th[:k], W


def pivoted_basis(W, tol=1e-8):
    """Gram matrix + pivoted Cholesky: drop (numerically) dependent columns, as the host does before uploading the basis"""
    G = W.T @ W
    k = G.shape[0]
    keep = []
    L = np.zeros((k, k))
    d = np.diag(G).copy()
    d0 = d.max()
    for _ in range(k):
        j = int(np.argmax(d))
        if d[j] <= tol * d0:
            break
        r = len(keep)
        L[j, r] = np.sqrt(d[j])
        for i in range(k):
            if i != j and i not in keep:
                L[i, r] = (G[i, j] - L[i, :r] @ L[j, :r]) / L[j, r]
                d[i] -= L[i, r] ** 2
        d[j] = -1.0
        keep.append(j)
    return W[:, sorted(keep)]


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    nc = 4
    # "LM iteration 1" (harvest) and a later iterate (use): other parameters, other trust-region radius
    q = synth.make_ba_scene(V, P, "band", seed=1004)
    p = synth.make_ba_scene(V, P, "band", seed=1004, rot_noise_deg=0.2, focal_noise=0.008)
    A1, b1, _, Linv1 = reduced_system_big(q, 1e4)
    A2, b2, _, Linv2 = reduced_system_big(p, 3e4)
    xd = np.linalg.solve(A2, b2)
    x, it0, st, hist = kernel_cg(A1, b1, 1e-13, harvest=256)
    print(f"V={q.V}: first solve (undeflated, harvesting) {it0} it, status {st}, err {np.abs(x - np.linalg.solve(A1, b1)).max():.1e}")
    x, it, st, _ = kernel_cg(A2, b2, 1e-13)
    print(f"later iterate, undeflated: {it} it (err {np.abs(x - xd).max() / np.abs(xd).max():.1e})")
    for k in (8, 16, 24):
        th, Wt = ritz_from_history(hist, k)
        Wt = pivoted_basis(Wt)
        Wy = Linv1.T @ Wt                      # unscaled unknowns y = Linv^T y~
        W2 = np.linalg.inv(Linv2).T @ Wy       # this solve's scaling: y~ = L^T y
        x, it, st, _ = kernel_cg(A2, b2, 1e-13, W2)
        x1, it1, st1, _ = kernel_cg(A1, b1, 1e-13, Wt)
        print(f"  k={k:2d} (kept {Wt.shape[1]}): later iterate {it:4d} it status {st} (err {np.abs(x - xd).max() / np.abs(xd).max():.1e}); "
              f"same system {it1} it; lowest Ritz {th[:4]}")


if __name__ == "__main__":
    main()
