"""CPU prototype (numpy/scipy, oracle Jacobians) of the NEXT step on stage 3: single-reduction CG (Chronopoulos-Gear, what k_cg runs)
with a small deflation basis W folded into the same reduction.  It checks the recurrences the kernel would use

    mu = G^-1 (AW)^T r                      G = W^T A W  (k x k, factored once per solve)
    p  = r + beta p - W mu ,   s = A p = w + beta s - (AW) mu          (w = A r)
    (p, A p) = delta - beta gamma / alpha_prev - mu^T G mu             (gamma = (r,r), delta = (w,r))

against textbook CG, and counts iterations to 1e-13 on the Jacobi-scaled reduced camera system of a synthetic PTZ scene for
  (a) no deflation, (b) W = exact lowest eigenvectors, (c) W = lowest eigenvectors of a DIFFERENT LM iterate's system (stale basis:
  other damping, perturbed parameters) — what recycling the basis across LM iterations would give.
    python tests/scripts/deflated_cg_prototype.py [V] [P]"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from ptz_calib_b200 import synth  # noqa: E402


def reduced_system(p, mu, keep=False):
    """Jacobi-scaled, LM-damped reduced camera system S~ (block-Jacobi scaled as k_scale_system does) and its right-hand side"""
    ev = orc.ba_eval(p)
    V, P, M = p.V, p.P, p.M
    sw = np.sqrt(p.track_weight[p.obs_track])
    J = ev.jac_obs * sw[:, None, None]
    r = ev.residuals[:M] * sw[:, None]
    cam_cols = [0, 2, 3, 4]  # PTZRay: fx, (fy: identically zero), w(3)
    nc = len(cam_cols)
    rows = np.repeat(np.arange(2 * M), nc)
    F = sp.csr_matrix((J[:, :, cam_cols].reshape(-1), (rows, (p.obs_view[:, None, None] * nc + np.arange(nc)[None, None, :]).repeat(2, 1).reshape(-1))),
                      shape=(2 * M, V * nc))
    rows3 = np.repeat(np.arange(2 * M), 3)
    E = sp.csr_matrix((J[:, :, 5:8].reshape(-1), (rows3, (p.obs_track[:, None, None] * 3 + np.arange(3)[None, None, :]).repeat(2, 1).reshape(-1))),
                      shape=(2 * M, P * 3))
    # Ceres' Jacobi scaling: columns scaled by 1 / (1 + norm)
    sF = 1.0 / (1.0 + np.sqrt(np.asarray(F.multiply(F).sum(0)).ravel()))
    sE = 1.0 / (1.0 + np.sqrt(np.asarray(E.multiply(E).sum(0)).ravel()))
    F, E = F @ sp.diags(sF), E @ sp.diags(sE)
    g = r.reshape(-1)
    A = (F.T @ F).tocsr()
    C = (E.T @ E).tocsr()
    B = (F.T @ E).tocsr()
    dA = np.clip(A.diagonal(), 1e-6, 1e32) / mu
    dC = np.clip(C.diagonal(), 1e-6, 1e32) / mu
    A = A + sp.diags(dA)
    Cd = (C + sp.diags(dC)).tocsc()
    Cinv = spla.splu(Cd)
    S = A - B @ sp.csr_matrix(Cinv.solve(B.T.toarray()))
    b = F.T @ g - B @ Cinv.solve(E.T @ g)
    S = np.asarray(S.todense())
    # block-Jacobi as a symmetric scaling
    n = V * nc
    Linv = np.zeros((n, n))
    for v in range(V):
        sl = slice(v * nc, (v + 1) * nc)
        Linv[sl, sl] = np.linalg.inv(np.linalg.cholesky(S[sl, sl]))
    St = Linv @ S @ Linv.T
    if keep:
        return 0.5 * (St + St.T), Linv @ b, sF, Linv
    return 0.5 * (St + St.T), Linv @ b


def cg_textbook(A, b, tol):
    x = np.zeros_like(b); r = b.copy(); p = r.copy(); rr = r @ r; b2 = b @ b; it = 0
    while np.sqrt(rr) > tol * np.sqrt(b2) and it < 5000:
        Ap = A @ p; a = rr / (p @ Ap); x += a * p; r -= a * Ap; rn = r @ r; p = r + (rn / rr) * p; rr = rn; it += 1
    return x, it


def cg_single_reduction(A, b, tol, W=None):
    """Chronopoulos-Gear CG, ONE fused reduction per iteration ((r,r), (Ar,r) and, with deflation, (AW)^T r)"""
    n = len(b)
    k = 0 if W is None else W.shape[1]
    if k:
        AW = A @ W
        G = W.T @ AW
        Gc = np.linalg.cholesky(G)
        solveG = lambda y: np.linalg.solve(Gc.T, np.linalg.solve(Gc, y))  # noqa: E731
        x = W @ solveG(W.T @ b)  # start with W^T r0 = 0
    else:
        x = np.zeros(n)
    r = b - A @ x
    p = np.zeros(n); s = np.zeros(n)
    alpha_prev, gamma_prev = 1.0, 1.0
    b2 = b @ b
    it = 0
    while it < 5000:
        w = A @ r                                    # sparse product (neighbour gathers)
        gamma, delta = r @ r, w @ r                  # ---- the one reduction
        c = AW.T @ r if k else None                  # ---- k more partials in the same reduction
        if np.sqrt(gamma) <= tol * np.sqrt(b2):
            break
        beta = 0.0 if it == 0 else gamma / gamma_prev
        if k:
            mu = solveG(c)
            pAp = delta - (beta * gamma / alpha_prev if it else 0.0) - mu @ (G @ mu)
            p = r + beta * p - W @ mu
            s = w + beta * s - AW @ mu
        else:
            pAp = delta - (beta * gamma / alpha_prev if it else 0.0)
            p = r + beta * p
            s = w + beta * s
        alpha = gamma / pAp
        x += alpha * p
        r -= alpha * s
        alpha_prev, gamma_prev = alpha, gamma
        it += 1
    return x, it


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    p = synth.make_ba_scene(V, P, "band", seed=1004)
    print(f"scene V={p.V} P={p.P} M={p.M}")
    A, b = reduced_system(p, 1e4)
    ev, evec = np.linalg.eigh(A)
    print("eigenvalues: min %.2e  #<1e-3: %d  #<1e-2: %d  #<0.1: %d  max %.3f" % (ev[0], (ev < 1e-3).sum(), (ev < 1e-2).sum(), (ev < 0.1).sum(), ev[-1]))
    xd = np.linalg.solve(A, b)
    x0, it0 = cg_textbook(A, b, 1e-13)
    x1, it1 = cg_single_reduction(A, b, 1e-13)
    print(f"textbook CG {it0} it (err {np.abs(x0 - xd).max() / np.abs(xd).max():.1e}); single-reduction CG {it1} it (err {np.abs(x1 - xd).max() / np.abs(xd).max():.1e})")
    # a different LM iterate: perturbed parameters, smaller trust region -> stale basis
    q = synth.make_ba_scene(V, P, "band", seed=1004, rot_noise_deg=0.3, focal_noise=0.01)
    A2, _ = reduced_system(q, 3e2)
    ev2, evec2 = np.linalg.eigh(A2)
    for k in (3, 6, 10, 16):
        xe, ite = cg_single_reduction(A, b, 1e-13, evec[:, :k])
        xs, its = cg_single_reduction(A, b, 1e-13, evec2[:, :k])
        print(f"k={k:2d}: exact basis {ite:4d} it (err {np.abs(xe - xd).max() / np.abs(xd).max():.1e})   stale basis {its:4d} it (err {np.abs(xs - xd).max() / np.abs(xd).max():.1e})")


if __name__ == "__main__" and not os.environ.get("ANALYTIC"):
    main()


def analytic_basis(p, A_scaling, kind):
    """Coarse vectors that need no eigen-solve: rotation fields theta_i = a + M d_i over the view directions d_i (world frame), mapped to the
    rvec tangent by the right Jacobian inverse (finite differences here), plus focal modes f_i ~ (1, d_i).  kind: 'gauge' (3), 'affine' (12),
    'affine+f' (16).  Returned in the coordinates of the scaled system: x~ = L^T (x / s)."""
    sF, Linv = A_scaling
    V, nc = p.V, 4
    R = synth.rodrigues_np(p.ext[:, :3])
    d = R[:, 2, :]  # optical axis in the world frame (third row of R)
    h = 1e-6
    Jinv = np.zeros((V, 3, 3))  # d rvec / d theta for R <- R exp([theta])
    for k in range(3):
        e = np.zeros((V, 3)); e[:, k] = h
        Jinv[:, :, k] = (synth.log_so3(R @ synth.rodrigues_np(e)) - synth.log_so3(R @ synth.rodrigues_np(-e))) / (2 * h)
    fields = [np.tile(np.eye(3)[k], (V, 1)) for k in range(3)]
    if kind != "gauge":
        for a in range(3):
            for c in range(3):
                th = np.zeros((V, 3)); th[:, a] = d[:, c]
                fields.append(th)
    cols = []
    for th in fields:
        v = np.zeros((V, nc))
        v[:, 1:4] = np.einsum("vij,vj->vi", Jinv, th)
        cols.append(v.reshape(-1))
    if kind == "affine+f":
        for m in [np.ones(V), d[:, 0], d[:, 1], d[:, 2]]:
            v = np.zeros((V, nc)); v[:, 0] = m * p.intr[:, 0]
            cols.append(v.reshape(-1))
    W = np.stack(cols, 1) / sF[:, None]
    W = np.linalg.solve(Linv.T, W) if False else (np.linalg.inv(Linv).T @ W)  # x~ = L^T x_s, L = inv(Linv)
    Q, _ = np.linalg.qr(W)
    return Q


def main2():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    p = synth.make_ba_scene(V, P, "band", seed=1004)
    # same construction as reduced_system, keeping the scalings
    global _keep
    A, b, sF, Linv = reduced_system(p, 1e4, keep=True)
    xd = np.linalg.solve(A, b)
    _, it = cg_single_reduction(A, b, 1e-13)
    print(f"analytic coarse spaces, V={p.V}: none {it} it")
    for kind in ("gauge", "affine", "affine+f"):
        W = analytic_basis(p, (sF, Linv), kind)
        x, it = cg_single_reduction(A, b, 1e-13, W)
        print(f"  {kind:9s} k={W.shape[1]:2d}: {it:4d} it (err {np.abs(x - xd).max() / np.abs(xd).max():.1e})")


if __name__ == "__main__" and os.environ.get("ANALYTIC") == "1":
    main2()


def reduced_system_big(p, mu):
    """same system as reduced_system, assembled track by track with numpy (dense S~ of 4V x 4V; works at cfg-4 size)"""
    ev = orc.ba_eval(p)
    V, P, M, nc = p.V, p.P, p.M, 4
    sw = np.sqrt(p.track_weight[p.obs_track])
    J = ev.jac_obs * sw[:, None, None]
    r = ev.residuals[:M] * sw[:, None]
    F = J[:, :, [0, 2, 3, 4]]  # [M,2,4]
    E = J[:, :, 5:8]           # [M,2,3]
    colF = np.zeros((V, nc)); np.add.at(colF, p.obs_view, (F ** 2).sum(1))
    colE = np.zeros((P, 3)); np.add.at(colE, p.obs_track, (E ** 2).sum(1))
    sF, sE = 1.0 / (1.0 + np.sqrt(colF)), 1.0 / (1.0 + np.sqrt(colE))
    F = F * sF[p.obs_view][:, None, :]
    E = E * sE[p.obs_track][:, None, :]
    U = np.zeros((V, nc, nc)); np.add.at(U, p.obs_view, np.einsum("mia,mib->mab", F, F))
    gc = np.zeros((V, nc)); np.add.at(gc, p.obs_view, np.einsum("mia,mi->ma", F, r))
    C = np.zeros((P, 3, 3)); np.add.at(C, p.obs_track, np.einsum("mia,mib->mab", E, E))
    gp = np.zeros((P, 3)); np.add.at(gp, p.obs_track, np.einsum("mia,mi->ma", E, r))
    U[:, np.arange(nc), np.arange(nc)] += np.clip(U[:, np.arange(nc), np.arange(nc)], 1e-6, 1e32) / mu
    C[:, np.arange(3), np.arange(3)] += np.clip(C[:, np.arange(3), np.arange(3)], 1e-6, 1e32) / mu
    Ci = np.linalg.inv(C)
    Wo = np.einsum("mia,mib->mab", F, E)  # [M,4,3] = F^T E per observation
    WC = np.einsum("mab,mbc->mac", Wo, Ci[p.obs_track])
    n = V * nc
    S = np.zeros((n, n))
    for v in range(V):
        S[v * nc:(v + 1) * nc, v * nc:(v + 1) * nc] = U[v]
    b = gc - 0.0
    np.add.at(b, p.obs_view, -np.einsum("mab,mb->ma", WC, gp[p.obs_track]))
    # pairs of observations of the same track (including o == o')
    order = np.lexsort((p.obs_view, p.obs_track))
    tt = p.obs_track[order]
    idx = np.arange(nc)
    d = 0
    while True:
        same = tt[:len(tt) - d] == tt[d:] if d < len(tt) else np.zeros(0, bool)
        if not same.any():
            break
        a = order[np.nonzero(same)[0]]
        c = order[np.nonzero(same)[0] + d]
        blk = -np.einsum("mab,mcb->mac", WC[a], Wo[c])  # [m,4,4] contribution to S[view a, view c]
        ra = (p.obs_view[a][:, None] * nc + idx[None, :])
        rc = (p.obs_view[c][:, None] * nc + idx[None, :])
        np.add.at(S, (ra[:, :, None].repeat(nc, 2), rc[:, None, :].repeat(nc, 1)), blk)
        if d > 0:
            np.add.at(S, (rc[:, :, None].repeat(nc, 2), ra[:, None, :].repeat(nc, 1)), blk.transpose(0, 2, 1))
        d += 1
    Linv = np.zeros((n, n))
    for v in range(V):
        sl = slice(v * nc, (v + 1) * nc)
        Linv[sl, sl] = np.linalg.inv(np.linalg.cholesky(S[sl, sl]))
    St = Linv @ S @ Linv.T
    return 0.5 * (St + St.T), Linv @ b.reshape(-1), sF.reshape(-1), Linv


def main3():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
    p = synth.make_ba_scene(V, P, "band", seed=1004)
    print(f"scene V={p.V} P={p.P} M={p.M}", flush=True)
    A, b, sF, Linv = reduced_system_big(p, 1e4)
    ev, evec = np.linalg.eigh(A)
    print("eigenvalues: min %.2e  #<1e-3: %d  #<1e-2: %d  #<0.1: %d  max %.3f" % (ev[0], (ev < 1e-3).sum(), (ev < 1e-2).sum(), (ev < 0.1).sum(), ev[-1]), flush=True)
    xd = np.linalg.solve(A, b)
    _, it = cg_single_reduction(A, b, 1e-13)
    print(f"none: {it} it", flush=True)
    for kind in ("gauge", "affine", "affine+f"):
        W = analytic_basis(p, (sF, Linv), kind)
        x, it = cg_single_reduction(A, b, 1e-13, W)
        print(f"  analytic {kind:9s} k={W.shape[1]:2d}: {it:4d} it (err {np.abs(x - xd).max() / np.abs(xd).max():.1e})", flush=True)
    for k in (6, 10, 16):
        x, it = cg_single_reduction(A, b, 1e-13, evec[:, :k])
        print(f"  exact eigenvectors k={k:2d}: {it:4d} it (err {np.abs(x - xd).max() / np.abs(xd).max():.1e})", flush=True)


if __name__ == "__main__" and os.environ.get("ANALYTIC") == "big":
    main3()


def main4():
    """init-CG: only the initial guess is projected (x0 = W G^-1 W^T b), the iteration itself is the plain one"""
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    p = synth.make_ba_scene(V, P, "band", seed=1004)
    A, b, sF, Linv = reduced_system_big(p, 1e4)
    ev, evec = np.linalg.eigh(A)
    xd = np.linalg.solve(A, b)
    _, it = cg_single_reduction(A, b, 1e-13)
    print(f"V={p.V}: none {it} it")
    bases = [("analytic k=16", analytic_basis(p, (sF, Linv), "affine+f")), ("exact k=10", evec[:, :10]), ("exact k=16", evec[:, :16])]
    for name, W in bases:
        AW = A @ W
        x0 = W @ np.linalg.solve(W.T @ AW, W.T @ b)
        r0 = b - A @ x0
        y, it = cg_single_reduction(A, r0, 1e-13 * np.linalg.norm(b) / np.linalg.norm(r0))
        x = x0 + y
        print(f"  init-CG {name}: {it} it, |r0|/|b| = {np.linalg.norm(r0) / np.linalg.norm(b):.2e} (err {np.abs(x - xd).max() / np.abs(xd).max():.1e})")


if __name__ == "__main__" and os.environ.get("ANALYTIC") == "init":
    main4()


def lanczos_ritz_from_cg(A, b, k, tol=1e-13):
    """plain CG on (A, b) keeping the normalised residuals; lowest k Ritz pairs of the Lanczos tridiagonal built from the CG coefficients"""
    x = np.zeros_like(b); r = b.copy(); p = r.copy(); rr = r @ r; b2 = b @ b
    Rm, alphas, betas = [], [], []
    while np.sqrt(rr) > tol * np.sqrt(b2) and len(alphas) < 5000:
        Rm.append(r / np.sqrt(rr))
        Ap = A @ p; a = rr / (p @ Ap); x += a * p; r = r - a * Ap; rn = r @ r; be = rn / rr; p = r + be * p; rr = rn
        alphas.append(a); betas.append(be)
    m = len(alphas)
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1.0 / alphas[j] + (betas[j - 1] / alphas[j - 1] if j > 0 else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = -np.sqrt(betas[j]) / alphas[j]
    th, Y = np.linalg.eigh(T)
    W = np.stack(Rm, 1) @ Y[:, :k]
    return th[:k], W, m


def main5():
    """recycling: Ritz vectors harvested from the CG run of ONE LM iterate, used as init-CG / deflation basis for ANOTHER iterate"""
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    p = synth.make_ba_scene(V, P, "band", seed=1004)
    q = synth.make_ba_scene(V, P, "band", seed=1004, rot_noise_deg=0.3, focal_noise=0.01)
    A1, b1, _, _ = reduced_system_big(q, 3e2)      # "first LM iteration": harvest here
    A, b, _, _ = reduced_system_big(p, 1e4)        # a later iterate: use there
    xd = np.linalg.solve(A, b)
    _, it = cg_single_reduction(A, b, 1e-13)
    print(f"V={p.V}: none {it} it")
    for k in (6, 10, 16):
        th, W, m = lanczos_ritz_from_cg(A1, b1, k)
        W, _ = np.linalg.qr(W)
        AW = A @ W
        x0 = W @ np.linalg.solve(W.T @ AW, W.T @ b)
        r0 = b - A @ x0
        y, it1 = cg_single_reduction(A, r0, 1e-13 * np.linalg.norm(b) / np.linalg.norm(r0))
        x2, it2 = cg_single_reduction(A, b, 1e-13, W)
        print(f"  Ritz vectors of the other iterate ({m} CG steps there), k={k:2d}: init-CG {it1} it (err {np.abs(x0 + y - xd).max() / np.abs(xd).max():.1e}), "
              f"deflated CG {it2} it (err {np.abs(x2 - xd).max() / np.abs(xd).max():.1e}); lowest Ritz values {th[:3]}")


if __name__ == "__main__" and os.environ.get("ANALYTIC") == "recycle":
    main5()
