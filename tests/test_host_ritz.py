"""Host part of the CG deflation (ptz-calib_b200/csrc/cg_ritz.hpp): the Lanczos tridiagonal rebuilt from the CG coefficients, its
lowest Ritz pairs (Sturm bisection + inverse iteration) and the ghost filter, against numpy on a CG run of the kernel's own
single-reduction recurrences (tests/scripts/deflated_cg_kernel_model.py restates k_cg)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ritz_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("ritz") / "host_ritz_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(out), os.path.join(ROOT, "tests", "host_ritz_check.cpp")], check=True)
    return str(out)


def cg_run(A, b, tol=1e-13, W=None):
    """single-reduction CG as k_cg runs it (owner update + lazy neighbours + Z term), returns x, iterations, (R, alpha, beta, gamma)"""
    n = len(b)
    k = 0 if W is None else W.shape[1]
    x = np.zeros(n); r = b.copy()
    if k:
        AW = A @ W; Z = A @ AW; E = W.T @ AW; Einv = np.linalg.inv(E)
        c0 = Einv @ (W.T @ b); x = W @ c0; r = b - AW @ c0
    w = np.zeros(n); s = np.zeros(n); p = np.zeros(n); mu = np.zeros(k)
    alpha = beta = 0.0; g_old = 0.0; g0 = b @ b
    R, abg = [], []
    it = 0
    while True:
        pn = r + beta * p - (W @ mu if k else 0.0)
        sn = w + beta * s - (AW @ mu if k else 0.0)
        x = x + alpha * pn
        rn = r - alpha * sn
        wn = A @ (r - alpha * (w + beta * s)) + (alpha * (Z @ mu) if k else 0.0)
        p, s, r, w = pn, sn, rn, wn
        R.append(rn.copy())
        g, d = rn @ rn, wn @ rn
        nu = AW.T @ rn if k else None
        if it > 0 and np.sqrt(g) <= tol * np.sqrt(g0):
            break
        beta = 0.0 if it == 0 else g / g_old
        mu = Einv @ nu if k else mu
        den = d - (beta * g / alpha if it else 0.0) - (mu @ nu if k else 0.0)
        assert den > 0
        alpha = g / den
        abg.append((alpha, beta, g))
        g_old = g
        it += 1
        assert it < 5000
    return x, it, (R, np.array(abg))


def spd_with_small_modes(n, seed, small=(2e-4, 3e-4, 5e-4, 4e-3, 9e-3, 2e-2)):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ev = np.concatenate([np.array(small), rng.uniform(0.1, 1.1, n - len(small))])
    return (Q * ev) @ Q.T, rng.normal(size=n), np.sort(ev)


def call(ritz_bin, abg, kmax):
    m = len(abg)
    txt = f"{m} {kmax}\n" + "\n".join(" ".join(repr(float(v)) for v in row) for row in abg) + "\n"
    out = subprocess.run([ritz_bin], input=txt, capture_output=True, text=True, check=True).stdout.split("\n")
    kd = int(out[0])
    val = np.array([float(v) for v in out[1].split()])
    Y = np.array([[float(v) for v in line.split()] for line in out[2 : 2 + m]])
    return kd, val, Y


def test_ritz_values_and_deflation(ritz_bin):
    A, b, ev = spd_with_small_modes(400, 5)
    x, it, (R, abg) = cg_run(A, b)
    assert np.abs(A @ x - b).max() <= 1e-10
    kd, val, Y = call(ritz_bin, abg, 16)
    assert kd == 16
    # reference: dense eigen-decomposition of the same tridiagonal
    m = len(abg)
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1 / abg[j, 0] + (abg[j, 1] / abg[j - 1, 0] if j else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = -np.sqrt(abg[j + 1, 1]) / abg[j, 0]
    th = np.linalg.eigvalsh(T)
    # every kept value is an eigenvalue of T; ghosts (repeats) are skipped, so the kept ones are strictly increasing
    assert all(np.abs(th - v).min() <= 1e-9 * max(v, 1e-12) for v in val)
    assert (np.diff(val) > 1e-6 * val[1:]).all()
    # the six planted small eigenvalues are found
    assert np.abs(val[:6] - ev[:6]).max() <= 1e-6 * ev[5]
    # the Ritz vectors deflate: same solution in far fewer iterations
    Wt = np.stack(R[:m], 1) @ Y
    x2, it2, _ = cg_run(A, b, W=Wt)
    assert np.abs(x2 - x).max() <= 1e-9 * np.abs(x).max()
    assert it2 < 0.7 * it, (it, it2)
    # a stale basis (another right-hand side, shifted spectrum) still works
    A3 = A + 0.01 * np.eye(len(b))
    b3 = np.random.default_rng(9).normal(size=len(b))
    x3, it3, _ = cg_run(A3, b3)
    x4, it4, _ = cg_run(A3, b3, W=Wt)
    assert np.abs(x4 - x3).max() <= 1e-9 * np.abs(x3).max() and it4 <= it3


def test_ritz_degenerate_inputs(ritz_bin):
    assert call(ritz_bin, np.array([[1.0, 0.0, 1.0]]), 4)[0] == 0          # one iteration: nothing to harvest
    assert call(ritz_bin, np.array([[1.0, 0.0, 1.0], [-1.0, 0.5, 0.5]]), 4)[0] == 0  # non-positive alpha
