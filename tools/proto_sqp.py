"""numpy prototype of the GPU algorithm (SQP + Goldfarb-Idnani dual active set) -- development aid.
Not shipped, not imported by the product; uses the oracle's problem definition to explore
convergence behaviour before the CUDA implementation is written."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import numpy as np
import ftmpc_oracle as o


def gi_qp(G, a, C, b, maxit=2000, tol=1e-10):
    """min 1/2 x'Gx + a'x  s.t. C x <= b.   Goldfarb-Idnani.  Returns x, lam(m), active list, status, iters"""
    n = G.shape[0]; m = C.shape[0]
    L = np.linalg.cholesky(G)
    J = np.linalg.inv(L).T.copy()          # J = L^-T
    x = -J @ (J.T @ a)
    A = []                                  # active indices (ordered)
    R = np.zeros((n, n)); q = 0
    u = np.zeros(0)
    Nn = -C                                 # normals: n_i' x >= beta_i
    beta = -b
    it = 0
    cnorm = np.linalg.norm(C, axis=1) + 1e-300
    while True:
        s = Nn @ x - beta
        s_sc = s.copy()
        s_sc[A] = np.inf
        p = int(np.argmin(s_sc / 1.0))
        if s_sc[p] >= -tol * max(1.0, cnorm[p]):
            lam = np.zeros(m); lam[A] = u
            return x, lam, list(A), 0, it
        npl = Nn[p]
        up = np.concatenate([u, [0.0]])
        while True:
            it += 1
            if it > maxit:
                lam = np.zeros(m); lam[A] = u
                return x, lam, list(A), 1, it
            d = J.T @ npl
            z = J[:, q:] @ d[q:]
            r = np.linalg.solve(R[:q, :q], d[:q]) if q > 0 else np.zeros(0)
            # step lengths
            t1 = np.inf; l = -1
            for j in range(q):
                if r[j] > 1e-14:
                    tj = up[j] / r[j]
                    if tj < t1: t1 = tj; l = j
            zn = z @ npl
            znorm2 = z @ z
            dep = znorm2 <= 1e-22 * max(1.0, (J @ d) @ (J @ d)) or zn <= 1e-14 * np.sqrt(znorm2) * np.linalg.norm(npl)
            sp = npl @ x - beta[p]
            t2 = np.inf if dep else -sp / zn
            t = min(t1, t2)
            if t == np.inf:
                lam = np.zeros(m); lam[A] = u
                return x, lam, list(A), 2, it            # infeasible
            if t2 == np.inf:
                up = up + t * np.concatenate([-r, [1.0]])
                # drop l
                J, R, q, A, up = drop(J, R, q, A, up, l)
                continue
            x = x + t * z
            up = up + t * np.concatenate([-r, [1.0]])
            if t == t2:
                # add p: householder on d[q:]
                d2 = d[q:].copy()
                alpha = np.linalg.norm(d2)
                sgn = 1.0 if d2[0] >= 0 else -1.0
                v = d2.copy(); v[0] += sgn * alpha
                vv = v @ v
                if vv > 0:
                    J[:, q:] -= np.outer(J[:, q:] @ v, v) * (2.0 / vv)
                R[:q, q] = d[:q]; R[q, q] = -sgn * alpha
                q += 1; A.append(p); u = up
                break
            else:
                J, R, q, A, up = drop(J, R, q, A, up, l)
                continue


def drop(J, R, q, A, up, l):
    # remove column l from R (q columns), restore triangular by Givens on rows, apply to J columns
    R[:, l:q - 1] = R[:, l + 1:q]
    R[:, q - 1] = 0
    for k in range(l, q - 1):
        a_, b_ = R[k, k], R[k + 1, k]
        h = np.hypot(a_, b_)
        if h == 0: continue
        c_, s_ = a_ / h, b_ / h
        rk = R[k, k:q - 1].copy(); rk1 = R[k + 1, k:q - 1].copy()
        R[k, k:q - 1] = c_ * rk + s_ * rk1
        R[k + 1, k:q - 1] = -s_ * rk + c_ * rk1
        jk = J[:, k].copy(); jk1 = J[:, k + 1].copy()
        J[:, k] = c_ * jk + s_ * jk1
        J[:, k + 1] = -s_ * jk + c_ * jk1
    A = A[:l] + A[l + 1:]
    up = np.concatenate([up[:l], up[l + 1:]])
    return J, R, q - 1, A, up


def terminal_hess(eN, h=1e-6):
    H = np.zeros((9, 9))
    for i in range(9):
        e = np.zeros(9); e[i] = h
        H[:, i] = (o.TERMINAL.grad(eN + e) - o.TERMINAL.grad(eN - e)) / (2 * h)
    return 0.5 * (H + H.T)


def psd(H):
    w, V = np.linalg.eigh(H)
    return (V * np.maximum(w, 0.0)) @ V.T


def gn_hessian(prob, X, U):
    N = prob.N
    G, Hx = prob.sensitivities(X, U)
    H = np.diag(np.tile(2 * prob.R, N))
    for t in range(N):
        H += G[t][0:9].T @ ((2 * prob.Q)[:, None] * G[t][0:9])
    HN = terminal_hess(X[N, 0:9] - prob.xref[N])
    H += G[N][0:9].T @ psd(HN) @ G[N][0:9]
    return H


def exact_hessian(prob, U, lam, h=1e-6):
    n = U.size
    def gl(Uv):
        f, g, c, J, X = prob.fun_and_grad(Uv)
        return g + J.T @ lam
    H = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = h
        H[:, i] = (gl(U + e) - gl(U - e)) / (2 * h)
    return 0.5 * (H + H.T)


def merit(prob, U, nu):
    c, X = prob.ineq(U)
    f = prob.objective_from(X, U.reshape(prob.N, 6))
    return f + nu * np.sum(np.maximum(c, 0.0)), f, c


def sqp(prob, U0=None, maxit=50, tol=1e-9, hess="gn", verbose=True):
    N = prob.N; n = 6 * N
    U = np.zeros(n) if U0 is None else U0.copy()
    nu = 1.0
    lam = None
    hist = []
    for it in range(maxit):
        f, g, c, Jc, X = prob.fun_and_grad(U)
        H = gn_hessian(prob, X, U.reshape(N, 6))
        mode = "gn"
        if hess in ("exact", "theta", "lm") and lam is not None:
            He = exact_hessian(prob, U, lam)
            if hess == "exact":
                try:
                    np.linalg.cholesky(He); H = He; mode = "ex"
                except np.linalg.LinAlgError:
                    mode = "gn(indef)"
            elif hess == "theta":
                S = He - H
                for th in (1.0, 0.5, 0.25, 0.125, 0.0):
                    try:
                        np.linalg.cholesky(H + th * S); break
                    except np.linalg.LinAlgError:
                        pass
                H = H + th * S; mode = "th%.3g" % th
            else:
                sc = np.mean(np.diag(H))
                for mu in (0.0, 1e-3, 1e-2, 1e-1, 1.0, 10.0, 100.0):
                    try:
                        np.linalg.cholesky(He + mu * sc * np.eye(n)); break
                    except np.linalg.LinAlgError:
                        pass
                H = He + mu * sc * np.eye(n); mode = "mu%.3g" % mu
        viol = c > 0
        rho = 1e4
        if viol.any():
            # Powell/Schittkowski relaxation: J d + (1-delta) c_viol <= 0, 0<=delta<=1, cost + rho/2 delta^2
            Ha = np.zeros((n + 1, n + 1)); Ha[:n, :n] = H; Ha[n, n] = rho
            ga = np.concatenate([g, [0.0]])
            Ca = np.zeros((Jc.shape[0] + 2, n + 1)); Ca[:-2, :n] = Jc; Ca[:-2, n] = np.where(viol, -c, 0.0)
            Ca[-2, n] = -1.0; Ca[-1, n] = 1.0
            ba = np.concatenate([-c, [0.0, 1.0]])
            da, lam_a, act, st, qit = gi_qp(Ha, ga, Ca, ba)
            d = da[:n]; delta = da[n]; lam_qp = lam_a[:-2]
        else:
            delta = 0.0
            d, lam_qp, act, st, qit = gi_qp(H, g, Jc, -c)
        if st != 0:
            print("QP status", st); return U, it, False, hist
        lam = lam_qp
        nu = max(nu, 1.1 * lam.max() if lam.size else 1.0)
        m0, f0, c0 = merit(prob, U, nu)
        dphi = g @ d - nu * np.sum(np.maximum(c, 0.0))
        alpha = 1.0
        while True:
            m1, f1, c1 = merit(prob, U + alpha * d, nu)
            if m1 <= m0 + 1e-4 * alpha * dphi or alpha < 1e-8: break
            alpha *= 0.5
        U = U + alpha * d
        dn = np.abs(d).max()
        hist.append(dn)
        if verbose:
            print(f"it {it:2d} {mode:8s} f {f:.9f} |d| {dn:.3e} alpha {alpha:.3g} nact {len(act)} qpit {qit} delta {delta:.3g} viol {max(0,c.max()):.2e} nu {nu:.3g}")
        if dn < tol:
            return U, it + 1, True, hist
    return U, maxit, False, hist


if __name__ == "__main__":
    np.set_printoptions(precision=9, linewidth=200, suppress=True)
    Nh = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    hess = sys.argv[2] if len(sys.argv) > 2 else "gn"
    prob, x0 = o.default_problem(Nh)
    t = time.time()
    with np.errstate(all="ignore"):
        U, its, ok, hist = sqp(prob, hess=hess)
    print("time", time.time() - t, "its", its, ok)
    print("u0", U[:6])
    k = o.kkt_residual(prob, U)
    print("kkt", k["stat"], k["viol"], len(k["active"]), "%.12f" % k["f"])
