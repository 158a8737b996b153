"""numpy emulation of the formulas in ptdeco_b200/csrc/eigh.cu (panel recurrences, compact-WY T,
twisted factorisation, back-transformation). Design-time check of the algebra; not a test of the
CUDA code and not used by the product."""
import numpy as np

NB = 8


def sytrd(A):
    d = A.shape[0]
    A = A.copy()
    dvec = np.zeros(d); evec = np.zeros(d); panels = []
    for j0 in range(0, d, NB):
        m = d - j0
        ncols = min(NB, m)
        V = np.zeros((m, NB)); W = np.zeros((m, NB)); T = np.zeros((NB, NB))
        At = A[j0:, j0:]
        for i in range(ncols):
            a = At[i, :].copy()  # row i == column i
            a[:i] = 0
            for r in range(i, m):
                a[r] -= V[r, :i] @ W[i, :i] + W[r, :i] @ V[i, :i]
            dvec[j0 + i] = a[i]
            if i + 1 >= m:
                break
            alpha = a[i + 1]
            xn2 = np.sum(a[i + 2:] ** 2)
            if xn2 > 0:
                beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
                tau = (beta - alpha) / beta
                scale = 1.0 / (alpha - beta)
            else:
                beta, tau, scale = alpha, 0.0, 0.0
            evec[j0 + i] = beta
            v = np.zeros(m); v[i + 1] = 1.0; v[i + 2:] = a[i + 2:] * scale
            V[:, i] = v
            if tau != 0:
                p = At @ v
                gW = W[:, :i].T @ v; gV = V[:, :i].T @ v; vp = v @ p
                dot = tau * (vp - 2 * gV @ gW)
                alpha2 = -0.5 * tau * dot
                w = tau * (p - V[:, :i] @ gW - W[:, :i] @ gV) + alpha2 * v
                w[:i + 1] = 0
                W[:, i] = w
                for c in range(i):
                    T[c, i] = -tau * (T[c, c:i] @ gV[c:i])
                T[i, i] = tau
        panels.append((j0, ncols, V.copy(), T.copy()))
        if m > NB:
            At[NB:, NB:] -= V[NB:] @ W[NB:].T + W[NB:] @ V[NB:].T
    return dvec, evec[:d - 1], panels


def twisted(D, E, lam):
    n = len(D)
    Dp = np.zeros(n); Dm = np.zeros(n)
    Dp[0] = D[0] - lam
    for i in range(n - 1):
        l = E[i] / Dp[i]
        Dp[i + 1] = (D[i + 1] - lam) - l * E[i]
    Dm[n - 1] = D[n - 1] - lam
    best, r = abs(Dp[n - 1]), n - 1
    for i in range(n - 2, -1, -1):
        u = E[i] / Dm[i + 1]
        Dm[i] = (D[i] - lam) - u * E[i]
        gam = abs(Dp[i] + Dm[i] - (D[i] - lam))
        if gam < best:
            best, r = gam, i
    z = np.zeros(n); z[r] = 1
    for i in range(r - 1, -1, -1):
        z[i] = -(E[i] / Dp[i]) * z[i + 1]
    for i in range(r, n - 1):
        z[i + 1] = -(E[i] / Dm[i + 1]) * z[i]
    return z / np.linalg.norm(z)


def sturm(D, E2, x):
    q = D[0] - x; c = int(q < 0)
    for i in range(1, len(D)):
        q = (D[i] - x) - E2[i - 1] / q
        c += int(q < 0)
    return c


def main():
    rng = np.random.default_rng(0)
    for d in (21, 40, 67):
        X = rng.standard_normal((3 * d, d)) * np.logspace(0, -2, d)
        A = X.T @ X / (3 * d)
        dv, ev, panels = sytrd(A)
        Tm = np.diag(dv) + np.diag(ev, 1) + np.diag(ev, -1)
        lam_ref = np.linalg.eigvalsh(A)
        print(d, "tridiag eigenvalue err", np.abs(np.linalg.eigvalsh(Tm) - lam_ref).max())
        # sturm sanity
        E2 = ev ** 2
        assert sturm(dv, E2, lam_ref[3] + 1e-9) == 4
        Z = np.stack([twisted(dv, ev, l) for l in lam_ref], axis=1)
        print("   Z orth", np.abs(Z.T @ Z - np.eye(d)).max(), "resid", np.abs(Tm @ Z - Z * lam_ref).max())
        U = Z.copy()
        for j0, nc, V, T in reversed(panels):
            m = d - j0
            if m < 2:
                continue
            X1 = V[:, :nc].T @ U[j0:]
            X2 = np.triu(T[:nc, :nc]) @ X1
            U[j0:] -= V[:, :nc] @ X2
        print("   U orth", np.abs(U.T @ U - np.eye(d)).max(), "resid", np.abs(A @ U - U * lam_ref).max())


if __name__ == "__main__":
    main()
