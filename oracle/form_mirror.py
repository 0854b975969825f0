"""CPU ORACLE CHECK — TEST INFRASTRUCTURE ONLY.

A literal, line-by-line sympy transcription of the reference's UFL residual for ONE cell
(P1 triangle or Q1 quadrilateral), used to cross-check the hand-vectorised numpy restatements
(`ns_oracle.element_F/J`, `q1_oracle.element_F/J`, `facet_F`):

  * the UFL operators are defined once from their definitions — `grad(f)_i = K_ji d f / d xi_j`,
    `nabla_grad(u)_ij = d_i u_j`, `div`, `sym`, `dot`, `inner` — and the form is then written with
    them exactly as src/solvers/stabilized_schur.py:69-121 does (the cited line is next to each
    statement), including `div(sigma(u_mid, p_sol, mu))` in the strong residual, which sympy
    differentiates through the non-affine geometry like UFL does (no closed-form Hessian);
  * the Jacobian is NOT re-derived: the tests differentiate this residual numerically (complex
    step) and compare with the oracle's hand-derived Jacobian.

Still "parity unpinned" with respect to DOLFINx itself; what this pins is that
the restatement says what the reference source says.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

XI, ETA = sp.symbols("xi eta", real=True)


def _basis(nv):
    if nv == 3:
        return [1 - XI - ETA, XI, ETA]                           # P1 reference basis
    return [(1 - XI) * (1 - ETA), XI * (1 - ETA), (1 - XI) * ETA, XI * ETA]   # Q1, tensor-ordered


class CellForms:
    """Residual of one cell as numeric callables of the nodal values.

    X: (nv, 2) vertex coordinates, Un: (nv, 2) previous velocity, h: cell diameter (DG0 value of
    `mesh.h`, :83-88).  theta / a0 / Uh: time scheme of ns_oracle.Problem (defaults: the mid-point
    scheme of stabilized_schur.py; theta = 1 mirrors stabilized_schur_bdf2.py:76-110)."""

    def __init__(self, X, Un, h, dt, rho, mu, f, eps0, theta=0.5, a0=1.0, Uh=None):
        nv = X.shape[0]
        self.nv = nv
        phi = _basis(nv)
        # geometry: x(xi, eta) = sum_a X_a phi_a ; J = dx/dxi ; K = J^-1
        x = [sum(sp.Float(X[a, i]) * phi[a] for a in range(nv)) for i in range(2)]
        J = sp.Matrix(2, 2, lambda i, j: sp.diff(x[i], (XI, ETA)[j]))
        self.detJ = sp.simplify(J.det())
        K = J.inv()

        def grad(fn):                     # UFL grad of a scalar: (d/dx_0, d/dx_1)
            return [sum(K[j, i] * sp.diff(fn, (XI, ETA)[j]) for j in range(2)) for i in range(2)]

        def nabla_grad(v):                # nabla_grad(v)[i][j] = d v_j / d x_i
            g = [grad(v[j]) for j in range(2)]
            return [[g[j][i] for j in range(2)] for i in range(2)]

        def div_vec(v):
            return sum(grad(v[i])[i] for i in range(2))

        def div_ten(A):                   # div(A)_i = d A_ij / d x_j
            return [sum(grad(A[i][j])[j] for j in range(2)) for i in range(2)]

        def sym(A):
            return [[(A[i][j] + A[j][i]) / 2 for j in range(2)] for i in range(2)]

        def dot_vv(a, b):
            return sum(a[i] * b[i] for i in range(2))

        def dot_v_ng(u, ng):              # dot(u, nabla_grad(w))_j = u_i d_i w_j
            return [sum(u[i] * ng[i][j] for i in range(2)) for j in range(2)]

        def inner_tt(A, B):
            return sum(A[i][j] * B[i][j] for i in range(2) for j in range(2))

        self._ops = dict(grad=grad, nabla_grad=nabla_grad, sym=sym)
        # unknown nodal values (symbols) and known fields
        self.Us = sp.symbols(f"U0:{2 * nv}")
        self.Ps = sp.symbols(f"P0:{nv}")
        u_sol = [sum(self.Us[2 * a + k] * phi[a] for a in range(nv)) for k in range(2)]
        p_sol = sum(self.Ps[a] * phi[a] for a in range(nv))
        u_prev = [sum(sp.Float(Un[a, k]) * phi[a] for a in range(nv)) for k in range(2)]
        Uh = Un if Uh is None else Uh
        u_hist = [sum(sp.Float(Uh[a, k]) * phi[a] for a in range(nv)) for k in range(2)]
        dt_, rho_, mu_ = sp.Float(dt), sp.Float(rho), sp.Float(mu)
        fvec = [sp.Float(f[0]), sp.Float(f[1])]
        th, a0_ = sp.Float(theta), sp.Float(a0)
        h_ = sp.Float(h)

        def epsilon(u):                   # src/solverBase.py:177-178
            return sym(nabla_grad(u))

        def sigma(u, p):                  # src/solverBase.py:180-182
            e = epsilon(u)
            return [[2 * mu_ * e[i][j] - (p if i == j else 0) for j in range(2)] for i in range(2)]

        u_mid = [th * u_sol[k] + (1 - th) * u_prev[k] for k in range(2)]            # :71 (theta = 1/2)
        dudt = [(a0_ * u_sol[k] - u_hist[k]) / dt_ for k in range(2)]                # :74 (u_sol - u_prev)/dt
        conv = dot_v_ng(u_mid, nabla_grad(u_mid))                                    # :75
        sig = sigma(u_mid, p_sol)
        # strong residual  :95-97
        dsig = div_ten(sig)
        R = [rho_ * (dudt[k] + conv[k]) - dsig[k] - rho_ * fvec[k] for k in range(2)]
        # stabilization parameters  :91-118 (functions of u_prev only)
        vnorm = sp.sqrt(dot_vv(u_prev, u_prev))
        eps = sp.Float(eps0)
        tau1 = h_ / sp.Piecewise((2 * vnorm, 2 * vnorm >= eps), (eps, True))         # :101-103
        tau2 = dt_ / 2                                                               # :104
        tau3 = (h_ * h_) / (4 * (mu_ / rho_))                                        # :105
        tau_supg = (1 / tau1 ** 2 + 1 / tau2 ** 2 + 1 / tau3 ** 2) ** sp.Rational(-1, 2)   # :106-108
        Re = (vnorm * h_) / (2 * (mu_ / rho_))                                       # :116
        z = sp.Piecewise((Re / 3, Re <= 3), (1, True))                               # :117
        tau_lsic = (vnorm * h_ * z) / 2                                              # :118

        integrands = []
        for a in range(nv):
            for k in range(2):
                v = [phi[a] if i == k else sp.Integer(0) for i in range(2)]
                Fv = rho_ * dot_vv(v, dudt)                                          # :74
                Fv += rho_ * dot_vv(v, conv)                                         # :75
                Fv -= dot_vv(v, [rho_ * fvec[0], rho_ * fvec[1]])                    # :76
                Fv += inner_tt(epsilon(v), sig)                                      # :77
                Fv += tau_supg * dot_vv(R, dot_v_ng(u_mid, nabla_grad(v)))           # :109
                Fv += tau_lsic * div_vec(u_mid) * rho_ * div_vec(v)                  # :119
                integrands.append(Fv)
        for a in range(nv):
            q = phi[a]
            Fq = q * div_vec(u_mid)                                                  # :80
            Fq += (1 / rho_) * tau_supg * dot_vv(R, grad(q))                          # :113
            integrands.append(Fq)
        args = (XI, ETA) + tuple(self.Us) + tuple(self.Ps)
        self._f = sp.lambdify(args, integrands + [sp.Abs(self.detJ)], modules="numpy", cse=True)
        # facet integrands need these as callables too
        self._u_mid, self._u_prev, self._p_sol, self._phi = u_mid, u_prev, p_sol, phi
        self._mu, self._rho, self._h = mu_, rho_, h_
        self._args = args

    def cell_residual(self, U, P, rule_u, rule_p):
        """(Fu (nv,2), Fp (nv,)) with the rules of the F_u and F_p block forms."""
        nv = self.nv
        out = np.zeros(3 * nv, dtype=np.result_type(U, P))
        for rule, sl in ((rule_u, slice(0, 2 * nv)), (rule_p, slice(2 * nv, 3 * nv))):
            pts, wts = rule
            for (xi, eta), w in zip(pts, wts):
                vals = self._f(xi, eta, *U.reshape(-1), *P)
                out[sl] += w * vals[-1] * np.array(vals[:-1], dtype=out.dtype)[sl]
        return out[:2 * nv].reshape(nv, 2), out[2 * nv:]

    def facet_residual(self, U, P, verts, nrm, length, rule, a_p=0.0, pconst=0.0, a_g=0.0, a_s=0.0, a_n=0.0, beta_n=0.0,
                       a_b=0.0, beta_b=0.0):
        """Fu (nv,2) of the ds terms on the straight facet between local vertices `verts` with unit
        outward normal `nrm`:  stabilized_schur.py:79 (a_p, a_g),
        stabilized_schur_pressure_backflow.py:192-201 (pconst, a_n), :208-209 (pconst, a_s), :214-217 (a_b)."""
        nv = self.nv
        grad, nabla_grad, sym = self._ops["grad"], self._ops["nabla_grad"], self._ops["sym"]
        n = [sp.Float(nrm[0]), sp.Float(nrm[1])]
        mu_, rho_, h_ = self._mu, self._rho, self._h
        u, un, p = self._u_mid, self._u_prev, self._p_sol
        ng = nabla_grad(u)
        eps_u = sym(ng)
        t_visc = [sum(2 * mu_ * eps_u[i][j] * n[j] for j in range(2)) for i in range(2)]     # (2 mu eps(u)) n
        un_n = sum(un[i] * n[i] for i in range(2))
        un_minus = (un_n - sp.Abs(un_n)) / 2                                                 # :215
        u_n = sum(u[i] * n[i] for i in range(2))
        u_t = [u[i] - u_n * n[i] for i in range(2)]                                          # :189
        exprs = []
        for a in range(nv):
            for k in range(2):
                v = [self._phi[a] if i == k else sp.Integer(0) for i in range(2)]
                v_n = sum(v[i] * n[i] for i in range(2))
                v_t = [v[i] - v_n * n[i] for i in range(2)]
                eps_v = sym(nabla_grad(v))
                tv = [sum(2 * mu_ * eps_v[i][j] * n[j] for j in range(2)) for i in range(2)]
                e = (a_p * p + pconst) * v_n                                                 # p n.v  /  p_c n.v
                e -= a_g * sum(mu_ * sum(ng[i][j] * n[j] for j in range(2)) * v[i] for i in range(2))   # :79
                e -= a_s * sum(t_visc[i] * v[i] for i in range(2))                           # :209
                e -= a_n * sum(t_visc[i] * v_t[i] for i in range(2))                         # :195-196
                e -= a_n * sum(tv[i] * u_t[i] for i in range(2))                             # :197-198
                e += a_n * (beta_n * mu_ / h_) * sum(u_t[i] * v_t[i] for i in range(2))      # :199-201
                e -= a_b * beta_b * rho_ * un_minus * sum(u[i] * v[i] for i in range(2))     # :216
                exprs.append(e)
        fn = sp.lambdify(self._args, exprs, modules="numpy", cse=True)
        va, vb = verts
        ref = {3: [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0)], 4: [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0)]}[nv]
        out = np.zeros(2 * nv, dtype=np.result_type(U, P))
        pts, wts = rule
        for s, w in zip(pts, wts):
            xi = (1 - s) * ref[va][0] + s * ref[vb][0]
            eta = (1 - s) * ref[va][1] + s * ref[vb][1]
            out += w * length * np.array(fn(xi, eta, *U.reshape(-1), *P), dtype=out.dtype)
        return out.reshape(nv, 2)


class SimplexForms:
    """The same literal transcription for one P1 simplex in d = 2 or 3 dimensions (cell and
    exterior-facet integrals): cross-check of `simplex_oracle` — groundwork for the tetrahedral kernels."""

    def __init__(self, X, Un, h, dt, rho, mu, f, eps0, theta=0.5, a0=1.0, Uh=None):
        nv, d = X.shape
        assert nv == d + 1
        self.nv, self.d = nv, d
        xi = sp.symbols(f"xi0:{d}", real=True)
        phi = [1 - sum(xi)] + list(xi)
        x = [sum(sp.Float(X[a, i]) * phi[a] for a in range(nv)) for i in range(d)]
        J = sp.Matrix(d, d, lambda i, j: sp.diff(x[i], xi[j]))
        detJ = J.det()
        K = J.inv()
        R_ = range(d)

        def grad(fn):
            return [sum(K[j, i] * sp.diff(fn, xi[j]) for j in R_) for i in R_]

        def nabla_grad(v):
            g = [grad(v[j]) for j in R_]
            return [[g[j][i] for j in R_] for i in R_]

        def div_vec(v):
            return sum(grad(v[i])[i] for i in R_)

        def div_ten(A):
            return [sum(grad(A[i][j])[j] for j in R_) for i in R_]

        def sym(A):
            return [[(A[i][j] + A[j][i]) / 2 for j in R_] for i in R_]

        def dot_vv(a, b):
            return sum(a[i] * b[i] for i in R_)

        def dot_v_ng(u, ng):
            return [sum(u[i] * ng[i][j] for i in R_) for j in R_]

        def inner_tt(A, B):
            return sum(A[i][j] * B[i][j] for i in R_ for j in R_)

        self.Us = sp.symbols(f"U0:{d * nv}")
        self.Ps = sp.symbols(f"P0:{nv}")
        u_sol = [sum(self.Us[d * a + k] * phi[a] for a in range(nv)) for k in R_]
        p_sol = sum(self.Ps[a] * phi[a] for a in range(nv))
        u_prev = [sum(sp.Float(Un[a, k]) * phi[a] for a in range(nv)) for k in R_]
        Uh = Un if Uh is None else Uh
        u_hist = [sum(sp.Float(Uh[a, k]) * phi[a] for a in range(nv)) for k in R_]
        dt_, rho_, mu_, h_ = sp.Float(dt), sp.Float(rho), sp.Float(mu), sp.Float(h)
        fvec = [sp.Float(v) for v in f]
        th, a0_ = sp.Float(theta), sp.Float(a0)

        def epsilon(u):                   # src/solverBase.py:177-178
            return sym(nabla_grad(u))

        def sigma(u, p):                  # src/solverBase.py:180-182
            e = epsilon(u)
            return [[2 * mu_ * e[i][j] - (p if i == j else 0) for j in R_] for i in R_]

        u_mid = [th * u_sol[k] + (1 - th) * u_prev[k] for k in R_]                  # :71
        dudt = [(a0_ * u_sol[k] - u_hist[k]) / dt_ for k in R_]                      # :74
        conv = dot_v_ng(u_mid, nabla_grad(u_mid))                                    # :75
        sig = sigma(u_mid, p_sol)
        dsig = div_ten(sig)
        R = [rho_ * (dudt[k] + conv[k]) - dsig[k] - rho_ * fvec[k] for k in R_]      # :95-97
        vnorm = sp.sqrt(dot_vv(u_prev, u_prev))                                      # :91-93
        eps = sp.Float(eps0)
        tau1 = h_ / sp.Piecewise((2 * vnorm, 2 * vnorm >= eps), (eps, True))         # :101-103
        tau_supg = (1 / tau1 ** 2 + 1 / (dt_ / 2) ** 2 + 1 / ((h_ * h_) / (4 * (mu_ / rho_))) ** 2) ** sp.Rational(-1, 2)
        Re = (vnorm * h_) / (2 * (mu_ / rho_))                                       # :116
        z = sp.Piecewise((Re / 3, Re <= 3), (1, True))                               # :117
        tau_lsic = (vnorm * h_ * z) / 2                                              # :118
        integrands = []
        for a in range(nv):
            for k in R_:
                v = [phi[a] if i == k else sp.Integer(0) for i in R_]
                Fv = rho_ * dot_vv(v, dudt) + rho_ * dot_vv(v, conv)                 # :74-75
                Fv -= dot_vv(v, [rho_ * fk for fk in fvec])                          # :76
                Fv += inner_tt(epsilon(v), sig)                                      # :77
                Fv += tau_supg * dot_vv(R, dot_v_ng(u_mid, nabla_grad(v)))           # :109
                Fv += tau_lsic * div_vec(u_mid) * rho_ * div_vec(v)                  # :119
                integrands.append(Fv)
        for a in range(nv):
            q = phi[a]
            integrands.append(q * div_vec(u_mid) + (1 / rho_) * tau_supg * dot_vv(R, grad(q)))   # :80, :113
        args = tuple(xi) + tuple(self.Us) + tuple(self.Ps)
        self._f = sp.lambdify(args, integrands + [sp.Abs(detJ)], modules="numpy", cse=True)
        self._args, self._phi = args, phi
        self._u_mid, self._u_prev, self._p_sol = u_mid, u_prev, p_sol
        self._mu, self._rho, self._h = mu_, rho_, h_
        self._nabla_grad, self._sym = nabla_grad, sym

    def facet_residual(self, U, P, lf, nrm, scale, rule, a_p=0.0, pconst=0.0, a_g=0.0, a_s=0.0, a_n=0.0, beta_n=0.0,
                       a_b=0.0, beta_b=0.0):
        """Fu (nv, d) of the ds terms on local facet `lf` (opposite vertex lf) with unit outward normal
        `nrm` and physical/reference measure ratio `scale`: stabilized_schur.py:79 (a_p, a_g),
        stabilized_schur_pressure_backflow.py:192-201 (pconst, a_n), :208-209 (pconst, a_s),
        :214-217 (a_b).  `rule`: points on the reference facet spanned by the facet's vertices in
        ascending local order."""
        nv, d = self.nv, self.d
        R_ = range(d)
        nabla_grad, sym = self._nabla_grad, self._sym
        n = [sp.Float(v) for v in nrm]
        mu_, rho_, h_ = self._mu, self._rho, self._h
        u, un, p = self._u_mid, self._u_prev, self._p_sol
        ng = nabla_grad(u)
        eps_u = sym(ng)
        t_visc = [sum(2 * mu_ * eps_u[i][j] * n[j] for j in R_) for i in R_]                 # (2 mu eps(u)) n
        un_n = sum(un[i] * n[i] for i in R_)
        un_minus = (un_n - sp.Abs(un_n)) / 2                                                 # :215
        u_n = sum(u[i] * n[i] for i in R_)
        u_t = [u[i] - u_n * n[i] for i in R_]                                                # :189
        exprs = []
        for a in range(nv):
            for k in R_:
                v = [self._phi[a] if i == k else sp.Integer(0) for i in R_]
                v_n = sum(v[i] * n[i] for i in R_)
                v_t = [v[i] - v_n * n[i] for i in R_]
                eps_v = sym(nabla_grad(v))
                tv = [sum(2 * mu_ * eps_v[i][j] * n[j] for j in R_) for i in R_]
                e = (a_p * p + pconst) * v_n                                                 # p n.v  /  p_c n.v
                e -= a_g * sum(mu_ * sum(ng[i][j] * n[j] for j in R_) * v[i] for i in R_)    # :79
                e -= a_s * sum(t_visc[i] * v[i] for i in R_)                                 # :209
                e -= a_n * sum(t_visc[i] * v_t[i] for i in R_)                               # :195-196
                e -= a_n * sum(tv[i] * u_t[i] for i in R_)                                   # :197-198
                e += a_n * (beta_n * mu_ / h_) * sum(u_t[i] * v_t[i] for i in R_)            # :199-201
                e -= a_b * beta_b * rho_ * un_minus * sum(u[i] * v[i] for i in R_)           # :216
                exprs.append(e)
        fn = sp.lambdify(self._args, exprs, modules="numpy", cse=True)
        ref = np.vstack([np.zeros((1, d)), np.eye(d)])           # reference cell vertices
        fverts = [v for v in range(nv) if v != lf]
        out = np.zeros(d * nv, dtype=np.result_type(U, P))
        pts, wts = rule
        for pt, w in zip(pts, wts):
            pt = np.atleast_1d(pt)
            lam = np.concatenate([[1.0 - pt.sum()], pt])
            xi = sum(lam[j] * ref[fverts[j]] for j in range(d))
            out += w * scale * np.array(fn(*xi, *U.reshape(-1), *P), dtype=out.dtype)
        return out.reshape(nv, d)

    def cell_residual(self, U, P, rule_u, rule_p):
        nv, d = self.nv, self.d
        out = np.zeros((d + 1) * nv, dtype=np.result_type(U, P))
        for rule, sl in ((rule_u, slice(0, d * nv)), (rule_p, slice(d * nv, (d + 1) * nv))):
            pts, wts = rule
            for pt, w in zip(pts, wts):
                vals = self._f(*pt, *U.reshape(-1), *P)
                out[sl] += w * vals[-1] * np.array(vals[:-1], dtype=out.dtype)[sl]
        return out[:d * nv].reshape(nv, d), out[d * nv:]


class CurlCurlForms:
    """Literal transcription of the cell integrals of src/solvers/stabilized_schur_pressurebc.py:85-160 (curl-curl
    viscous term, rotational convection) on one P1 simplex, d = 2 or 3: cross-check of `curlcurl_oracle`."""

    def __init__(self, X, Un, h, dt, rho, mu, f, eps0):
        nv, d = X.shape
        assert nv == d + 1
        self.nv, self.d = nv, d
        xi = sp.symbols(f"xi0:{d}", real=True)
        phi = [1 - sum(xi)] + list(xi)
        x = [sum(sp.Float(X[a, i]) * phi[a] for a in range(nv)) for i in range(d)]
        J = sp.Matrix(d, d, lambda i, j: sp.diff(x[i], xi[j]))
        detJ = J.det()
        K = J.inv()
        R_ = range(d)

        def dx(fn, i):                                    # w.dx(i)
            return sum(K[j, i] * sp.diff(fn, xi[j]) for j in R_)

        def grad(fn):
            return [dx(fn, i) for i in R_]

        def nabla_grad(v):
            return [[dx(v[j], i) for j in R_] for i in R_]

        def div(v):
            return sum(dx(v[i], i) for i in R_)

        def dot(a, b):
            return sum(a[i] * b[i] for i in R_)

        if d == 2:                                        # :96-110
            def _rot(w):
                return dx(w[1], 0) - dx(w[0], 1)

            def curl_curl_inner(u, v):
                return _rot(u) * _rot(v)

            def cross_curl_vec(w):
                om = _rot(w)
                return [-om * w[1], om * w[0]]
        else:                                             # :112-122
            def curl(w):
                return [dx(w[2], 1) - dx(w[1], 2), dx(w[0], 2) - dx(w[2], 0), dx(w[1], 0) - dx(w[0], 1)]

            def cross(a, b):
                return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

            def curl_curl_inner(u, v):
                return dot(curl(u), curl(v))

            def cross_curl_vec(w):
                return cross(curl(w), w)

        self.Us = sp.symbols(f"U0:{d * nv}")
        self.Ps = sp.symbols(f"P0:{nv}")
        u_sol = [sum(self.Us[d * a + k] * phi[a] for a in range(nv)) for k in R_]
        p_sol = sum(self.Ps[a] * phi[a] for a in range(nv))
        u_prev = [sum(sp.Float(Un[a, k]) * phi[a] for a in range(nv)) for k in R_]
        dt_, rho_, mu_, h_ = sp.Float(dt), sp.Float(rho), sp.Float(mu), sp.Float(h)
        fvec = [sp.Float(v) for v in f]
        u_mid = [(u_sol[k] + u_prev[k]) / 2 for k in R_]                               # :89
        vnorm = sp.sqrt(dot(u_prev, u_prev))                                           # :141
        Rs = [rho_ * ((u_sol[k] - u_prev[k]) / dt_ + cross_curl_vec(u_mid)[k]) for k in R_]    # :143-144
        gp = grad(p_sol)
        Rs = [Rs[k] + gp[k] - rho_ * fvec[k] for k in R_]                              # :145
        eps = sp.Float(eps0)
        tau1 = h_ / sp.Piecewise((2 * vnorm, 2 * vnorm >= eps), (eps, True))           # :148
        tau = (1 / tau1 ** 2 + 1 / (dt_ / 2) ** 2 + 1 / ((h_ * h_) / (4 * (mu_ / rho_))) ** 2) ** sp.Rational(-1, 2)
        Re = (vnorm * h_) / (2 * (mu_ / rho_))                                         # :155
        z = sp.Piecewise((Re / 3, Re <= 3), (1, True))
        tau_lsic = (vnorm * h_ * z) / 2                                                # :157
        integrands = []
        for a in range(nv):
            for k in R_:
                v = [phi[a] if i == k else sp.Integer(0) for i in R_]
                Fv = rho_ * dot(v, [(u_sol[i] - u_prev[i]) / dt_ for i in R_])         # :126
                Fv += mu_ * curl_curl_inner(u_mid, v)                                  # :127
                Fv -= p_sol * div(v)                                                   # :128
                Fv += rho_ * dot(cross_curl_vec(u_mid), v)                             # :129
                Fv -= rho_ * sp.Rational(1, 2) * dot(u_mid, u_mid) * div(v)            # :130
                Fv -= rho_ * dot(v, fvec)                                              # :131
                ngv = nabla_grad(v)
                Fv += dot([tau * r for r in Rs], [sum(u_mid[i] * ngv[i][j] for i in R_) for j in R_])   # :152
                Fv += tau_lsic * div(u_mid) * rho_ * div(v)                            # :158
                integrands.append(Fv)
        for a in range(nv):
            q = phi[a]
            integrands.append(q * div(u_mid) + (1 / rho_) * dot([tau * r for r in Rs], grad(q)))        # :132, :153
        args = tuple(xi) + tuple(self.Us) + tuple(self.Ps)
        self._f = sp.lambdify(args, integrands + [sp.Abs(detJ)], modules="numpy", cse=True)
        self._args, self._phi, self._u_mid, self._mu, self._h = args, phi, u_mid, mu_, h_
        if d == 2:                                        # _cross_curl_n, :105-107 / :121-122
            self._cross_curl_n = lambda w, n: [-_rot(w) * n[1], _rot(w) * n[0]]
        else:
            self._cross_curl_n = lambda w, n: cross(curl(w), n)

    def facet_residual(self, U, P, lf, nrm, scale, rule, pconst=0.0, a_n=0.0, beta_n=0.0):
        """Fu (nv, d) of the ds terms of stabilized_schur_pressurebc.setup on local facet `lf`: p_c dot(v, n)
        (:189-190) and the Nitsche terms (:193-201)."""
        nv, d = self.nv, self.d
        R_ = range(d)
        n = [sp.Float(v) for v in nrm]
        u, mu_, h_ = self._u_mid, self._mu, self._h
        dot = lambda a, b: sum(a[i] * b[i] for i in R_)
        u_T = [u[i] - dot(u, n) * n[i] for i in R_]                                    # :195
        exprs = []
        for a in range(nv):
            for k in R_:
                v = [self._phi[a] if i == k else sp.Integer(0) for i in R_]
                v_T = [v[i] - dot(v, n) * n[i] for i in R_]                            # :196
                e = pconst * dot(v, n)                                                 # :189-190
                e += a_n * (-mu_ * dot(self._cross_curl_n(u, n), v_T)                  # :199
                            - mu_ * dot(self._cross_curl_n(v, n), u_T)                 # :200
                            + (beta_n * mu_ / h_) * dot(u_T, v_T))                     # :201
                exprs.append(e)
        fn = sp.lambdify(self._args, exprs, modules="numpy", cse=True)
        ref = np.vstack([np.zeros((1, d)), np.eye(d)])
        fverts = [v for v in range(nv) if v != lf]
        out = np.zeros(d * nv, dtype=np.result_type(U, P))
        pts, wts = rule
        for pt, w in zip(pts, wts):
            pt = np.atleast_1d(pt)
            lam = np.concatenate([[1.0 - pt.sum()], pt])
            xi = sum(lam[j] * ref[fverts[j]] for j in range(d))
            out += w * scale * np.array(fn(*xi, *U.reshape(-1), *P), dtype=out.dtype)
        return out.reshape(nv, d)

    def cell_residual(self, U, P, rule_u, rule_p):
        nv, d = self.nv, self.d
        out = np.zeros((d + 1) * nv, dtype=np.result_type(U, P))
        for rule, sl in ((rule_u, slice(0, d * nv)), (rule_p, slice(d * nv, (d + 1) * nv))):
            pts, wts = rule
            for pt, w in zip(pts, wts):
                vals = self._f(*np.atleast_1d(pt), *U.reshape(-1), *P)
                out[sl] += w * vals[-1] * np.array(vals[:-1], dtype=out.dtype)[sl]
        return out[:d * nv].reshape(nv, d), out[d * nv:]
