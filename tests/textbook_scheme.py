"""An INDEPENDENT statement of the scheme the hot path implements, written from the published
algorithms rather than from the reference's source structure:

* WENO5-JS in its textbook form (Jiang & Shu 1996: candidate polynomials of cell averages, smoothness
  indicators 13/12 (.)^2 + 1/4 (.)^2, ideal weights 1/10, 6/10, 3/10, alpha_k = d_k / (eps + beta_k)^2),
  applied to the primitive variables;
* the HLLC solver in Toro's form (star states U*_K, F = F_K + s (U*_K - U_K)) with the Davis-type wave
  speeds s_L = min(u_L - c_L, u_R - c_R), s_R = max(u_R + c_R, u_L + c_L) and the stiffened-gas mixture
  sound speed c^2 = ((Gamma + 1) p + Pi) / (rho Gamma);
* the five-equation model of Allaire et al. in quasi-conservative form,
  d(alpha_i)/dt + div(alpha_i u) = alpha_i div(u), face velocities from the Riemann solver;
* SSP-RK3 (Shu & Osher).

The reference evaluates the same scheme through grid-dependent coefficient arrays
(m_weno.fpp:168-363) that reduce to these constants on uniform grids only up to rounding, so the
oracle and this file agree to ~1e-13, not bitwise: tests/test_oracle_vs_textbook.py gates at 1e-11.
Vectorised numpy over (variable, [z,] y, x); uniform grids; boundary codes <= -3 (extrapolation),
-1 (periodic), -2 (reflective).  TEST INFRASTRUCTURE: a second pin of the oracle, used nowhere else."""
import numpy as np


def _pad(q, b, axis, code_beg, code_end, mom_normal):
    """Ghost cells along one axis of q (E, Ny, Nx)."""
    def side(code, first):
        n = q.shape[axis]
        idx = np.arange(b)
        if code <= -3:                                   # extrapolation: copy the edge cell
            src = np.zeros(b, dtype=int) if first else np.full(b, n - 1)
        elif code == -1:                                 # periodic
            src = n - b + idx if first else idx
        elif code == -2:                                 # reflective: mirror image, normal momentum changes sign
            src = b - 1 - idx if first else n - 1 - idx
        else:
            raise ValueError(code)
        g = np.take(q, src, axis=axis).copy()
        if code == -2:
            g[mom_normal] = -g[mom_normal]
        return g
    return np.concatenate([side(code_beg, True), q, side(code_end, False)], axis=axis)


def _weno5(v, axis, eps):
    """Left- and right-face values of every cell that has two neighbours on each side along `axis`."""
    v = np.moveaxis(v, axis, -1)
    a, b, c, d, e = v[..., :-4], v[..., 1:-3], v[..., 2:-2], v[..., 3:-1], v[..., 4:]      # j-2 .. j+2
    b0 = 13.0 / 12.0 * (a - 2 * b + c) ** 2 + 0.25 * (a - 4 * b + 3 * c) ** 2
    b1 = 13.0 / 12.0 * (b - 2 * c + d) ** 2 + 0.25 * (b - d) ** 2
    b2 = 13.0 / 12.0 * (c - 2 * d + e) ** 2 + 0.25 * (3 * c - 4 * d + e) ** 2

    def combine(p0, p1, p2, d0, d1, d2):
        a0, a1, a2 = d0 / (eps + b0) ** 2, d1 / (eps + b1) ** 2, d2 / (eps + b2) ** 2
        return (a0 * p0 + a1 * p1 + a2 * p2) / (a0 + a1 + a2)
    vR = combine((2 * a - 7 * b + 11 * c) / 6.0, (-b + 5 * c + 2 * d) / 6.0, (2 * c + 5 * d - e) / 6.0, 0.1, 0.6, 0.3)
    vL = combine((-a + 5 * b + 2 * c) / 6.0, (2 * b + 5 * c - d) / 6.0, (11 * c - 7 * d + 2 * e) / 6.0, 0.3, 0.6, 0.1)
    return np.moveaxis(vL, -1, axis), np.moveaxis(vR, -1, axis)


class Textbook:
    def __init__(self, nf, nd, gammas, pi_infs, dx, bc, eps=1e-16, Re=None):
        """gammas / pi_infs in the reference's convention: Gamma_i = 1/(gamma_i - 1),
        Pi_i = gamma_i pi_inf_i/(gamma_i - 1).  dx[d]: uniform cell width; bc[d] = (beg, end).
        Re[f] = (shear, bulk) Reynolds numbers of fluid f (<= 0: that fluid has none): Navier-Stokes stress
        tau = (grad u + grad u^T - 2/3 div u I)/Re_shear + div u I/Re_bulk with the mixture rule
        1/Re = sum_f alpha_f/Re_f, central differences, 1-D / 2-D (the reference's weno_Re_flux = F branch)."""
        self.nf, self.nd, self.E = nf, nd, 2 * nf + nd + 1
        self.G, self.P = np.asarray(gammas[:nf], float), np.asarray(pi_infs[:nf], float)
        self.dx, self.bc, self.eps = dx, bc, eps
        self.iRe = None
        if Re is not None and any(r > 0 for f in Re[:nf] for r in f):
            self.iRe = np.array([[1.0 / Re[f][i] if Re[f][i] > 0 else 0.0 for f in range(nf)] for i in range(2)])
            self.has = [any(Re[f][i] > 0 for f in range(nf)) for i in range(2)]

    def mixture(self, ar, al):
        rho = ar.sum(axis=0)
        Gm = np.tensordot(self.G, al, axes=1)
        Pm = np.tensordot(self.P, al, axes=1)
        return rho, Gm, Pm

    def primitive(self, q):
        nf, nd = self.nf, self.nd
        rho, Gm, Pm = self.mixture(q[:nf], q[nf + nd + 1:])
        rho = np.maximum(rho, 1e-16)
        w = q.copy()
        w[nf:nf + nd] = q[nf:nf + nd] / rho
        w[nf + nd] = (q[nf + nd] - 0.5 * (q[nf:nf + nd] * w[nf:nf + nd]).sum(axis=0) - Pm) / Gm
        return w

    def hllc(self, L, R, n):
        """Toro's HLLC for the five-equation model; n = index of the normal velocity."""
        nf, nd = self.nf, self.nd
        out = []
        for W in (L, R):
            rho, Gm, Pm = self.mixture(W[:nf], W[nf + nd + 1:])
            p, u = W[nf + nd], W[nf + n]
            En = Gm * p + Pm + 0.5 * rho * (W[nf:nf + nd] ** 2).sum(axis=0)
            c = np.sqrt(((Gm + 1.0) * p + Pm) / (rho * Gm))
            out.append((rho, p, u, En, c))
        (rL, pL, uL, EL, cL), (rR, pR, uR, ER, cR) = out
        sL = np.minimum(uL - cL, uR - cR)
        sR = np.maximum(uR + cR, uL + cL)
        sS = (pR - pL + rL * uL * (sL - uL) - rR * uR * (sR - uR)) / (rL * (sL - uL) - rR * (sR - uR))
        left = ~np.signbit(sS)                           # s_S >= +0: the left star state is upwind
        F = np.empty_like(L)
        W = np.where(left, L, R)
        rho, p, u, En = (np.where(left, x, y) for x, y in ((rL, rR), (pL, pR), (uL, uR), (EL, ER)))
        sK = np.where(left, sL, sR)
        s = np.where(left, np.minimum(0.0, sL), np.maximum(0.0, sR))     # 0 outside the fan: plain F_K
        xi = (sK - u) / (sK - sS)
        # conserved vector U_K and its star state U*_K
        U = np.concatenate([W[:nf], rho * W[nf:nf + nd], En[None], W[nf + nd + 1:]])
        Us = xi * U
        Us[nf + n] = xi * rho * sS
        Us[nf + nd] = xi * (En + (sS - u) * (rho * sS + p / (sK - u)))
        FK = u * U
        FK[nf + n] = FK[nf + n] + p
        FK[nf + nd] = u * (En + p)
        F = FK + s * (Us - U)
        uf = u + s * (xi - 1.0)                          # velocity that advects the volume fractions
        vs = W[nf:nf + nd].copy()                        # face velocity: upwind tangential components,
        vs[n] = uf                                       # contact-weighted normal component
        return F, uf, vs

    def _viscous_flux(self, F, L, R, vs, w2, d, q_shape):
        """Subtract the viscous stress of every face of direction d from the flux F (momenta) and its
        work from the energy flux.  w2: primitive variables on the grid padded by 3 in every direction."""
        nf, nd, b = self.nf, self.nd, 3
        al_L, al_R = L[nf + nd + 1:], R[nf + nd + 1:]
        iRe = [0.5 * (np.maximum(np.tensordot(self.iRe[i], al_L, axes=1), 1e-16) +
                      np.maximum(np.tensordot(self.iRe[i], al_R, axes=1), 1e-16)) for i in range(2)]
        ax = w2.ndim - 1 - d                             # axis of direction d in w2 (variable axis first)
        n = q_shape[ax]
        vel = w2[nf:nf + nd]

        def faces_of(a):                                 # cells -1 .. N along d at interior transverse indices -> pairs
            idx = [slice(None)] * a.ndim
            for dd in range(nd):
                axd = a.ndim - 1 - dd
                idx[axd] = slice(b - 1, b + n + 1) if dd == d else slice(b, a.shape[axd] - b)
            c = a[tuple(idx)]
            lo = np.take(c, np.arange(0, n + 1), axis=ax)
            hi = np.take(c, np.arange(1, n + 2), axis=ax)
            return lo, hi
        grad = [[None] * nd for _ in range(nd)]          # grad[dd][v] = d vel_v / d x_dd at the faces
        for v in range(nd):
            lo, hi = faces_of(vel[v][None])
            grad[d][v] = ((hi - lo) / self.dx[d])[0]
            for dd in range(nd):
                if dd == d:
                    continue
                axd = vel[v].ndim - 1 - dd
                central = (np.roll(vel[v], -1, axis=axd) - np.roll(vel[v], 1, axis=axd)) / (2.0 * self.dx[dd])
                lo, hi = faces_of(central[None])
                grad[dd][v] = (0.5 * (lo + hi))[0]
        div = sum(grad[v][v] for v in range(nd))
        for i in range(nd):                              # tau_{d i}
            tau = 0.0
            if self.has[0]:
                tau = tau + (grad[d][i] + grad[i][d] - (2.0 / 3.0 * div if i == d else 0.0)) * iRe[0]
            if self.has[1] and i == d:
                tau = tau + div * iRe[1]
            F[nf + i] = F[nf + i] - tau
            F[nf + nd] = F[nf + nd] - vs[i] * tau
        return F

    def rhs(self, q):
        """q: (E, [Nz,] Ny, Nx) interior cells -> dq/dt."""
        nf, nd, E, b = self.nf, self.nd, self.E, 3
        out = np.zeros_like(q)
        w2 = None
        if self.iRe is not None:                         # ghosts in every direction, x first (corners)
            q2 = q
            for d in range(nd):
                q2 = _pad(q2, b, q.ndim - 1 - d, self.bc[d][0], self.bc[d][1], nf + d)
            w2 = self.primitive(q2)
        for d in range(nd):
            axis = q.ndim - 1 - d                        # x is the last axis
            qg = _pad(q, b, axis, self.bc[d][0], self.bc[d][1], nf + d)
            w = self.primitive(qg)
            vL, vR = _weno5(w, axis, self.eps)           # cells -1 .. N+1
            n = q.shape[axis]
            sl = lambda a, lo, hi: np.take(a, np.arange(lo, hi), axis=axis)
            L, R = sl(vR, 0, n + 1), sl(vL, 1, n + 2)
            F, uf, vs = self.hllc(L, R, d)               # faces -1/2 .. N+1/2
            if w2 is not None:
                F = self._viscous_flux(F, L, R, vs, w2, d, q.shape)
            dF = sl(F, 0, n) - sl(F, 1, n + 1)
            out += dF / self.dx[d]
            al = q[nf + nd + 1:]
            out[nf + nd + 1:] += al * (sl(uf[None], 1, n + 1) - sl(uf[None], 0, n))[0] / self.dx[d]
        return out

    def step(self, q, dt):
        q1 = q + dt * self.rhs(q)
        q2 = (3.0 * q + q1 + dt * self.rhs(q1)) / 4.0
        return (q + 2.0 * q2 + 2.0 * dt * self.rhs(q2)) / 3.0
