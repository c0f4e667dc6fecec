"""An INDEPENDENT second statement of one HD Runge-Kutta substep, used to pin oracle/specter_oracle.py.

TEST INFRASTRUCTURE.  The reference cannot be built here or on the GPU box (no Fortran compiler, MPI or FFTW:
profiles/r2a_gpu_box_probe.txt), so the oracle cannot be compared with the Fortran binary.  What can be excluded is a
self-consistent restatement error: this file states the same mathematics a second time, from the numbered specification
in SURVEY.md Appendix A (each step cites the Fortran it follows), sharing NO code with the oracle and none of its
building blocks:

  * every transform is a dense matrix product with DFT matrices built from exactly reduced integer phases -- no FFT
    library, no half-spectrum tricks: the x direction is carried as the FULL Hermitian spectrum of a real line;
  * all arithmetic is numpy longdouble (80-bit extended on x86: 64-bit mantissa), so this side's own rounding is ~1e-19;
  * the pass structure is the mathematical one (3-D transforms as three separate axis products, the projection as
    written in Appendix A item 8), not the oracle's vectorised one.

tests/test_oracle.py::test_oracle_matches_independent_statement holds the oracle to this to 1e-12 on an 8 x 8 x 48
grid (rounding of the FP64 oracle itself, amplified by the cancellations of the continuation table: |dir| ~ 3.5e3).
"""
import numpy as np

LD = np.longdouble
CLD = np.clongdouble
PI = LD("3.14159265358979323846264338327950288")


def _phase_matrix(n, sign):
    """E[a, b] = exp(sign * 2 pi i a b / n) with the phase reduced exactly in integers first."""
    a = np.arange(n)
    m = (np.outer(a, a) % n).astype(LD)
    ang = 2 * PI * m / LD(n)
    return (np.cos(ang) + 1j * LD(sign) * np.sin(ang)).astype(CLD)


class Independent:
    def __init__(self, nx, ny, nz, Cz, d, Lx, Ly, Lz, tdir, ord_):
        self.nx, self.ny, self.nz, self.C, self.d, self.ord = nx, ny, nz, Cz, d, ord_
        self.nxh = nx // 2 + 1
        self.nph = nz - Cz
        self.Lz = LD(Lz)
        # Appendix A 1 (specter.fpp:683-749)
        self.Dkx, self.Dky = LD(1) / LD(Lx), LD(1) / LD(Ly)
        self.dz = LD(Lz) / LD(nz - Cz - 1)
        self.Dkz = 2 * PI / (self.dz * nz)
        self.z = self.dz * np.arange(nz).astype(LD)
        # Appendix A 2 (specter.fpp:772-789): index n/2 holds -n/2 Dk
        def kvec(n, Dk):
            k = np.empty(n, dtype=LD)
            for i in range(n):
                k[i] = LD(i) * Dk if i < n // 2 else LD(i - n) * Dk
            return k
        self.kx_full = kvec(nx, self.Dkx)            # all nx entries; the half spectrum uses 0..nx/2
        self.kx = self.kx_full[: self.nxh].copy()    # kx(nx/2+1) = -(nx/2) Dkx
        self.ky = kvec(ny, self.Dky)
        self.kz = kvec(nz, self.Dkz)
        self.Fx, self.Bx = _phase_matrix(nx, -1), _phase_matrix(nx, +1)
        self.Fy, self.By = _phase_matrix(ny, -1), _phase_matrix(ny, +1)
        self.Fz, self.Bz = _phase_matrix(nz, -1), _phase_matrix(nz, +1)
        # Appendix A 4 (fcgram_mod.f90:180-257): raw little-endian f64, Fortran order; dir = A Q^T
        A = np.fromfile(f"{tdir}/A{Cz}-{d}.dat", dtype="<f8").reshape((d, Cz)).T.astype(LD)   # A(C, d) column-major
        Q = np.fromfile(f"{tdir}/Q{d}.dat", dtype="<f8").reshape((d, d)).T.astype(LD)        # Q(d, d) column-major
        self.dir = A @ Q.T                                                                    # (C, d)

    # ---- Appendix A 4: FC-Gram continuation of the rows above the physical region (fftp.fpp:757-772) ----
    def continue_z(self, f):
        """f[..., z] with nz rows; rows nph.. are replaced."""
        n, C, d = self.nz, self.C, self.d
        out = f.copy()
        for ii in range(1, C + 1):
            acc = 0
            for jj in range(1, d + 1):
                acc = acc + self.dir[ii - 1, jj - 1] * f[..., n - C - d + jj - 1] + self.dir[C - ii, jj - 1] * f[..., d - jj]
            out[..., n - C + ii - 1] = acc
        return out

    # ---- Appendix A 3: transforms (fftp.fpp), spectral a[kx, ky, kz], real r[z, y, x] ----
    def to_real(self, a):
        """fftp3d_complex_to_real: unnormalised backward transforms; c2r in x drops Im of the kx = 0 and nx/2 entries."""
        nx, nxh = self.nx, self.nxh
        m = np.einsum("zk,ijk->ijz", self.Bz, a.astype(CLD))        # z backward
        m = np.einsum("yj,ijz->iyz", self.By, m)                    # y backward
        full = np.zeros((nx,) + m.shape[1:], dtype=CLD)             # Hermitian completion in x
        full[:nxh] = m
        full[0] = full[0].real
        full[nx // 2] = full[nx // 2].real
        for i in range(1, nx // 2):
            full[nx - i] = np.conj(m[i])
        r = np.einsum("xi,iyz->zyx", self.Bx, full)
        return r.real.astype(LD)

    def to_spectral(self, r):
        """fftp3d_real_to_complex with the continuation: x, y forward on the physical planes, rows above continued, z forward."""
        m = np.einsum("ix,zyx->iyz", self.Fx[: self.nxh], r.astype(CLD))
        m = np.einsum("jy,iyz->ijz", self.Fy, m)
        m = self.continue_z(m)
        return np.einsum("kz,ijz->ijk", self.Fz, m)

    def z_backward(self, a):
        return np.einsum("zk,ijk->ijz", self.Bz, a.astype(CLD))

    def z_forward_continued(self, m):
        return np.einsum("kz,ijz->ijk", self.Fz, self.continue_z(m.astype(CLD)))

    # ---- Appendix A 5-7 ----
    def deriv(self, a, direction):
        k = {1: self.kx[:, None, None], 2: self.ky[None, :, None], 3: self.kz[None, None, :]}[direction]
        return 1j * k * a

    def gradre(self, v):
        """(u . grad) u, pseudospec_hd.f90:245-316: products on the physical planes only, scaled 1/N^2."""
        N = LD(self.nx) * self.ny * self.nz
        u = [self.to_real(c) for c in v]
        out = []
        for c in range(3):
            acc = np.zeros_like(u[0])
            for dd in range(3):
                acc = acc + u[dd] * self.to_real(self.deriv(v[c], dd + 1))
            acc[self.nph:] = 0
            out.append(self.to_spectral(acc / (N * N)))
        return out

    def fc_filter(self, a):
        """pseudospec_hd.f90:1099-1109."""
        alpha = 16 * np.log(LD(10))
        def fac(k, n, Dk):
            return np.exp(-alpha * (2 * k / (LD(n) * Dk)) ** 100)
        return a * fac(self.kx, self.nx, self.Dkx)[:, None, None] * fac(self.ky, self.ny, self.Dky)[None, :, None] \
                 * fac(self.kz, self.nz, self.Dkz)[None, None, :]

    # ---- Appendix A 8: no-slip walls and the projection (vboundary.f90:116-148, boundary_mod.fpp:197-402) ----
    def impose_and_project(self, v, pr, o, vwall0=(0, 0), vwallL=(0, 0)):
        nz, nph = self.nz, self.nph
        top = nph - 1
        kx, ky, kz = self.kx, self.ky, self.kz
        tmp = LD(1) / LD(o) if o == self.ord else LD(o + 1) / LD(o)
        w = []
        for c, k in ((0, kx[:, None]), (1, ky[None, :])):
            m = self.z_backward(v[c]) / LD(nz)                     # (i)
            for row, wall in ((0, vwall0), (top, vwallL)):         # (ii)
                m[:, :, row] = 1j * k * pr[:, :, row] * tmp
                m[0, 0, row] = LD(self.nx) * self.ny * LD(wall[c])
            w.append(self.z_forward_continued(m))                  # (iii)
        vx, vy, vz = w[0], w[1], v[2].astype(CLD)
        kk2 = kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            dd = -1j * (kx[:, None, None] * vx + ky[None, :, None] * vy + kz[None, None, :] * vz) / kk2   # (iv)
        dd[0, 0, 0] = 0
        vx = vx - 1j * kx[:, None, None] * dd
        vy = vy - 1j * ky[None, :, None] * dd
        vz = vz - 1j * kz[None, None, :] * dd
        wz = self.z_backward(vz) / LD(nz)                          # (v)
        bc1, bc2 = wz[:, :, 0], wz[:, :, top]
        kh = np.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2)           # (vi) Neumann-Neumann harmonic correction
        phi = np.zeros((self.nxh, self.ny, nz), dtype=CLD)
        dphi = np.zeros_like(phi)
        Lz = self.Lz
        for i in range(self.nxh):
            for j in range(self.ny):
                k_h = kh[i, j]
                if k_h > 0:
                    t = 1 / (k_h * (1 - np.exp(-2 * k_h * Lz)))
                    c1 = (bc2[i, j] - bc1[i, j] * np.exp(-k_h * Lz)) * t
                    c2 = (-bc1[i, j] + bc2[i, j] * np.exp(-k_h * Lz)) * t
                    ep, em = np.exp(k_h * (self.z - Lz)), np.exp(-k_h * self.z)
                    phi[i, j] = c1 * ep + c2 * em
                    dphi[i, j] = k_h * (c1 * ep - c2 * em)
                else:
                    phi[i, j] = bc1[i, j].real * self.z
                    dphi[i, j] = bc1[i, j].real
        pnew = self.z_backward(dd) / LD(nz) + phi                  # (vii)
        ph, dph = self.z_forward_continued(phi), self.z_forward_continued(dphi)   # (viii)
        vx = vx - 1j * kx[:, None, None] * ph
        vy = vy - 1j * ky[None, :, None] * ph
        vz = vz - dph
        return [vx, vy, vz], pnew

    # ---- Appendix A 7: one substep (hd_rkstep2.f90:3-36) ----
    def rkstep2(self, v, v0, f, pr, o, dt, nu):
        nl = [self.fc_filter(c) for c in self.gradre(v)]
        kk2 = self.kx[:, None, None] ** 2 + self.ky[None, :, None] ** 2 + self.kz[None, None, :] ** 2
        new = []
        for c in range(3):
            new.append(v0[c].astype(CLD) + LD(dt) * (LD(nu) * (-kk2 * v[c]) - nl[c] + f[c]) / LD(o))
        return self.impose_and_project(new, pr.astype(CLD), o)


# =====================================================================================================================
# BOUSS and MHD: the same independent machinery (dense DFT matrices, longdouble) for the other two solvers of
# BASELINE.json.  Written from the Fortran includes and boundary modules cited at each step, not from the oracle.
# =====================================================================================================================
class IndependentSolvers(Independent):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        tdir = a[8] if len(a) > 8 else kw["tdir"]
        d = self.d
        # load_neumann_tables (fcgram_mod.f90:300-365): Q(d,d); Q<ord>n<d>.dat = dxp, Qn(d,d); neu = Q(d,:) Qn^T, last entry
        # scaled by (dz/dxp)^ord
        Q = np.fromfile(f"{tdir}/Q{d}.dat", dtype="<f8").reshape((d, d)).T.astype(LD)
        self.neu = {}
        for order in (1, 2):
            raw = np.fromfile(f"{tdir}/Q{order}n{d}.dat", dtype="<f8")
            dxp, Qn = LD(raw[0]), raw[1:1 + d * d].reshape((d, d)).T.astype(LD)
            w = np.array([sum(Q[d - 1, m] * Qn[k, m] for m in range(d)) for k in range(d)], dtype=LD)
            w[d - 1] = w[d - 1] * (self.dz / dxp) ** order
            self.neu[order] = w

    # ---- spectral operators (pseudospec_hd.f90:161-202) ----
    def curl(self, a, b, c):
        kx, ky, kz = self.kx[:, None, None], self.ky[None, :, None], self.kz[None, None, :]
        return [1j * ky * c - 1j * kz * b, 1j * kz * a - 1j * kx * c, 1j * kx * b - 1j * ky * a]

    def _products_to_spectral(self, r):
        N = LD(self.nx) * self.ny * self.nz
        r = r / (N * N)
        r[self.nph:] = 0
        return self.to_spectral(r)

    def cross(self, P, Q):
        """FFT3[P x Q]/N^2 with the products on the physical planes (prodre pseudospec_hd.f90:357-399, vector pseudospec_mhd.f90:87-102)."""
        p = [self.to_real(c) for c in P]
        q = [self.to_real(c) for c in Q]
        return [self._products_to_spectral(p[1] * q[2] - p[2] * q[1]),
                self._products_to_spectral(p[2] * q[0] - p[0] * q[2]),
                self._products_to_spectral(p[0] * q[1] - p[1] * q[0])]

    def advect(self, v, th):
        """FFT3[v . grad th]/N^2 (pseudospec_phd.f90:58-110)."""
        acc = 0
        for dd in range(3):
            acc = acc + self.to_real(v[dd]) * self.to_real(self.deriv(th, dd + 1))
        return self._products_to_spectral(acc)

    def mixed(self, a):
        """goto_domain_w_boundaries (boundary_mod.fpp:72-135): z backward, 1/nz."""
        return self.z_backward(a) / LD(self.nz)

    # ---- BOUSS (include/bouss/bouss_rkstep2.f90:3-59) ----
    def bouss_rkstep2(self, v, th, v0, th0, f, fs, pr, o, dt, nu, kappa, xmom=1.0, xtemp=1.0):
        nl = self.gradre(v)
        adv = self.advect(v, th)
        nl[2] = nl[2] - LD(xmom) * th          # buoyancy
        adv = adv - LD(xtemp) * v[2]           # heat current
        nl = [self.fc_filter(c) for c in nl]
        adv = self.fc_filter(adv)
        kk2 = self.kx[:, None, None] ** 2 + self.ky[None, :, None] ** 2 + self.kz[None, None, :] ** 2
        new = [v0[c].astype(CLD) + LD(dt) * (LD(nu) * (-kk2 * v[c]) - nl[c] + f[c]) / LD(o) for c in range(3)]
        thn = th0.astype(CLD) + LD(dt) * (LD(kappa) * (-kk2 * th) - adv + fs) / LD(o)
        new, pnew = self.impose_and_project(new, pr.astype(CLD), o)
        # s_imposebc (sboundary.f90:67-119): constant (zero) temperature at both walls
        m = self.mixed(thn)
        m[:, :, 0] = 0
        m[:, :, self.nph - 1] = 0
        thn = self.fc_filter(self.z_forward_continued(m))
        # the theta "hack" (:57-59): to real space and back, which re-continues theta from its physical values
        N = LD(self.nx) * self.ny * self.nz
        thn = self.to_spectral(self.to_real(thn) / N)
        return new, thn, pnew

    # ---- MHD with conducting walls (include/mhd/mhd_rkstep2.f90:3-84, bboundary.f90:100-189) ----
    def _neumann(self, m, order):
        d, top, w = self.d, self.nph - 1, self.neu[order]
        lo = w[d - 1] * m[:, :, 0]
        hi = w[d - 1] * m[:, :, top]
        for k in range(1, d):
            lo = lo + w[k - 1] * m[:, :, d - k]            # f(dz+1-k)        fcgram_mod.f90:461-465
            hi = hi + w[k - 1] * m[:, :, top - d + k]      # f(nz-Cz-dz+k)    :483-487
        m[:, :, 0], m[:, :, top] = lo, hi

    def a_imposebc_and_project(self, a):
        nz, top = self.nz, self.nph - 1
        kx, ky, kz = self.kx[:, None, None], self.ky[None, :, None], self.kz[None, None, :]
        ax, ay, az = [c.astype(CLD).copy() for c in a]
        az[0, 0, 0] = 0                                                    # bboundary.f90:147-149
        w = []
        for c in (ax, ay):                                                 # int_conducting_z :192-236
            m = self.mixed(c)
            m[:, :, 0] = 0
            m[:, :, top] = 0
            w.append(self.z_forward_continued(m))
        ax, ay = w
        # sol_project(ax, ay, az, ph, 0, 0, 0) (boundary_mod.fpp:197-402)
        kk2 = kx ** 2 + ky ** 2 + kz ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            dd = -1j * (kx * ax + ky * ay + kz * az) / kk2
        dd[0, 0, 0] = 0
        ax, ay, az = ax - 1j * kx * dd, ay - 1j * ky * dd, az - 1j * kz * dd
        C1 = self.z_backward(dd / LD(nz))
        bc1, bc2 = -C1[:, :, 0], -C1[:, :, top]
        kh = np.sqrt(self.kx[:, None] ** 2 + self.ky[None, :] ** 2)
        phi = np.zeros((self.nxh, self.ny, nz), dtype=CLD)
        dphi = np.zeros_like(phi)
        Lz = self.Lz
        for i in range(self.nxh):                                          # laplace_z, pure Dirichlet :499-528, 635-675
            for j in range(self.ny):
                k_h = kh[i, j]
                if i == 0 and j == 0:
                    c1, c2 = (bc2[0, 0] - bc1[0, 0]) / Lz, bc1[0, 0]
                    phi[0, 0] = c1.real * self.z + c2.real
                    dphi[0, 0] = c1.real
                else:
                    t = 1 / (1 - np.exp(-2 * k_h * Lz))
                    c1 = (bc2[i, j] - bc1[i, j] * np.exp(-k_h * Lz)) * t
                    c2 = (bc1[i, j] - bc2[i, j] * np.exp(-k_h * Lz)) * t
                    ep, em = np.exp(k_h * (self.z - Lz)), np.exp(-k_h * self.z)
                    phi[i, j] = c1 * ep + c2 * em
                    dphi[i, j] = k_h * (c1 * ep - c2 * em)
        ph = C1 + phi
        ph_hat, dph_hat = self.z_forward_continued(phi), self.z_forward_continued(dphi)
        ax, ay, az = ax - 1j * kx * ph_hat, ay - 1j * ky * ph_hat, az - dph_hat
        out = []
        for c, order in ((ax, 2), (ay, 2), (az, 1)):                       # conducting_z :239-290
            m = self.mixed(c)
            m[:, :, 0] = 0
            m[:, :, top] = 0
            self._neumann(m, order)
            out.append(self.z_forward_continued(m))
        return out, ph

    def mhd_rkstep2(self, v, a, v0, a0, f, mf, pr, o, dt, nu, mu, b0=(0.0, 0.0, 0.0)):
        N = LD(self.nx) * self.ny * self.nz
        B = self.curl(*a)
        for c in range(3):
            B[c][0, 0, 0] = LD(b0[c]) * N                                  # mhd_rkstep2.f90:11-15
        J = self.curl(*B)                                                  # written over a (:18-20)
        nl = self.cross(self.curl(*v), v)                                  # prodre: omega x v
        lor = self.cross(J, B)
        nl = [self.fc_filter(nl[c] - lor[c]) for c in range(3)]
        emf = [self.fc_filter(c) for c in self.cross(v, B)]
        kk2 = self.kx[:, None, None] ** 2 + self.ky[None, :, None] ** 2 + self.kz[None, None, :] ** 2
        vn = [v0[c].astype(CLD) + LD(dt) * (LD(nu) * (-kk2 * v[c]) - nl[c] + f[c]) / LD(o) for c in range(3)]
        an = [a0[c].astype(CLD) + LD(dt) * (-LD(mu) * J[c] + emf[c] + mf[c]) / LD(o) for c in range(3)]
        vn, pnew = self.impose_and_project(vn, pr.astype(CLD), o)
        an, ph = self.a_imposebc_and_project(an)
        return vn, an, pnew, ph


# =====================================================================================================================
# The remaining wall kinds and solvers of SURVEY 8(f) row 2: vacuum (Robin) walls of the vector potential, ROTBOUSS and
# MHDBOUSS.  Same machinery, written from bboundary.f90:100-344, boundary_mod.fpp:264-369 / 563-678,
# fcgram_mod.f90:612-635 and the rotbouss / mhdbouss includes.
# =====================================================================================================================
class IndependentWalls(IndependentSolvers):
    def _robin(self, m, wall):
        """robin_reconstruct (fcgram_mod.f90:612-635) with a = khom: f' + khom f = g, g stored in the wall row."""
        d, top, w = self.d, self.nph - 1, self.neu[1]
        kh = np.sqrt(self.kx[:, None] ** 2 + self.ky[None, :] ** 2)
        if wall == 0:
            acc = w[d - 1] * m[:, :, 0]
            for k in range(1, d):
                acc = acc + w[k - 1] * m[:, :, d - k]
            m[:, :, 0] = acc / (kh * w[d - 1] + 1)
        else:
            acc = w[d - 1] * m[:, :, top]
            for k in range(1, d):
                acc = acc + w[k - 1] * m[:, :, top - d + k]
            m[:, :, top] = acc / (kh * w[d - 1] + 1)

    def _neumann_wall(self, m, order, wall):
        d, top, w = self.d, self.nph - 1, self.neu[order]
        if wall == 0:
            acc = w[d - 1] * m[:, :, 0]
            for k in range(1, d):
                acc = acc + w[k - 1] * m[:, :, d - k]
            m[:, :, 0] = acc
        else:
            acc = w[d - 1] * m[:, :, top]
            for k in range(1, d):
                acc = acc + w[k - 1] * m[:, :, top - d + k]
            m[:, :, top] = acc

    def a_imposebc_and_project_walls(self, a, s, e):
        """bboundary.f90:100-189 for wall kinds s (z = 0) and e (z = Lz): 0 conducting, 1 vacuum."""
        nz, top = self.nz, self.nph - 1
        kx, ky, kz = self.kx[:, None, None], self.ky[None, :, None], self.kz[None, None, :]
        ax, ay, az = [c.astype(CLD).copy() for c in a]
        az[0, 0, 0] = 0
        if s == 0 or e == 0:                                               # int_conducting_z on the conducting walls only
            w = []
            for c in (ax, ay):
                m = self.mixed(c)
                if s == 0:
                    m[:, :, 0] = 0
                if e == 0:
                    m[:, :, top] = 0
                w.append(self.z_forward_continued(m))
            ax, ay = w
        # sol_project(ax, ay, az, ph, 0, 2 s, 2 e)
        kk2 = kx ** 2 + ky ** 2 + kz ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            dd = -1j * (kx * ax + ky * ay + kz * az) / kk2
        dd[0, 0, 0] = 0
        ax, ay, az = ax - 1j * kx * dd, ay - 1j * ky * dd, az - 1j * kz * dd
        C1s = dd / LD(nz)
        C1 = self.z_backward(C1s)
        C2 = self.z_backward(1j * kz * C1s)                                # derivk(C1, C2, 3), then to z (:289-292)
        kh = np.sqrt(self.kx[:, None] ** 2 + self.ky[None, :] ** 2)
        bc1 = -C1[:, :, 0] if s == 0 else C2[:, :, 0] - kh * C1[:, :, 0]                    # :294-321
        bc2 = -C1[:, :, top] if e == 0 else -(C2[:, :, top] + kh * C1[:, :, top])           # :323-352
        Lz = self.Lz
        phi = np.zeros((self.nxh, self.ny, nz), dtype=CLD)
        dphi = np.zeros_like(phi)
        for i in range(self.nxh):
            for j in range(self.ny):
                k_h, b1, b2 = kh[i, j], bc1[i, j], bc2[i, j]
                if (s, e) == (0, 0):                                       # pure Dirichlet :499-528
                    mean = ((b2 - b1) / Lz, b1)
                    if k_h > 0:
                        t = 1 / (1 - np.exp(-2 * k_h * Lz))
                        c1, c2 = (b2 - b1 * np.exp(-k_h * Lz)) * t, (b1 - b2 * np.exp(-k_h * Lz)) * t
                elif (s, e) == (1, 1):                                     # pure Robin :563-592
                    mean = (b1, 0)
                    if k_h > 0:
                        c1, c2 = b2 / (2 * k_h), b1 / (2 * k_h)
                elif (s, e) == (0, 1):                                     # Dirichlet bottom, Robin top :595-624
                    mean = (b2, b1)
                    if k_h > 0:
                        c1 = b2 / (2 * k_h)
                        c2 = (b1 * 2 * k_h - b2 * np.exp(-k_h * Lz)) / (2 * k_h)
                else:
                    raise ValueError("[ERROR] Unsupported BC combination in call to laplace_z. Aborting...")
                if i == 0 and j == 0:                                      # :635-639
                    phi[0, 0] = np.real(mean[0]) * self.z + np.real(mean[1])
                    dphi[0, 0] = np.real(mean[0])
                else:
                    ep, em = np.exp(k_h * (self.z - Lz)), np.exp(-k_h * self.z)
                    phi[i, j] = c1 * ep + c2 * em
                    dphi[i, j] = k_h * (c1 * ep - c2 * em)
        ph = C1 + phi
        ph_hat, dph_hat = self.z_forward_continued(phi), self.z_forward_continued(dphi)
        ax, ay, az = ax - 1j * kx * ph_hat, ay - 1j * ky * ph_hat, az - dph_hat
        out = []
        for c, order in ((ax, 2), (ay, 2), (az, 1)):
            m = self.mixed(c)
            for wall, kind in ((0, s), (1, e)):                            # z = 0 first, then z = Lz (:168-181)
                m[:, :, 0 if wall == 0 else top] = 0
                if kind == 0:
                    self._neumann_wall(m, order, wall)                     # conducting_z :239-290
                else:
                    self._robin(m, wall)                                   # insulating_z :294-344
            out.append(self.z_forward_continued(m))
        return out, ph

    # ---- ROTBOUSS (include/rotbouss/rotbouss_rkstep2.f90:3-56): Coriolis term, no theta filter / round trip ----
    def rotbouss_rkstep2(self, v, th, v0, th0, f, fs, pr, o, dt, nu, kappa, xmom, xtemp, omega, vwall0=(0, 0), vwallL=(0, 0)):
        ox, oy, oz = [LD(w) for w in omega]
        nl = self.gradre(v)
        adv = self.advect(v, th)
        nl[0] = nl[0] + 2 * (oy * v[2] - oz * v[1])
        nl[1] = nl[1] + 2 * (oz * v[0] - ox * v[2])
        nl[2] = nl[2] + 2 * (ox * v[1] - oy * v[0]) - LD(xmom) * th
        adv = adv - LD(xtemp) * v[2]
        nl = [self.fc_filter(c) for c in nl]
        adv = self.fc_filter(adv)
        kk2 = self.kx[:, None, None] ** 2 + self.ky[None, :, None] ** 2 + self.kz[None, None, :] ** 2
        new = [v0[c].astype(CLD) + LD(dt) * (LD(nu) * (-kk2 * v[c]) - nl[c] + f[c]) / LD(o) for c in range(3)]
        thn = th0.astype(CLD) + LD(dt) * (LD(kappa) * (-kk2 * th) - adv + fs) / LD(o)
        new, pnew = self.impose_and_project(new, pr.astype(CLD), o, vwall0, vwallL)
        m = self.mixed(thn)                                                # s_imposebc: zero walls, forward (no filter here)
        m[:, :, 0] = 0
        m[:, :, self.nph - 1] = 0
        thn = self.z_forward_continued(m)
        return new, thn, pnew

    # ---- MHDBOUSS (include/mhdbouss/mhdbouss_rkstep2.f90:3-106) ----
    def mhdbouss_rkstep2(self, v, a, th, v0, a0, th0, f, mf, fs, pr, o, dt, nu, mu, kappa, xmom, xtemp, b0, s, e):
        N = LD(self.nx) * self.ny * self.nz
        B = self.curl(*a)
        for c in range(3):
            B[c][0, 0, 0] = LD(b0[c]) * N
        J = self.curl(*B)
        nl = self.cross(self.curl(*v), v)
        adv = self.advect(v, th)
        lor = self.cross(J, B)
        nl = [nl[c] - lor[c] for c in range(3)]
        nl[2] = nl[2] - LD(xmom) * th
        adv = adv - LD(xtemp) * v[2]
        nl = [self.fc_filter(c) for c in nl]
        adv = self.fc_filter(adv)
        emf = [self.fc_filter(c) for c in self.cross(v, B)]
        kk2 = self.kx[:, None, None] ** 2 + self.ky[None, :, None] ** 2 + self.kz[None, None, :] ** 2
        vn = [v0[c].astype(CLD) + LD(dt) * (LD(nu) * (-kk2 * v[c]) - nl[c] + f[c]) / LD(o) for c in range(3)]
        an = [a0[c].astype(CLD) + LD(dt) * (-LD(mu) * J[c] + emf[c] + mf[c]) / LD(o) for c in range(3)]
        thn = th0.astype(CLD) + LD(dt) * (LD(kappa) * (-kk2 * th) - adv + fs) / LD(o)
        vn, pnew = self.impose_and_project(vn, pr.astype(CLD), o)
        an, ph = self.a_imposebc_and_project_walls(an, s, e)
        m = self.mixed(thn)
        m[:, :, 0] = 0
        m[:, :, self.nph - 1] = 0
        thn = self.z_forward_continued(m)
        thn = self.to_spectral(self.to_real(thn) / N)                      # the round trip without the filter (:103-106)
        return vn, an, thn, pnew, ph
