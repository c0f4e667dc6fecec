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
