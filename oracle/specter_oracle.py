"""CPU oracle for the SPECTER per-RK-substep hot path (numpy / scipy.fft, FP64).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / reported baseline.
The product path (``specter_b200``) never imports anything from ``oracle/``.

It is a literal restatement -- same pass structure, same (non)normalisation,
same quirks -- of the reference's Fortran for this path.  Every function cites
the reference ``file:line`` (relative to /root/reference/src) it follows.

PARITY PINNING: "parity unpinned" against the reference binary.  The reference
(Fortran + MPI + FFTW) cannot be compiled in this image nor on the GPU box (no
gfortran / MPI / FFTW on either: DESIGN.md 9, profiles/r2a_gpu_box_probe.txt)
and ships no golden vectors: its own tests (src/tests/*.f90) only *print*
error norms of analytic identities.  The oracle is therefore pinned against
(i) those analytic known-answer identities re-stated as assertions
(tests/test_oracle.py), (ii) the reference's FC-Gram table fixtures, and
(iii) INDEPENDENT second statements of the HD, BOUSS, MHD, ROTBOUSS (moving
walls) and MHDBOUSS (conducting, vacuum and mixed walls) substeps
(tests/independent_hd.py: dense DFT matrices, full Hermitian x spectrum,
longdouble; no shared code), which it matches to better than 1e-12 of the field
maxima over two substeps each, and (iv) real-space evaluations of every global
quantity (energy, enstrophy-like column, helicity, cross, divergence, variance,
product, wall checks) with the same independent transform.  FFTW is a third-party
dependency absent from /root/reference (only hint of a version: 3.3.8,
src/Makefile.in:59); the DFT is mathematically fixed (forward sign -1, backward
+1, both unnormalised), so bitwise parity with an FFTW build is unpinned only
beyond FFT rounding.

Array conventions (memory-identical to the Fortran arrays):
  spectral / mixed  Fortran ``a(nz,ny,ista:iend)``  <->  numpy ``a[i,j,k]`` shape (nxl,ny,nz), C order
  real              Fortran ``r(nx,ny,ksta:kend)``  <->  numpy ``r[k,j,i]`` shape (nzl,ny,nx), C order
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft

IM = 1j
_WORKERS = int(os.environ.get("SPECTER_ORACLE_WORKERS", "0")) or (os.cpu_count() or 1)


def set_workers(n: int) -> None:
    global _WORKERS
    _WORKERS = max(1, int(n))


# ----------------------------------------------------------------------------
# slab partition                                             fftp/fftp.fpp:1154-1184
# ----------------------------------------------------------------------------
def range_(n1: int, n2: int, nprocs: int, irank: int):
    """``range`` (fftp.fpp:1177-1181): 1-based inclusive [sta,end] of rank irank."""
    iwork1 = (n2 - n1 + 1) // nprocs
    iwork2 = (n2 - n1 + 1) % nprocs
    ista = irank * iwork1 + n1 + min(irank, iwork2)
    iend = ista + iwork1 - 1
    if iwork2 > irank:
        iend += 1
    return ista, iend


# ----------------------------------------------------------------------------
# FC-Gram tables                                     fftp/fcgram_mod.f90:180-365
# ----------------------------------------------------------------------------
def load_dirichlet_tables(tdir: str, C: int, d: int) -> np.ndarray:
    """``load_dirichlet_tables`` (fcgram_mod.f90:228-252): dir = A . Q^T, shape (C,d).

    Files are raw little-endian f64 streams in Fortran (column-major) order.
    """
    A = np.fromfile(os.path.join(tdir, f"A{C}-{d}.dat"), dtype="<f8")
    Q = np.fromfile(os.path.join(tdir, f"Q{d}.dat"), dtype="<f8")
    if A.size != C * d or Q.size != d * d:
        raise ValueError("FC-Gram table size mismatch")
    A = A.reshape(d, C).T  # A(C,d) column-major
    Q = Q.reshape(d, d).T
    return np.ascontiguousarray(A @ Q.T)


def load_neumann_tables(tdir: str, d: int, dz: float, order: int) -> np.ndarray:
    """``load_neumann_tables`` (fcgram_mod.f90:261-365): neu = Q(d,:) . Qn^T,
    last entry scaled by (dz/dxp)^order.  order=1 -> Q1n, order=2 -> Q2n."""
    Q = np.fromfile(os.path.join(tdir, f"Q{d}.dat"), dtype="<f8").reshape(d, d).T
    raw = np.fromfile(os.path.join(tdir, f"Q{order}n{d}.dat"), dtype="<f8")
    dxp = raw[0]
    Qn = raw[1:].reshape(d, d).T
    neu = Qn @ Q[d - 1, :]  # neu(k) = sum_j Q(d,j) Qn(k,j)   (fcgram_mod.f90:355-358)
    neu = np.array(neu, dtype=np.float64)
    neu[d - 1] *= (dz / dxp) ** order
    return neu


# ----------------------------------------------------------------------------
# grids, wavenumbers                                      specter.fpp:683-800
# ----------------------------------------------------------------------------
@dataclass
class Grid:
    nx: int
    ny: int
    nz: int
    Cz: int
    oz: int
    Lx: float = 1.0
    Ly: float = 1.0
    Lz: float = 1.0
    tdir: str = ""
    nprocs: int = 1
    myrank: int = 0
    ord: int = 2
    x: np.ndarray = field(init=False, repr=False)
    y: np.ndarray = field(init=False, repr=False)
    z: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        nx, ny, nz, Cz = self.nx, self.ny, self.nz, self.Cz
        pi = np.pi
        # periodic x,y (specter.fpp:683-724); non-periodic z (:742-749)
        self.dx = self.Lx * 2.0 * pi / nx
        self.Dkx = 1.0 / self.Lx
        self.dy = self.Ly * 2.0 * pi / ny
        self.Dky = 1.0 / self.Ly
        if Cz == 0:
            self.dz = self.Lz * 2.0 * pi / nz
            self.Dkz = 1.0 / self.Lz
        else:
            self.dz = self.Lz / (nz - Cz - 1)
            self.Dkz = 2.0 * pi / (self.dz * nz)
        self.x = self.dx * np.arange(nx, dtype=np.float64)
        self.y = self.dy * np.arange(ny, dtype=np.float64)
        self.z = self.dz * np.arange(nz, dtype=np.float64)
        # wavenumbers (specter.fpp:772-789): index n/2+1 holds -n/2
        self.kx_full = self._kvec(nx) * self.Dkx
        self.ky = self._kvec(ny) * self.Dky
        self.kz = self._kvec(nz) * self.Dkz
        self.nxh = nx // 2 + 1
        self.ista, self.iend = range_(1, self.nxh, self.nprocs, self.myrank)
        self.ksta, self.kend = range_(1, nz, self.nprocs, self.myrank)
        self.pkend = min(nz - Cz, self.kend)  # specter.fpp:750
        self.nxl = self.iend - self.ista + 1
        self.nzl = self.kend - self.ksta + 1
        self.kx = self.kx_full[self.ista - 1 : self.iend]  # local kx(ista:iend)
        # kk2, khom (specter.fpp:791-800)
        self.kk2 = (
            self.kx[:, None, None] ** 2 + self.ky[None, :, None] ** 2 + self.kz[None, None, :] ** 2
        )
        self.khom = np.sqrt(self.kk2[:, :, 0])
        self.N = float(nx) * float(ny) * float(nz)
        if self.tdir and Cz > 0:
            self.dir = load_dirichlet_tables(self.tdir, Cz, self.oz)
        else:
            self.dir = None
        self.neu = None
        self.neu2 = None

    @staticmethod
    def _kvec(n: int) -> np.ndarray:
        k = np.empty(n, dtype=np.float64)
        if n == 1:
            k[0] = 0.0
            return k
        for i in range(1, n // 2 + 1):
            k[i - 1] = float(i - 1)
            k[i + n // 2 - 1] = float(i - n // 2 - 1)
        return k

    def load_neumann(self):
        self.neu = load_neumann_tables(self.tdir, self.oz, self.dz, 1)
        self.neu2 = load_neumann_tables(self.tdir, self.oz, self.dz, 2)

    # shapes
    def cshape(self):
        return (self.nxl, self.ny, self.nz)

    def rshape(self):
        return (self.nzl, self.ny, self.nx)


# ----------------------------------------------------------------------------
# transforms                                                   fftp/fftp.fpp
# ----------------------------------------------------------------------------
def fc_continue_z(g: Grid, a: np.ndarray) -> None:
    """FC-Gram continuation, in place on the last axis (fftp.fpp:757-772).

    f(n-C+ii) = sum_jj dir(ii,jj) f(n-C-d+jj) + dir(C-ii+1,jj) f(d-jj+1), summed
    in the reference's order jj=1..d.
    """
    n, C, d = g.nz, g.Cz, g.oz
    if C <= 0:
        return
    dirm = g.dir  # (C,d)
    dflip = dirm[::-1, :]  # dir(C-ii+1, jj)
    # jj = 1
    acc = dirm[:, 0] * a[..., n - C - d : n - C - d + 1] + dflip[:, 0] * a[..., d - 1 : d]
    for jj in range(2, d + 1):
        acc = acc + dirm[:, jj - 1] * a[..., n - C - d + jj - 1 : n - C - d + jj] \
                  + dflip[:, jj - 1] * a[..., d - jj : d - jj + 1]
    a[..., n - C :] = acc


def fftp1d_real_to_complex_z(g: Grid, a: np.ndarray) -> np.ndarray:
    """fftp.fpp:720-786: continuation + unnormalised forward c2c along z, in place."""
    fc_continue_z(g, a)
    a[...] = sfft.fft(a, axis=-1, norm="backward", workers=_WORKERS)
    return a


def fftp1d_complex_to_real_z(g: Grid, a: np.ndarray) -> np.ndarray:
    """fftp.fpp:1060-1094: unnormalised backward c2c along z, in place."""
    a[...] = sfft.ifft(a, axis=-1, norm="forward", workers=_WORKERS)
    return a


def fftp2d_real_to_complex_xy(g: Grid, r: np.ndarray) -> np.ndarray:
    """fftp.fpp:428-524 (single rank): 2-D r2c over (x,y) per z-plane, then the
    transpose (x,y,z)->(z,y,x).  r[k,j,i] -> out[i,j,k]."""
    c = sfft.rfft(r, axis=2, norm="backward", workers=_WORKERS)
    c = sfft.fft(c, axis=1, norm="backward", workers=_WORKERS)
    return np.ascontiguousarray(c.transpose(2, 1, 0))


def fftp2d_complex_to_real_xy(g: Grid, a: np.ndarray) -> np.ndarray:
    """fftp.fpp:824-919 (single rank): transpose, then 2-D c2r (FFTW semantics:
    complex backward along y, then Hermitian c2r along x which ignores the
    imaginary parts of the kx=0 and kx=nx/2 entries).  a[i,j,k] -> r[k,j,i]."""
    c = a.transpose(2, 1, 0)
    c = sfft.ifft(c, axis=1, norm="forward", workers=_WORKERS)
    return sfft.irfft(c, n=g.nx, axis=2, norm="forward", workers=_WORKERS)


def fftp3d_real_to_complex(g: Grid, r: np.ndarray) -> np.ndarray:
    """fftp.fpp:388-425 (y periodic branch)."""
    out = fftp2d_real_to_complex_xy(g, r)
    return fftp1d_real_to_complex_z(g, out)


def fftp3d_complex_to_real(g: Grid, a: np.ndarray) -> np.ndarray:
    """fftp.fpp:789-821.  (The reference destroys its input; this does not.)"""
    c = a.copy()
    fftp1d_complex_to_real_z(g, c)
    return fftp2d_complex_to_real_xy(g, c)


# ----------------------------------------------------------------------------
# spectral operators                               pseudo/pseudospec_hd.f90
# ----------------------------------------------------------------------------
def derivk(g: Grid, a: np.ndarray, dir: int) -> np.ndarray:
    """pseudospec_hd.f90:28-94: b = im*k_dir*a."""
    if dir == 1:
        return (IM * g.kx)[:, None, None] * a
    if dir == 2:
        return (IM * g.ky)[None, :, None] * a
    return (IM * g.kz)[None, None, :] * a


def laplak(g: Grid, a: np.ndarray) -> np.ndarray:
    """pseudospec_hd.f90:97-127: b = -kk2*a."""
    return -g.kk2 * a


def curlk(g: Grid, a: np.ndarray, b: np.ndarray, dir: int) -> np.ndarray:
    """pseudospec_hd.f90:130-206."""
    if dir == 1:
        c1 = derivk(g, a, 3)
        c2 = derivk(g, b, 2)
        return c2 - c1
    if dir == 2:
        c1 = derivk(g, a, 3)
        c2 = derivk(g, b, 1)
        return c1 - c2
    c1 = derivk(g, a, 2)
    c2 = derivk(g, b, 1)
    return c2 - c1


def fc_filter_factors(g: Grid):
    """The three separable factors of ``fc_filter`` (pseudospec_hd.f90:1099-1109)."""
    alpha = 16.0 * np.log(10.0)
    p2 = 100.0  # 2*p, p=50d0
    fx = np.exp(-alpha * (2 * g.kx / g.nx / g.Dkx) ** p2)
    fy = np.exp(-alpha * (2 * g.ky / g.ny / g.Dky) ** p2)
    fz = np.exp(-alpha * (2 * g.kz / g.nz / g.Dkz) ** p2)
    return fx, fy, fz


def fc_filter(g: Grid, a: np.ndarray) -> np.ndarray:
    """pseudospec_hd.f90:1082-1115: a*fx*fy*fz (left-to-right), in place."""
    fx, fy, fz = fc_filter_factors(g)
    a *= fx[:, None, None]
    a *= fy[None, :, None]
    a *= fz[None, None, :]
    return a


def _phys(g: Grid):
    """Local physical z-planes ksta..pkend as a slice of the local real array."""
    return slice(0, g.pkend - g.ksta + 1)


def gradre(g: Grid, a, b, c):
    """(A.grad)A, pseudospec_hd.f90:209-319.  Products only on k<=pkend; the other
    planes of rx,ry,rz are uninitialised in the reference (we use 0) and are
    overwritten by the continuation in the forward transform."""
    ph = _phys(g)
    rx = np.zeros(g.rshape())
    ry = np.zeros(g.rshape())
    rz = np.zeros(g.rshape())
    comps = (a, b, c)
    for d_ in (1, 2, 3):
        r1 = fftp3d_complex_to_real(g, comps[d_ - 1])
        r2 = fftp3d_complex_to_real(g, derivk(g, a, d_))
        r3 = fftp3d_complex_to_real(g, derivk(g, b, d_))
        r4 = fftp3d_complex_to_real(g, derivk(g, c, d_))
        rx[ph] += r1[ph] * r2[ph]
        ry[ph] += r1[ph] * r3[ph]
        rz[ph] += r1[ph] * r4[ph]
    tmp = 1.0 / g.N ** 2
    rx[ph] *= tmp
    ry[ph] *= tmp
    rz[ph] *= tmp
    return (fftp3d_real_to_complex(g, rx), fftp3d_real_to_complex(g, ry),
            fftp3d_real_to_complex(g, rz))


def prodre(g: Grid, a, b, c):
    """curl(A) x A, pseudospec_hd.f90:322-402."""
    ph = _phys(g)
    r1 = fftp3d_complex_to_real(g, curlk(g, b, c, 1))
    r2 = fftp3d_complex_to_real(g, curlk(g, a, c, 2))
    r3 = fftp3d_complex_to_real(g, curlk(g, a, b, 3))
    r4 = fftp3d_complex_to_real(g, a)
    r5 = fftp3d_complex_to_real(g, b)
    r6 = fftp3d_complex_to_real(g, c)
    tmp = 1.0 / g.N ** 2
    rx = np.zeros(g.rshape()); ry = np.zeros(g.rshape()); rz = np.zeros(g.rshape())
    rx[ph] = (r2[ph] * r6[ph] - r5[ph] * r3[ph]) * tmp
    ry[ph] = (r3[ph] * r4[ph] - r6[ph] * r1[ph]) * tmp
    rz[ph] = (r1[ph] * r5[ph] - r4[ph] * r2[ph]) * tmp
    return (fftp3d_real_to_complex(g, rx), fftp3d_real_to_complex(g, ry),
            fftp3d_real_to_complex(g, rz))


# ----------------------------------------------------------------------------
# diagnostics                      pseudospec_hd.f90:405-635, 778-940, 1118-1235
# ----------------------------------------------------------------------------
def _weights(g: Grid) -> np.ndarray:
    w = np.full(g.nxl, 2.0)
    if g.ista == 1:
        w[0] = 1.0
    return w


def _mean_phys(g: Grid, R1: np.ndarray, tmp: float) -> float:
    """Weighted sum over kx (1 for kx=0 else 2), ky and physical z rows."""
    nph = g.nz - g.Cz
    return float(np.sum(_weights(g)[:, None, None] * R1[:, :, :nph] * tmp))


def _abs2_iz(g: Grid, c: np.ndarray) -> np.ndarray:
    c1 = c.copy()
    fftp1d_complex_to_real_z(g, c1)
    return c1.real ** 2 + c1.imag ** 2


def energy(g: Grid, a, b, c, kin: int) -> float:
    """pseudospec_hd.f90:405-635.  kin=0 reproduces the reference's quirk: the
    y-component of the curl is added twice and the z-component never (:527,:540)."""
    tmp = 1.0 / g.N ** 2 / float(g.nz - g.Cz)
    if kin == 1:
        R1 = _abs2_iz(g, a) + _abs2_iz(g, b) + _abs2_iz(g, c)
    elif kin == 0:
        R1 = _abs2_iz(g, curlk(g, b, c, 1))
        R1 = R1 + _abs2_iz(g, curlk(g, a, c, 2))
        R1 = R1 + _abs2_iz(g, curlk(g, a, c, 2))
    else:
        C1 = curlk(g, b, c, 1)
        C2 = curlk(g, a, c, 2)
        C3 = curlk(g, a, b, 3)
        C4 = curlk(g, C2, C3, 1)
        C3n = curlk(g, C1, C3, 2)
        C1n = curlk(g, C1, C2, 3)
        R1 = _abs2_iz(g, C4) + _abs2_iz(g, C3n) + _abs2_iz(g, C1n)
    return _mean_phys(g, R1, tmp)


def cross(g: Grid, a, b, c, d, e, f, kin: int) -> float:
    """pseudospec_hd.f90:778-940 (kin=1: <A.B>; kin=0: <curl A . curl B>)."""
    tmp = 1.0 / g.N ** 2 / float(g.nz - g.Cz)

    def prod(p, q):
        c1 = p.copy(); c2 = q.copy()
        fftp1d_complex_to_real_z(g, c1); fftp1d_complex_to_real_z(g, c2)
        return (c1 * np.conj(c2)).real

    if kin == 1:
        r1 = prod(a, d) + prod(b, e) + prod(c, f)
    else:
        r1 = prod(curlk(g, b, c, 1), curlk(g, e, f, 1))
        r1 = r1 + prod(curlk(g, a, c, 2), curlk(g, d, f, 2))
        r1 = r1 + prod(curlk(g, a, b, 3), curlk(g, d, e, 3))
    return _mean_phys(g, r1, tmp)


def divergence(g: Grid, a, b, c) -> float:
    """pseudospec_hd.f90:1118-1235."""
    tmp = 1.0 / g.N ** 2 / float(g.nz - g.Cz)
    C2 = derivk(g, a, 1)
    C2 = C2 + derivk(g, b, 2)
    C2 = C2 + derivk(g, c, 3)
    return _mean_phys(g, _abs2_iz(g, C2), tmp)


def bouncheck_z(g: Grid, a, b=None):
    """boundary_mod.fpp:681-801: mean |a|^2 (+|b|^2) on rows z=0 and z=Lz."""
    tmp = (1.0 / g.N) ** 2
    top_row = g.nz - g.Cz - 1
    A2 = _abs2_iz(g, a)
    R1 = A2[:, :, 0]
    R2 = A2[:, :, top_row]
    if b is not None:
        B2 = _abs2_iz(g, b)
        R1 = R1 + B2[:, :, 0]
        R2 = R2 + B2[:, :, top_row]
    w = _weights(g)[:, None]
    return float(np.sum(w * R1 * tmp)), float(np.sum(w * R2 * tmp))


def hdcheck(g: Grid, a, b, c, d, e, f):
    """pseudospec_hd.f90:943-1005 -> the balance.txt columns (eng, ens, pot)."""
    eng = energy(g, a, b, c, 1)
    ens = energy(g, a, b, c, 0)
    pot = cross(g, a, b, c, d, e, f, 1)
    return eng, ens, pot


def vdiagnostic(g: Grid, a, b, c):
    """vboundary.f90:214-269 -> noslip_diagnostic.txt columns."""
    tm1 = divergence(g, a, b, c)
    tmp, tmq = bouncheck_z(g, a, b)
    tmr, tms = bouncheck_z(g, c)
    return tm1, tmp, tmq, tmr, tms


def normvec(g: Grid, a, b, c, d: float, kin: int):
    """pseudospec_hd.f90:1238-1286."""
    rmp = np.sqrt(d / energy(g, a, b, c, kin))
    a *= rmp; b *= rmp; c *= rmp


# ----------------------------------------------------------------------------
# boundary module                                    boundary/boundary_mod.fpp
# ----------------------------------------------------------------------------
def goto_domain_w_boundaries(g: Grid, *fields):
    """boundary_mod.fpp:72-150: z-IFFT, then x 1/nz on physical rows only."""
    tmp = 1.0 / float(g.nz)
    nph = g.nz - g.Cz
    for a in fields:
        fftp1d_complex_to_real_z(g, a)
        a[:, :, :nph] *= tmp


def goto_3d_fourier(g: Grid, *fields):
    """boundary_mod.fpp:153-194."""
    for a in fields:
        fftp1d_real_to_complex_z(g, a)


def poisson_inhomogeneous(g: Grid, a, b, c) -> np.ndarray:
    """boundary_mod.fpp:405-448.  The 0/0 at mode (1,1,1) is overwritten by 0."""
    num = (g.kx[:, None, None] * a + g.ky[None, :, None] * b + g.kz[None, None, :] * c)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = -IM * num / g.kk2
    if g.ista == 1:
        d[0, 0, 0] = 0.0
    return d


def laplace_z(g: Grid, bc: np.ndarray, bczsta: int, bczend: int):
    """boundary_mod.fpp:451-678.  bc[i,j,0:2] -> a[i,j,k], b[i,j,k] for all nz rows."""
    Lz = g.Lz
    kh = g.khom
    bc1 = bc[:, :, 0]
    bc2 = bc[:, :, 1]
    coef1 = np.empty_like(bc1)
    coef2 = np.empty_like(bc1)
    with np.errstate(divide="ignore", invalid="ignore"):
        e1 = np.exp(-kh * Lz)
        if bczsta == bczend and bczsta == 0:  # pure Dirichlet  (:499-528)
            tmp = 1.0 / (1 - np.exp(-2 * kh * Lz))
            coef1[...] = (bc2 - bc1 * e1) * tmp
            coef2[...] = (bc1 - bc2 * e1) * tmp
            if g.ista == 1:
                coef1[0, 0] = (bc2[0, 0] - bc1[0, 0]) / Lz
                coef2[0, 0] = bc1[0, 0]
        elif bczsta == bczend and bczsta == 1:  # pure Neumann   (:531-560)
            tmp = 1.0 / (kh * (1 - np.exp(-2 * kh * Lz)))
            coef1[...] = (bc2 - bc1 * e1) * tmp
            coef2[...] = (-bc1 + bc2 * e1) * tmp
            if g.ista == 1:
                coef1[0, 0] = bc1[0, 0]
                coef2[0, 0] = 0.0
        elif bczsta == bczend and bczsta == 2:  # pure Robin     (:563-592)
            tmp = 1.0 / (2 * kh)
            coef1[...] = bc2 * tmp
            coef2[...] = bc1 * tmp
            if g.ista == 1:
                coef1[0, 0] = bc1[0, 0]
                coef2[0, 0] = 0.0
        elif bczsta == 0 and bczend == 2:  # Dirichlet bottom / Robin top (:595-624)
            tmp = 1.0 / (2 * kh)
            coef1[...] = bc2 * tmp
            coef2[...] = (bc1 * 2 * kh - bc2 * e1) * tmp
            if g.ista == 1:
                coef1[0, 0] = bc2[0, 0]
                coef2[0, 0] = bc1[0, 0]
        else:
            raise ValueError("[ERROR] Unsupported BC combination in call to laplace_z.")
    z = g.z
    ep = np.exp(kh[:, :, None] * (z[None, None, :] - Lz))
    em = np.exp(-kh[:, :, None] * z[None, None, :])
    a = coef1[:, :, None] * ep + coef2[:, :, None] * em
    b = kh[:, :, None] * (coef1[:, :, None] * ep - coef2[:, :, None] * em)
    if g.ista == 1:  # (0,0) mode: linear profile using real(coef)  (:635-639)
        a[0, 0, :] = coef1[0, 0].real * z + coef2[0, 0].real
        b[0, 0, :] = coef1[0, 0].real
    return a, b


def sol_project(g: Grid, a, b, c, bctarget: int, bczsta: int, bczend: int):
    """boundary_mod.fpp:197-402.  a,b,c updated in place; returns d (mixed domain)."""
    d = poisson_inhomogeneous(g, a, b, c)
    a -= IM * g.kx[:, None, None] * d
    b -= IM * g.ky[None, :, None] * d
    c -= IM * g.kz[None, None, :] * d
    tmp = 1.0 / float(g.nz)
    top = g.nz - g.Cz - 1
    if bctarget == 0:
        C1 = d * tmp
    elif bctarget == 1:
        C1 = c * tmp
    else:
        raise ValueError("bctarget")
    C2 = None
    if bczsta == 2 or bczend == 2:
        C2 = derivk(g, C1, 3)
        fftp1d_complex_to_real_z(g, C2)
    fftp1d_complex_to_real_z(g, C1)
    bc = np.empty((g.nxl, g.ny, 2), dtype=np.complex128)
    if bczsta == 0:
        bc[:, :, 0] = -C1[:, :, 0] if bctarget == 0 else C1[:, :, 0]
    elif bczsta == 2:
        bc[:, :, 0] = C2[:, :, 0] - g.khom * C1[:, :, 0]
    else:
        raise ValueError("[ERROR] Unsupported BC kind in call to sol_project.")
    if bczend == 0:
        bc[:, :, 1] = -C1[:, :, top] if bctarget == 0 else C1[:, :, top]
    elif bczend == 2:
        bc[:, :, 1] = -(C2[:, :, top] + g.khom * C1[:, :, top])
    else:
        raise ValueError("[ERROR] Unsupported BC kind in call to sol_project.")
    C2, C3 = laplace_z(g, bc, bczsta + bctarget, bczend + bctarget)
    if bctarget == 0:
        d = C1 + C2
    else:
        fftp1d_complex_to_real_z(g, d)
        d = d * tmp + C2
    fftp1d_real_to_complex_z(g, C2)
    fftp1d_real_to_complex_z(g, C3)
    a -= IM * g.kx[:, None, None] * C2
    b -= IM * g.ky[None, :, None] * C2
    c -= C3
    return d


def noslip_z(g: Grid, o: int, vx, vy, pr, vbound, pos: int):
    """vboundary.f90:154-211 (vx,vy in the mixed domain)."""
    ind = 0 if pos == 0 else g.nz - g.Cz - 1
    tmp = 1.0 / float(o)
    if o != g.ord:
        tmp = float(o + 1) * tmp
    vx[:, :, ind] = IM * g.kx[:, None] * pr[:, :, ind] * tmp
    vy[:, :, ind] = IM * g.ky[None, :] * pr[:, :, ind] * tmp
    if g.ista == 1:
        vx[0, 0, ind] = g.nx * g.ny * vbound[0]
        vy[0, 0, ind] = g.nx * g.ny * vbound[1]


def v_imposebc_and_project(g: Grid, vx, vy, vz, pr, rki: int,
                           v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0)):
    """vboundary.f90:67-151.  vx,vy,vz in place; returns the new pr."""
    goto_domain_w_boundaries(g, vx, vy)
    noslip_z(g, rki, vx, vy, pr, v_zsta, 0)
    noslip_z(g, rki, vx, vy, pr, v_zend, 1)
    goto_3d_fourier(g, vx, vy)
    return sol_project(g, vx, vy, vz, 1, 0, 0)


# ----------------------------------------------------------------------------
# the HD substep                 include/hd/hd_rkstep{1,2}.f90, specter.fpp:1142-1161
# ----------------------------------------------------------------------------
@dataclass
class HDState:
    vx: np.ndarray
    vy: np.ndarray
    vz: np.ndarray
    pr: np.ndarray
    fx: np.ndarray
    fy: np.ndarray
    fz: np.ndarray


def hd_rkstep2(g: Grid, s: HDState, C1, C2, C3, o: int, dt: float, nu: float,
               v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0)):
    """One RK substep, include/hd/hd_rkstep2.f90:3-36."""
    rmp = 1.0 / float(o)
    C4, C5, C6 = gradre(g, s.vx, s.vy, s.vz)
    fc_filter(g, C4); fc_filter(g, C5); fc_filter(g, C6)
    s.vx = laplak(g, s.vx); s.vy = laplak(g, s.vy); s.vz = laplak(g, s.vz)
    s.vx = C1 + dt * (nu * s.vx - C4 + s.fx) * rmp
    s.vy = C2 + dt * (nu * s.vy - C5 + s.fy) * rmp
    s.vz = C3 + dt * (nu * s.vz - C6 + s.fz) * rmp
    s.pr = v_imposebc_and_project(g, s.vx, s.vy, s.vz, s.pr, o, v_zsta, v_zend)


def hd_step(g: Grid, s: HDState, dt: float, nu: float, on_substep=None, **kw):
    """One full time step: rkstep1 copy + ``ord`` substeps (specter.fpp:1142-1161)."""
    C1 = s.vx.copy(); C2 = s.vy.copy(); C3 = s.vz.copy()
    for o in range(g.ord, 0, -1):
        hd_rkstep2(g, s, C1, C2, C3, o, dt, nu, **kw)
        if on_substep is not None:
            on_substep(o, s)


# ----------------------------------------------------------------------------
# synthetic inputs               pseudospec_mod.fpp:101-119, initialv.f90, initialfv.f90
# ----------------------------------------------------------------------------
class Randu:
    """Park-Miller ``randu`` (pseudospec_mod.fpp:101-119); ``am=1./im`` is
    evaluated in single precision before widening."""
    IQ, IR, MASK, IA, IMOD = 127773, 2836, 123459876, 16807, 2147483647

    def __init__(self, seed: int):
        self.idum = int(seed)
        self.am = float(np.float32(1.0) / np.float32(self.IMOD))

    def __call__(self) -> float:
        idum = self.idum ^ self.MASK
        k = idum // self.IQ
        idum = self.IA * (idum - k * self.IQ) - self.IR * k
        if idum < 0:
            idum += self.IMOD
        r = self.am * idum
        r = (r - 0.5) * 2
        self.idum = idum ^ self.MASK
        return r


_CR_ROOTS = [np.float64(np.float32(v)) for v in (
    4.73004074, 7.85320462, 10.99560784, 14.13716549,
    17.27875966, 20.42035225, 23.56194490, 26.70353756)]


def initialv(g: Grid, seed=1000, kdn=2.0, kup=4.0, vparam0=1.0, vparam1=2.0, u0=1.0, rnd=None):
    """Chandrasekhar-Reid no-slip solenoidal noise, initialv.f90:25-203 (1 rank,
    single-stream randu order).  Returns vx,vy,vz in (kz,ky,kx)."""
    assert g.nprocs == 1
    rnd = rnd or Randu(seed)
    nph = g.nz - g.Cz
    z = g.z[:nph]
    Lz = g.Lz
    pi = np.pi
    C1 = np.zeros(g.cshape(), dtype=np.complex128)
    C2 = np.zeros_like(C1)
    C3 = np.zeros_like(C1)
    ny = g.ny
    for kk in range(int(vparam0), int(vparam1) + 1):
        if kk <= 8:
            rm1 = float(_CR_ROOTS[kk - 1])
        elif kk % 2 == 1:
            rm1 = (2 * kk - 0.5) * pi
        else:
            rm1 = (2 * kk + 0.5) * pi
        s = rm1 * (z / Lz - 0.5)
        if kk % 2 == 1:
            dphi = rm1 / Lz * (np.sinh(s) / np.cosh(0.5 * rm1) + np.sin(s) / np.cos(0.5 * rm1))
            phi = np.cosh(s) / np.cosh(0.5 * rm1) - np.cos(s) / np.cos(0.5 * rm1)
        else:
            dphi = rm1 / Lz * (np.cosh(s) / np.sinh(0.5 * rm1) - np.cos(s) / np.sin(0.5 * rm1))
            phi = np.sinh(s) / np.sinh(0.5 * rm1) - np.sin(s) / np.sin(0.5 * rm1)
        psi = np.sin(kk * pi * z / Lz)

        def add(i, j, mirror):
            k2 = g.kk2[i, j, 0]
            if not (k2 <= kup ** 2 and k2 >= kdn ** 2):
                return
            rmp = 2 * pi * rnd()
            rmq = rnd() / np.sqrt(k2 + rm1 ** 2) ** 3
            ph = np.cos(rmp) + IM * np.sin(rmp)
            C1[i, j, :nph] += rmq * ph * dphi
            C2[i, j, :nph] += rmq * ph * phi
            if mirror:
                jm = (ny - j) % ny  # Fortran index ny-j+2
                C1[i, jm, :nph] = np.conj(C1[i, j, :nph])
                C2[i, jm, :nph] = np.conj(C2[i, j, :nph])
            if kk % 2 == 0:
                rmp = pi / 2 + rmp
            rmq = rnd() / np.sqrt(k2 + kk ** 2 * pi ** 2 / Lz ** 2) ** 2
            ph = np.cos(rmp) + IM * np.sin(rmp)
            C3[i, j, :nph] += rmq * ph * psi
            if mirror:
                C3[i, jm, :nph] = np.conj(C3[i, j, :nph])

        for j in range(0, ny // 2 + 1):
            add(0, j, True)
        for i in range(1, g.nxl):
            for j in range(ny):
                add(i, j, False)
    vz = -(derivk(g, derivk(g, C2, 1), 1) + derivk(g, derivk(g, C2, 2), 2))
    vx = derivk(g, C1, 1) + derivk(g, C3, 2)
    vy = derivk(g, C1, 2) - derivk(g, C3, 1)
    fftp1d_real_to_complex_z(g, vx)
    fftp1d_real_to_complex_z(g, vy)
    fftp1d_real_to_complex_z(g, vz)
    normvec(g, vx, vy, vz, u0, 1)
    return vx, vy, vz


def initialfv(g: Grid, f0=1.0):
    """initialfv.f90:25-31: uniform body force in x, mode (1,1,1) scaled by N."""
    fx = np.zeros(g.cshape(), dtype=np.complex128)
    fy = np.zeros_like(fx)
    fz = np.zeros_like(fx)
    if g.ista == 1:
        fx[0, 0, 0] = f0 * g.N
    return fx, fy, fz


def make_hd_state(g: Grid, **ic) -> HDState:
    vx, vy, vz = initialv(g, **ic)
    fx, fy, fz = initialfv(g)
    pr = np.zeros(g.cshape(), dtype=np.complex128)
    return HDState(vx, vy, vz, pr, fx, fy, fz)


# ----------------------------------------------------------------------------
# field files and the output / restart blocks of the driver
# ----------------------------------------------------------------------------
def io_path(dir, fname, nmb):
    """binary_io.f90:202-204."""
    import os
    return os.path.join(str(dir), f"{fname}.{nmb}.out")


def io_write(g: Grid, dir, fname, nmb, var: np.ndarray):
    """io_write, mpiio/binary_io.f90:165-223 with the view of io_init (:15-86): raw native reals, Fortran order,
    global extent (nx, ny, nz-Cz); this rank's planes ksta..min(kend, nz-Cz) at the offset of its first plane."""
    nk = max(0, min(g.kend, g.nz - g.Cz) - g.ksta + 1)
    data = np.ascontiguousarray(var[:nk], dtype=np.float64)
    path = io_path(dir, fname, nmb)
    mode = "r+b" if __import__("os").path.exists(path) else "w+b"
    with open(path, mode) as f:
        f.seek((g.ksta - 1) * g.nx * g.ny * 8)
        f.write(data.tobytes())


def io_read(g: Grid, dir, fname, nmb) -> np.ndarray:
    """io_read, mpiio/binary_io.f90:89-160; planes that are not in the file stay zero."""
    nk = max(0, min(g.kend, g.nz - g.Cz) - g.ksta + 1)
    out = np.zeros(g.rshape())
    with open(io_path(dir, fname, nmb), "rb") as f:
        f.seek((g.ksta - 1) * g.nx * g.ny * 8)
        out[:nk] = np.frombuffer(f.read(nk * g.nx * g.ny * 8), dtype=np.float64).reshape(nk, g.ny, g.nx)
    return out


def hd_output(g: Grid, s: "HDState", odir, ext, dt, outs=0):
    """The BIN block of specter.fpp:1005-1053 (HD fields)."""
    rmp = 1.0 / (float(g.nx) * float(g.ny) * float(g.nz))
    C1, C2, C3 = s.vx * rmp, s.vy * rmp, s.vz * rmp
    if outs >= 1:
        io_write(g, odir, "wx", ext, fftp3d_complex_to_real(g, curlk(g, C2, C3, 1)))
        io_write(g, odir, "wy", ext, fftp3d_complex_to_real(g, curlk(g, C1, C3, 2)))
        io_write(g, odir, "wz", ext, fftp3d_complex_to_real(g, curlk(g, C1, C2, 3)))
    io_write(g, odir, "vx", ext, fftp3d_complex_to_real(g, C1))
    io_write(g, odir, "vy", ext, fftp3d_complex_to_real(g, C2))
    io_write(g, odir, "vz", ext, fftp3d_complex_to_real(g, C3))
    rmp = 1.0 / (float(g.nx) * float(g.ny) * dt)
    io_write(g, odir, "pr", ext, fftp2d_complex_to_real_xy(g, s.pr * rmp))


def hd_restart(g: Grid, idir, ext, dt):
    """The stat != 0 branch of specter.fpp:886-912: returns (vx, vy, vz, pr) with pr back in p' units."""
    v = [fftp3d_real_to_complex(g, io_read(g, idir, n, ext)) for n in ("vx", "vy", "vz")]
    pr = fftp2d_real_to_complex_xy(g, io_read(g, idir, "pr", ext))
    pr[:, :, : g.nz - g.Cz] *= dt
    return v[0], v[1], v[2], pr


def analytic_field(g: Grid, kind="sin"):
    """The src/tests fields: sin4x cos8y sin6z (fc_dirichlet.f90:14, energy.f90:13)
    or sin4x cos8y exp(.4 z/Lz) (poisson.f90:21) on the local real slab."""
    z = g.z[g.ksta - 1 : g.kend][:, None, None]
    y = g.y[None, :, None]
    x = g.x[None, None, :]
    if kind == "sin":
        return np.sin(4 * x) * np.cos(8 * y) * np.sin(6 * z)
    return np.sin(4 * x) * np.cos(8 * y) * np.exp(0.4 * z / g.Lz)


# ============================================================================
# Boussinesq and vector-potential MHD            (configs 3 and 4 of BASELINE.json)
# ============================================================================
def advect(g: Grid, a, b, c, d):
    """A.grad(d), pseudospec_phd.f90:23-113."""
    ph = _phys(g)
    r3 = np.zeros(g.rshape())
    for comp, dir_ in ((a, 1), (b, 2), (c, 3)):
        r1 = fftp3d_complex_to_real(g, comp)
        r2 = fftp3d_complex_to_real(g, derivk(g, d, dir_))
        r3[ph] += r1[ph] * r2[ph]
    r3[ph] *= 1.0 / g.N ** 2
    return fftp3d_real_to_complex(g, r3)


def vector(g: Grid, a, b, c, d, e, f):
    """A x B in real space, pseudospec_mhd.f90:22-105."""
    ph = _phys(g)
    r1 = fftp3d_complex_to_real(g, a); r2 = fftp3d_complex_to_real(g, b); r3 = fftp3d_complex_to_real(g, c)
    r4 = fftp3d_complex_to_real(g, d); r5 = fftp3d_complex_to_real(g, e); r6 = fftp3d_complex_to_real(g, f)
    tmp = 1.0 / g.N ** 2
    o1 = np.zeros(g.rshape()); o2 = np.zeros(g.rshape()); o3 = np.zeros(g.rshape())
    o1[ph] = (r2[ph] * r6[ph] - r5[ph] * r3[ph]) * tmp
    o2[ph] = (r3[ph] * r4[ph] - r6[ph] * r1[ph]) * tmp
    o3[ph] = (r1[ph] * r5[ph] - r4[ph] * r2[ph]) * tmp
    return (fftp3d_real_to_complex(g, o1), fftp3d_real_to_complex(g, o2), fftp3d_real_to_complex(g, o3))


def variance(g: Grid, a, kin: int) -> float:
    """pseudospec_phd.f90:116-196."""
    tmp = 1.0 / g.N ** 2 / float(g.nz - g.Cz)
    at = a.copy() if kin == 1 else g.kk2 * a
    return _mean_phys(g, _abs2_iz(g, at), tmp)


def normsca(g: Grid, a, b: float, kin: int):
    """pseudospec_phd.f90:324-368."""
    a *= np.sqrt(b / variance(g, a, kin))


def s_imposebc(g: Grid, th):
    """sboundary.f90:67-119 with `constant' walls (s_constant_z :122-165 sets the rows to 0)."""
    goto_domain_w_boundaries(g, th)
    th[:, :, 0] = 0.0
    th[:, :, g.nz - g.Cz - 1] = 0.0
    goto_3d_fourier(g, th)


@dataclass
class BoussState(HDState):
    th: np.ndarray = None
    fs: np.ndarray = None


def bouss_rkstep2(g: Grid, s: BoussState, C1, C2, C3, C7, o: int, dt: float, nu: float, kappa: float,
                  xmom: float = 1.0, xtemp: float = 1.0, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0)):
    """include/bouss/bouss_rkstep2.f90:3-59, including the theta `hack' (:57-59)."""
    rmp = 1.0 / float(o)
    C4, C5, C6 = gradre(g, s.vx, s.vy, s.vz)
    C8 = advect(g, s.vx, s.vy, s.vz, s.th)
    C6 = C6 - xmom * s.th
    C8 = C8 - xtemp * s.vz
    for q in (C4, C5, C6, C8):
        fc_filter(g, q)
    s.vx = laplak(g, s.vx); s.vy = laplak(g, s.vy); s.vz = laplak(g, s.vz); s.th = laplak(g, s.th)
    s.vx = C1 + dt * (nu * s.vx - C4 + s.fx) * rmp
    s.vy = C2 + dt * (nu * s.vy - C5 + s.fy) * rmp
    s.vz = C3 + dt * (nu * s.vz - C6 + s.fz) * rmp
    s.th = C7 + dt * (kappa * s.th - C8 + s.fs) * rmp
    s.pr = v_imposebc_and_project(g, s.vx, s.vy, s.vz, s.pr, o, v_zsta, v_zend)
    s_imposebc(g, s.th)
    fc_filter(g, s.th)
    R1 = fftp3d_complex_to_real(g, s.th)
    R1 = R1 / g.nx / g.ny / g.nz
    s.th = fftp3d_real_to_complex(g, R1)


def bouss_step(g: Grid, s: BoussState, dt, nu, kappa, on_substep=None, **kw):
    C1 = s.vx.copy(); C2 = s.vy.copy(); C3 = s.vz.copy(); C7 = s.th.copy()
    for o in range(g.ord, 0, -1):
        bouss_rkstep2(g, s, C1, C2, C3, C7, o, dt, nu, kappa, **kw)
        if on_substep is not None:
            on_substep(o, s)


def initials(g: Grid, rnd: "Randu", c0=0.5, skdn=3.0, skup=8.0, cparam0=1.0, cparam1=2.0):
    """examples/initials.f90_random (1 rank).  Note rmp uses kk2 at z-INDEX kk (as written there)."""
    nph = g.nz - g.Cz
    z = g.z[:nph]
    pi, Lz, ny = np.pi, g.Lz, g.ny
    th = np.zeros(g.cshape(), dtype=np.complex128)
    for kk in range(int(cparam0), int(cparam1) + 1):
        def put(i, j, mirror):
            k2 = g.kk2[i, j, 0]
            if not (k2 <= skup ** 2 and k2 >= skdn ** 2):
                return
            rmq = 2 * pi * rnd()
            rmp = 1.0 / np.sqrt(g.kk2[i, j, kk - 1])
            ph = np.cos(rmq) + IM * np.sin(rmq)
            th[i, j, :nph] = rmp * ph * np.sin(2 * kk * pi / Lz * z)
            th[i, j, :nph] = th[i, j, :nph] + rmp * ph * np.sin((2 * kk - 1) * pi / Lz * z)
            if mirror:
                th[i, (ny - j) % ny, :nph] = np.conj(th[i, j, :nph])
        for j in range(1, ny // 2 + 1):
            put(0, j, True)
        for i in range(1, g.nxl):
            for j in range(ny):
                put(i, j, False)
    fftp1d_real_to_complex_z(g, th)
    normsca(g, th, c0, 1)
    return th


def make_bouss_state(g: Grid, seed=1000, **ic) -> BoussState:
    """Velocity from initialv, then theta from the same randu stream (specter.fpp:846-874 order)."""
    rnd = Randu(seed)
    vx, vy, vz = initialv(g, rnd=rnd, **ic)
    th = initials(g, rnd)
    fx, fy, fz = initialfv(g)
    z = np.zeros(g.cshape(), dtype=np.complex128)
    return BoussState(vx, vy, vz, z.copy(), fx, fy, fz, th, z.copy())


def neumann_reconstruct(g: Grid, f, boun: int, order: int):
    """fcgram_mod.f90:368-511, z branches (:456-499).  boun 5 = z=0, 6 = z=Lz; the prescribed normal
    derivative sits in the wall row on entry."""
    neu = g.neu if order == 1 else g.neu2
    d = g.oz
    if boun == 5:
        acc = neu[d - 1] * f[:, :, 0]
        for k in range(1, d):
            acc = acc + neu[k - 1] * f[:, :, d - k]          # f(dz+1-k)
        f[:, :, 0] = acc
    elif boun == 6:
        top = g.nz - g.Cz - 1
        acc = neu[d - 1] * f[:, :, top]
        for k in range(1, d):
            acc = acc + neu[k - 1] * f[:, :, top - d + k]    # f(nz-Cz-dz+k)
        f[:, :, top] = acc
    else:
        raise ValueError("[ERROR] Neumann reconstruction not performed.")


def a_imposebc_and_project(g: Grid, ax, ay, az, bczsta: int = 0, bczend: int = 0):
    """bboundary.f90:100-189 for conducting walls (bc kind 0) at both ends.  Returns ph."""
    if bczsta != 0 or bczend != 0:
        raise ValueError("[ERROR] Unsupported boundary conditions in Z direction (oracle: conducting only).")
    if g.ista == 1:
        az[0, 0, 0] = 0.0
    top = g.nz - g.Cz - 1
    goto_domain_w_boundaries(g, ax, ay)
    for ind in (0, top):                                   # int_conducting_z :192-236
        ax[:, :, ind] = 0.0
        ay[:, :, ind] = 0.0
    goto_3d_fourier(g, ax, ay)
    ph = sol_project(g, ax, ay, az, 0, 0, 0)
    goto_domain_w_boundaries(g, ax, ay, az)
    for pos, ind in ((0, 0), (1, top)):                     # conducting_z :239-290
        ax[:, :, ind] = 0.0
        ay[:, :, ind] = 0.0
        az[:, :, ind] = 0.0
        neumann_reconstruct(g, ax, 5 + pos, 2)
        neumann_reconstruct(g, ay, 5 + pos, 2)
        neumann_reconstruct(g, az, 5 + pos, 1)
    goto_3d_fourier(g, ax, ay, az)
    return ph


@dataclass
class MhdState(HDState):
    ax: np.ndarray = None
    ay: np.ndarray = None
    az: np.ndarray = None
    ph: np.ndarray = None
    mx: np.ndarray = None
    my: np.ndarray = None
    mz: np.ndarray = None


def mhd_rkstep2(g: Grid, s: MhdState, C1, C2, C3, C9, C10, C11, o: int, dt: float, nu: float, mu: float,
                b0=(0.0, 0.0, 0.0)):
    """include/mhd/mhd_rkstep2.f90:3-84 (ax,ay,az hold J when the RK update reads them)."""
    rmp = 1.0 / float(o)
    C12 = curlk(g, s.ay, s.az, 1); C13 = curlk(g, s.ax, s.az, 2); C14 = curlk(g, s.ax, s.ay, 3)
    if g.ista == 1:
        C12[0, 0, 0] = b0[0] * g.N; C13[0, 0, 0] = b0[1] * g.N; C14[0, 0, 0] = b0[2] * g.N
    s.ax = curlk(g, C13, C14, 1); s.ay = curlk(g, C12, C14, 2); s.az = curlk(g, C12, C13, 3)
    C4, C5, C6 = prodre(g, s.vx, s.vy, s.vz)
    C15, C16, C17 = vector(g, s.ax, s.ay, s.az, C12, C13, C14)
    C4 = C4 - C15; C5 = C5 - C16; C6 = C6 - C17
    for q in (C4, C5, C6):
        fc_filter(g, q)
    C15, C16, C17 = vector(g, s.vx, s.vy, s.vz, C12, C13, C14)
    for q in (C15, C16, C17):
        fc_filter(g, q)
    s.vx = laplak(g, s.vx); s.vy = laplak(g, s.vy); s.vz = laplak(g, s.vz)
    s.vx = C1 + dt * (nu * s.vx - C4 + s.fx) * rmp
    s.vy = C2 + dt * (nu * s.vy - C5 + s.fy) * rmp
    s.vz = C3 + dt * (nu * s.vz - C6 + s.fz) * rmp
    s.ax = C9 + dt * (-mu * s.ax + C15 + s.mx) * rmp
    s.ay = C10 + dt * (-mu * s.ay + C16 + s.my) * rmp
    s.az = C11 + dt * (-mu * s.az + C17 + s.mz) * rmp
    s.pr = v_imposebc_and_project(g, s.vx, s.vy, s.vz, s.pr, o)
    s.ph = a_imposebc_and_project(g, s.ax, s.ay, s.az)


def mhd_step(g: Grid, s: MhdState, dt, nu, mu, on_substep=None, **kw):
    C = [q.copy() for q in (s.vx, s.vy, s.vz, s.ax, s.ay, s.az)]
    for o in range(g.ord, 0, -1):
        mhd_rkstep2(g, s, *C, o, dt, nu, mu, **kw)
        if on_substep is not None:
            on_substep(o, s)


def initialb(g: Grid, rnd: "Randu", a0=1.0, mkdn=2.0, mkup=4.0, aparam0=1.0, aparam1=2.0):
    """initialb.f90 (1 rank): toroidal vector potential from a random stream function."""
    nph = g.nz - g.Cz
    z = g.z[:nph]
    pi, Lz, ny = np.pi, g.Lz, g.ny
    C3 = np.zeros(g.cshape(), dtype=np.complex128)
    for kk in range(int(aparam0), int(aparam1) + 1):
        def put(i, j, mirror):
            k2 = g.kk2[i, j, 0]
            if not (k2 <= mkup ** 2 and k2 >= mkdn ** 2):
                return
            rmp = 2 * pi * rnd()
            rmq = rnd() / np.sqrt(k2 + kk ** 2 * pi ** 2 / Lz ** 2) ** 2
            C3[i, j, :nph] = rmq * (np.cos(rmp) + IM * np.sin(rmp)) * np.sin(kk * pi * z / Lz)
            if mirror:
                C3[i, (ny - j) % ny, :nph] = np.conj(C3[i, j, :nph])
        for j in range(1, ny // 2 + 1):
            put(0, j, True)
        for i in range(1, g.nxl):
            for j in range(ny):
                put(i, j, False)
    ax = derivk(g, C3, 2)
    ay = derivk(g, -C3, 1)
    az = np.zeros_like(ax)
    fftp1d_real_to_complex_z(g, ax); fftp1d_real_to_complex_z(g, ay); fftp1d_real_to_complex_z(g, az)
    normvec(g, ax, ay, az, a0, 0)
    return ax, ay, az


def make_mhd_state(g: Grid, seed=1000, **ic) -> MhdState:
    if g.neu is None:
        g.load_neumann()
    rnd = Randu(seed)
    vx, vy, vz = initialv(g, rnd=rnd, **ic)
    ax, ay, az = initialb(g, rnd)
    z = np.zeros(g.cshape(), dtype=np.complex128)
    return MhdState(vx, vy, vz, z.copy(), z.copy(), z.copy(), z.copy(), ax, ay, az, z.copy(), z.copy(), z.copy(), z.copy())


# ----------------------------------------------------------------------------
# vacuum (Robin) walls, ROTBOUSS / MHDBOUSS substeps, the remaining diagnostics
# ----------------------------------------------------------------------------
def robin_reconstruct(g: Grid, f, boun: int, a: np.ndarray):
    """fcgram_mod.f90:514-644, z branches (:612-635): wall value from f' + a f = g_wall stored in the wall row;
    a[i,j] is the (real) coefficient per (kx,ky) (the callers pass khom)."""
    if g.neu is None:
        raise ValueError("[ERROR] Neumann table not loaded in robin_reconstruct. Aborting...")
    neu = g.neu
    d = g.oz
    if boun == 5:
        acc = neu[d - 1] * f[:, :, 0]
        for k in range(1, d):
            acc = acc + neu[k - 1] * f[:, :, d - k]
        f[:, :, 0] = acc / (a * neu[d - 1] + 1)
    elif boun == 6:
        top = g.nz - g.Cz - 1
        acc = neu[d - 1] * f[:, :, top]
        for k in range(1, d):
            acc = acc + neu[k - 1] * f[:, :, top - d + k]
        f[:, :, top] = acc / (a * neu[d - 1] + 1)
    else:
        raise ValueError("[ERROR] Robin reconstruction not performed. Wrong boundary specified. Aborting...")


B_CONDUCTING, B_VACUUM = 0, 1   # b_parsebc, bboundary.f90:15-40


def a_imposebc_and_project_bc(g: Grid, ax, ay, az, bczsta: int, bczend: int):
    """bboundary.f90:100-189 for every wall combination it accepts (0 conducting, 1 vacuum).  Returns ph."""
    if bczsta not in (0, 1) or bczend not in (0, 1):
        raise ValueError("[ERROR] Unsupported boundary conditions in Z direction. Aborting...")
    if g.neu is None:
        g.load_neumann()
    if g.ista == 1:
        az[0, 0, 0] = 0.0
    top = g.nz - g.Cz - 1
    walls = ((0, 0, bczsta), (1, top, bczend))
    if bczsta == 0 or bczend == 0:
        goto_domain_w_boundaries(g, ax, ay)
        for pos, ind, kind in walls:                       # int_conducting_z :192-236
            if kind == 0:
                ax[:, :, ind] = 0.0
                ay[:, :, ind] = 0.0
        goto_3d_fourier(g, ax, ay)
    ph = sol_project(g, ax, ay, az, 0, 2 * bczsta, 2 * bczend)
    goto_domain_w_boundaries(g, ax, ay, az)
    for pos, ind, kind in walls:
        ax[:, :, ind] = 0.0
        ay[:, :, ind] = 0.0
        az[:, :, ind] = 0.0
        if kind == 0:                                      # conducting_z :239-290
            neumann_reconstruct(g, ax, 5 + pos, 2)
            neumann_reconstruct(g, ay, 5 + pos, 2)
            neumann_reconstruct(g, az, 5 + pos, 1)
        else:                                              # insulating_z :294-344
            robin_reconstruct(g, ax, 5 + pos, g.khom)
            robin_reconstruct(g, ay, 5 + pos, g.khom)
            robin_reconstruct(g, az, 5 + pos, g.khom)
    goto_3d_fourier(g, ax, ay, az)
    return ph


def helicity(g: Grid, a, b, c) -> float:
    """pseudospec_hd.f90:638-775: <A . curl A> over the physical rows."""
    tmp = 1.0 / g.N ** 2 / float(g.nz - g.Cz)

    def prod(p, q):
        c1 = p.copy(); c2 = q.copy()
        fftp1d_complex_to_real_z(g, c1); fftp1d_complex_to_real_z(g, c2)
        return (c1 * np.conj(c2)).real

    r1 = prod(a, curlk(g, b, c, 1))
    r1 = r1 + prod(b, curlk(g, a, c, 2))
    r1 = r1 + prod(c, curlk(g, a, b, 3))
    return _mean_phys(g, r1, tmp)


def product(g: Grid, a, b) -> float:
    """pseudospec_phd.f90:199-272."""
    tmp = 1.0 / g.N ** 2 / float(g.nz - g.Cz)
    at = a.copy(); bt = b.copy()
    fftp1d_complex_to_real_z(g, at); fftp1d_complex_to_real_z(g, bt)
    return _mean_phys(g, (at * np.conj(bt)).real, tmp)


def pscheck(g: Grid, a, b):
    """pseudospec_phd.f90:275-321 -> the scalar.txt columns (eng, ens, pot)."""
    return variance(g, a, 1), variance(g, a, 0), product(g, a, b)


def maxabs(g: Grid, a, b, c, kin: int) -> float:
    """pseudospec_hd.f90:1008-1079 (one rank): max |curl|, |laplacian| or |field| over the physical rows."""
    if kin == 0:
        c1, c2, c3 = curlk(g, b, c, 1), curlk(g, a, c, 2), curlk(g, a, b, 3)
    elif kin == 1:
        c1, c2, c3 = laplak(g, a), laplak(g, b), laplak(g, c)
    else:
        c1, c2, c3 = a, b, c
    r1 = fftp3d_complex_to_real(g, c1); r2 = fftp3d_complex_to_real(g, c2); r3 = fftp3d_complex_to_real(g, c3)
    ph = _phys(g)
    dloc = float(np.max(r1[ph] ** 2 + r2[ph] ** 2 + r3[ph] ** 2))
    return np.sqrt(dloc) / g.N


def mhdcheck(g: Grid, a, b, c, ma, mb, mc):
    """pseudospec_mhd.f90:109-212 with hel=1, crs=1 -> (eng, ens, cur, engk, engm, helk, helm, crh, asq):
    the columns of balance.txt, energy.txt, helicity.txt and cross.txt."""
    engk = energy(g, a, b, c, 1)
    ens = energy(g, a, b, c, 0)
    engm = energy(g, ma, mb, mc, 0)
    cur = energy(g, ma, mb, mc, 2)
    helk = helicity(g, a, b, c)
    helm = helicity(g, ma, mb, mc)
    asq = energy(g, ma, mb, mc, 1)
    crh = cross(g, a, b, c, derivk(g, ma, 1), derivk(g, mb, 2), derivk(g, mc, 3), 1)
    return engk + engm, ens, cur, engk, engm, helk, helm, crh, asq


def robcheck(g: Grid, a, b, c):
    """bboundary.f90:434-602: mean squared residual of the vacuum condition da/dn + khom a at both walls,
    tangential (a,b) and normal (c) parts."""
    tmp = 1.0 / g.N ** 2
    top = g.nz - g.Cz - 1
    w = _weights(g)[:, None]

    def resid(q):
        C1 = q.copy()
        C2 = derivk(g, C1, 3)
        fftp1d_complex_to_real_z(g, C1); fftp1d_complex_to_real_z(g, C2)
        r0 = -C2[:, :, 0] + g.khom * C1[:, :, 0]
        r1 = C2[:, :, top] + g.khom * C1[:, :, top]
        return r0.real ** 2 + r0.imag ** 2, r1.real ** 2 + r1.imag ** 2

    a0, a1 = resid(a); b0, b1 = resid(b); c0, c1 = resid(c)
    return (float(np.sum(w * (a0 + b0) * tmp)), float(np.sum(w * (a1 + b1) * tmp)),
            float(np.sum(w * c0 * tmp)), float(np.sum(w * c1 * tmp)))


def bdiagnostic(g: Grid, a, b, c, bczsta: int = 0, bczend: int = 0):
    """bboundary.f90:348-430 -> {'conducting': 6 columns, 'vacuum': 6 columns} of the two diagnostic files."""
    c1 = curlk(g, b, c, 1); c2 = curlk(g, a, c, 2); c3 = curlk(g, a, b, 3)
    tm1 = divergence(g, a, b, c)
    tm2 = divergence(g, c1, c2, c3)
    out = {}
    if bczsta == 0 or bczend == 0:
        tmr, tms = bouncheck_z(g, c3)
        c4 = curlk(g, c2, c3, 1)
        c2b = curlk(g, c1, c3, 2)
        tmp, tmq = bouncheck_z(g, c4, c2b)
        out["conducting"] = (tm1, tm2, tmp, tmq, tmr, tms)
    if bczsta == 1 or bczend == 1:
        out["vacuum"] = (tm1, tm2) + robcheck(g, a, b, c)
    return out


def sdiagnostic(g: Grid, a):
    """sboundary.f90:168-210 -> scalar_constant_diagnostic.txt columns."""
    return bouncheck_z(g, a)


def rotbouss_rkstep2(g: Grid, s: BoussState, C1, C2, C3, C7, o: int, dt: float, nu: float, kappa: float,
                     xmom: float = 1.0, xtemp: float = 1.0, omega=(0.0, 0.0, 0.0),
                     v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0)):
    """include/rotbouss/rotbouss_rkstep2.f90:3-56 (no theta filter / round trip at the end, unlike BOUSS)."""
    rmp = 1.0 / float(o)
    ox, oy, oz_ = omega
    C4, C5, C6 = gradre(g, s.vx, s.vy, s.vz)
    C8 = advect(g, s.vx, s.vy, s.vz, s.th)
    C4 = C4 + 2 * (oy * s.vz - oz_ * s.vy)
    C5 = C5 + 2 * (oz_ * s.vx - ox * s.vz)
    C6 = C6 + 2 * (ox * s.vy - oy * s.vx) - xmom * s.th
    C8 = C8 - xtemp * s.vz
    for q in (C4, C5, C6, C8):
        fc_filter(g, q)
    s.vx = laplak(g, s.vx); s.vy = laplak(g, s.vy); s.vz = laplak(g, s.vz); s.th = laplak(g, s.th)
    s.vx = C1 + dt * (nu * s.vx - C4 + s.fx) * rmp
    s.vy = C2 + dt * (nu * s.vy - C5 + s.fy) * rmp
    s.vz = C3 + dt * (nu * s.vz - C6 + s.fz) * rmp
    s.th = C7 + dt * (kappa * s.th - C8 + s.fs) * rmp
    s.pr = v_imposebc_and_project(g, s.vx, s.vy, s.vz, s.pr, o, v_zsta, v_zend)
    s_imposebc(g, s.th)


def rotbouss_step(g: Grid, s: BoussState, dt, nu, kappa, on_substep=None, **kw):
    C1 = s.vx.copy(); C2 = s.vy.copy(); C3 = s.vz.copy(); C7 = s.th.copy()
    for o in range(g.ord, 0, -1):
        rotbouss_rkstep2(g, s, C1, C2, C3, C7, o, dt, nu, kappa, **kw)
        if on_substep is not None:
            on_substep(o, s)


@dataclass
class MhdBoussState(MhdState):
    th: np.ndarray = None
    fs: np.ndarray = None


def mhdbouss_rkstep2(g: Grid, s: MhdBoussState, C1, C2, C3, C7, C9, C10, C11, o: int, dt: float, nu: float,
                     mu: float, kappa: float, xmom: float = 1.0, xtemp: float = 1.0, b0=(0.0, 0.0, 0.0),
                     bczsta: int = 0, bczend: int = 0):
    """include/mhdbouss/mhdbouss_rkstep2.f90:3-106 (theta round trip without the filter; static walls)."""
    rmp = 1.0 / float(o)
    C12 = curlk(g, s.ay, s.az, 1); C13 = curlk(g, s.ax, s.az, 2); C14 = curlk(g, s.ax, s.ay, 3)
    if g.ista == 1:
        C12[0, 0, 0] = b0[0] * g.N; C13[0, 0, 0] = b0[1] * g.N; C14[0, 0, 0] = b0[2] * g.N
    s.ax = curlk(g, C13, C14, 1); s.ay = curlk(g, C12, C14, 2); s.az = curlk(g, C12, C13, 3)
    C4, C5, C6 = prodre(g, s.vx, s.vy, s.vz)
    C8 = advect(g, s.vx, s.vy, s.vz, s.th)
    C15, C16, C17 = vector(g, s.ax, s.ay, s.az, C12, C13, C14)
    C4 = C4 - C15; C5 = C5 - C16; C6 = C6 - C17
    C6 = C6 - xmom * s.th
    C8 = C8 - xtemp * s.vz
    for q in (C4, C5, C6, C8):
        fc_filter(g, q)
    C15, C16, C17 = vector(g, s.vx, s.vy, s.vz, C12, C13, C14)
    for q in (C15, C16, C17):
        fc_filter(g, q)
    s.vx = laplak(g, s.vx); s.vy = laplak(g, s.vy); s.vz = laplak(g, s.vz); s.th = laplak(g, s.th)
    s.vx = C1 + dt * (nu * s.vx - C4 + s.fx) * rmp
    s.vy = C2 + dt * (nu * s.vy - C5 + s.fy) * rmp
    s.vz = C3 + dt * (nu * s.vz - C6 + s.fz) * rmp
    s.ax = C9 + dt * (-mu * s.ax + C15 + s.mx) * rmp
    s.ay = C10 + dt * (-mu * s.ay + C16 + s.my) * rmp
    s.az = C11 + dt * (-mu * s.az + C17 + s.mz) * rmp
    s.th = C7 + dt * (kappa * s.th - C8 + s.fs) * rmp
    s.pr = v_imposebc_and_project(g, s.vx, s.vy, s.vz, s.pr, o)
    s.ph = a_imposebc_and_project_bc(g, s.ax, s.ay, s.az, bczsta, bczend)
    s_imposebc(g, s.th)
    R1 = fftp3d_complex_to_real(g, s.th)
    R1 = R1 / g.nx / g.ny / g.nz
    s.th = fftp3d_real_to_complex(g, R1)


def mhdbouss_step(g: Grid, s: MhdBoussState, dt, nu, mu, kappa, on_substep=None, **kw):
    C = [q.copy() for q in (s.vx, s.vy, s.vz, s.th, s.ax, s.ay, s.az)]
    for o in range(g.ord, 0, -1):
        mhdbouss_rkstep2(g, s, *C, o, dt, nu, mu, kappa, **kw)
        if on_substep is not None:
            on_substep(o, s)


def make_mhdbouss_state(g: Grid, seed=1000, **ic) -> MhdBoussState:
    """initialv + initials + initialb drawn from one randu stream in the driver's order (specter.fpp:861-874:
    velocity, then scalar, then vector potential)."""
    rnd = Randu(seed)
    vx, vy, vz = initialv(g, rnd=rnd, **ic)
    th = initials(g, rnd)
    ax, ay, az = initialb(g, rnd)
    fx, fy, fz = initialfv(g)
    z = np.zeros(g.cshape(), dtype=np.complex128)
    return MhdBoussState(vx, vy, vz, z.copy(), fx, fy, fz, ax, ay, az, z.copy(), z.copy(), z.copy(), z.copy(),
                         th, z.copy())


def solver_output(g: Grid, s, odir, ext, dt, outs=0):
    """The whole BIN block of specter.fpp:1005-1128: HD fields, then th (SCALAR_), then a, b, j, ph (MAGFIELD_)."""
    hd_output(g, s, odir, ext, dt, outs)
    rmp = 1.0 / (float(g.nx) * float(g.ny) * float(g.nz))
    if getattr(s, "th", None) is not None:
        io_write(g, odir, "th", ext, fftp3d_complex_to_real(g, s.th * rmp))
    if getattr(s, "ax", None) is not None:
        C1, C2, C3 = s.ax * rmp, s.ay * rmp, s.az * rmp
        if outs >= 1:
            io_write(g, odir, "bx", ext, fftp3d_complex_to_real(g, curlk(g, C2, C3, 1)))
            io_write(g, odir, "by", ext, fftp3d_complex_to_real(g, curlk(g, C1, C3, 2)))
            io_write(g, odir, "bz", ext, fftp3d_complex_to_real(g, curlk(g, C1, C2, 3)))
        if outs == 2:
            for n, c in (("jx", C1), ("jy", C2), ("jz", C3)):
                io_write(g, odir, n, ext, fftp3d_complex_to_real(g, laplak(g, c)))
        for n, c in (("ax", C1), ("ay", C2), ("az", C3)):
            io_write(g, odir, n, ext, fftp3d_complex_to_real(g, c))
        io_write(g, odir, "ph", ext, fftp2d_complex_to_real_xy(g, s.ph * (1.0 / (float(g.nx) * float(g.ny) * dt))))


def solver_restart(g: Grid, idir, ext, dt, scalar=False, magnetic=False):
    """The stat != 0 branch of specter.fpp:886-957 -> dict of spectral fields (pr, ph back in primed units)."""
    vx, vy, vz, pr = hd_restart(g, idir, ext, dt)
    out = {"vx": vx, "vy": vy, "vz": vz, "pr": pr}
    if scalar:
        out["th"] = fftp3d_real_to_complex(g, io_read(g, idir, "th", ext))
    if magnetic:
        for n in ("ax", "ay", "az"):
            out[n] = fftp3d_real_to_complex(g, io_read(g, idir, n, ext))
        ph = fftp2d_real_to_complex_xy(g, io_read(g, idir, "ph", ext))
        ph[:, :, : g.nz - g.Cz] *= dt
        out["ph"] = ph
    return out


# ----------------------------------------------------------------------------
# BOOTS regridder                                          tools/boots.fpp
# ----------------------------------------------------------------------------
def boots_points(nzt: int, nzp: int):
    """Continuation points of the old and the new grid (boots.fpp:181-182 with GCD :388-400): the two
    z periods coincide, Lz (1 + 1/g) with g = gcd(nzt-1, nzp-1).  nzp = nz-Cz, the new physical rows."""
    import math
    g = math.gcd(nzt - 1, nzp - 1)
    return (nzt - 1) // g - 1, (nzp - 1) // g - 1


def boots_suffix(nx: int, ny: int, nzp: int) -> str:
    """boots.fpp:174: '_P' i5.5 '-' i5.5 '-' i5.5 of nx, ny, nz-Cz."""
    return "_P%05d-%05d-%05d" % (nx, ny, nzp)


def boots_prolongate(C1t: np.ndarray, nx: int, ny: int, M: int) -> np.ndarray:
    """The Fourier-space zero padding of boots.fpp:275-300, loop for loop (1-based indices as written there,
    later assignments overwrite earlier ones).  C1t[i,j,k] (nxt/2+1, nyt, m) -> B1[i,j,k] (nx/2+1, ny, M).
    As written the second loop of each direction starts one index early, so the old mode nyt/2-1 (and m/2-1)
    is copied twice -- kept."""
    nxth, nyt, m = C1t.shape
    nxt = 2 * (nxth - 1)
    fact = 1.0 / (float(nxt) * float(nyt) * float(m))
    B1 = np.zeros((nx // 2 + 1, ny, M), dtype=np.complex128)
    for j in range(1, nyt // 2 + 2):
        B1[:nxth, j - 1, 0 : m // 2 + 1] = C1t[:, j - 1, 0 : m // 2 + 1] * fact
        for k in range(M - m // 2, M + 1):
            B1[:nxth, j - 1, k - 1] = C1t[:, j - 1, k - M + m - 1] * fact
    for j in range(ny - nyt // 2, ny + 1):
        B1[:nxth, j - 1, 0 : m // 2 + 1] = C1t[:, j - ny + nyt - 1, 0 : m // 2 + 1] * fact
        for k in range(M - m // 2, M + 1):
            B1[:nxth, j - 1, k - 1] = C1t[:, j - ny + nyt - 1, k - M + m - 1] * fact
    return B1


def boots_regrid(vt: np.ndarray, nx: int, ny: int, nzp: int, ozt: int, tdir: str) -> np.ndarray:
    """One file of the BOOTS3D loop (boots.fpp:228-312) on one rank: vt[k,j,i] on the old physical grid
    (nzt, nyt, nxt) -> FC continuation with Czt points + 3-D r2c on (nxt, nyt, nzt+Czt) -> zero padding ->
    periodic 3-D c2r on (nx, ny, nzp+Czn) -> the first nzp planes (the io plan of :187 has nz-Cz planes)."""
    nzt, nyt, nxt = vt.shape
    if not (1 <= nxt <= nx and 1 <= nyt <= ny and 1 <= nzt <= nzp):
        raise ValueError("MAIN: prolongation specification incorrect")      # boots.fpp:146-160
    Czt, Czn = boots_points(nzt, nzp)
    m, M = nzt + Czt, nzp + Czn
    if not ((Czt == 0 and ozt == 0) or (Czt > 0 and ozt > 0)):                # fcgram_mod.f90:128-141
        raise ValueError("Mismatch in continuation or matching points in z direction. Aborting...")
    gt = Grid(nxt, nyt, m, Czt, ozt, tdir=tdir)
    r = np.zeros((m, nyt, nxt))
    r[:nzt] = vt
    C1t = fftp3d_real_to_complex(gt, r)
    B1 = boots_prolongate(C1t, nx, ny, M)
    gn = Grid(nx, ny, M, 0, 0)
    br = fftp3d_complex_to_real(gn, B1)
    return np.ascontiguousarray(br[:nzp])


def boots_files(idir, odir, tdir, fnlist: str, nxt: int, nyt: int, nzt: int, ozt: int, nx: int, ny: int, nzp: int):
    """The file loop of boots.fpp:228-247, 305-312: names separated by ';', read as `idir/name` (bmangle = 0),
    written as `odir/name` + suffix.  Returns the list of files written."""
    out = []
    for fname in [s.strip() for s in fnlist.split(";") if s.strip()]:
        vt = np.fromfile(os.path.join(str(idir), fname), dtype=np.float64).reshape(nzt, nyt, nxt)
        br = boots_regrid(vt, nx, ny, nzp, ozt, tdir)
        fout = os.path.join(str(odir), fname + boots_suffix(nx, ny, nzp))
        br.tofile(fout)
        out.append(fout)
    return out


# ----------------------------------------------------------------------------
# global-quantity text files            include/*/*_global.f90 and the WRITE
# statements of hdcheck / pscheck / mhdcheck / vdiagnostic / sdiagnostic / bdiagnostic
# ----------------------------------------------------------------------------
def fortran_e(x: float, w: int, d: int) -> str:
    """One `1P Ew.d` field as gfortran / ifort print it: d.dddE+ee right-justified in w columns; three-digit
    exponents drop the E (1.234560-100); NaN / Infinity by name; all stars when the field is too narrow."""
    if x != x:
        s = "NaN"
    elif x in (float("inf"), float("-inf")):
        s = "Infinity" if x > 0 else "-Infinity"
    else:
        m, e = ("%.*E" % (d, x)).split("E")
        s = m + (("E%+03d" % int(e)) if abs(int(e)) < 100 else ("%+04d" % int(e)))
    return "*" * w if len(s) > w else s.rjust(w)


def _append_row(path, fields):
    with open(path, "a") as f:
        f.write("".join(fortran_e(x, w, d) for x, w, d in fields) + "\n")


def solver_global(g: Grid, s, solver: str, odir, t: int, dt: float, bczsta: int = 0, bczend: int = 0):
    """The `<solver>_global.f90` include (hdcheck / mhdcheck with hel = 1 (and crs = 1), pscheck, vdiagnostic,
    sdiagnostic, bdiagnostic) with the reference's FORMATs: pseudospec_hd.f90:991-1001, pseudospec_phd.f90:313-318,
    pseudospec_mhd.f90:189-209, vboundary.f90:260-264, sboundary.f90:201-205, bboundary.f90:400-425."""
    tl = (t - 1) * dt
    p = lambda name: os.path.join(str(odir), name)
    mag = solver in ("MHD", "MHDBOUSS")
    sca = solver in ("BOUSS", "ROTBOUSS", "MHDBOUSS")
    if not mag:
        eng, ens, pot = hdcheck(g, s.vx, s.vy, s.vz, s.fx, s.fy, s.fz)
        _append_row(p("balance.txt"), [(tl, 13, 6), (eng, 23, 16), (ens, 23, 16), (pot, 24, 16)])
        _append_row(p("helicity.txt"), [(tl, 13, 6), (helicity(g, s.vx, s.vy, s.vz), 24, 16)])
    else:
        eng, ens, cur, engk, engm, helk, helm, crh, asq = mhdcheck(g, s.vx, s.vy, s.vz, s.ax, s.ay, s.az)
        _append_row(p("balance.txt"), [(tl, 13, 6), (eng, 23, 16), (ens, 23, 16), (cur, 23, 16)])
        _append_row(p("energy.txt"), [(tl, 13, 6), (engk, 23, 16), (engm, 23, 16)])
        _append_row(p("helicity.txt"), [(tl, 13, 6), (helk, 24, 16), (helm, 24, 16)])
        _append_row(p("cross.txt"), [(tl, 13, 6), (crh, 23, 16), (asq, 24, 16)])
    if sca:
        e1, e2, e3 = pscheck(g, s.th, s.fs)
        _append_row(p("scalar.txt"), [(tl, 13, 6), (e1, 22, 14), (e2, 22, 14), (e3, 23, 14)])
    _append_row(p("noslip_diagnostic.txt"), [(tl, 13, 6)] + [(x, 13, 6) for x in vdiagnostic(g, s.vx, s.vy, s.vz)])
    if mag:
        d = bdiagnostic(g, s.ax, s.ay, s.az, bczsta, bczend)
        for kind in ("conducting", "vacuum"):
            if kind in d:
                _append_row(p(kind + "_diagnostic.txt"), [(tl, 13, 6)] + [(x, 13, 6) for x in d[kind]])
    if sca:
        _append_row(p("scalar_constant_diagnostic.txt"), [(tl, 13, 6)] + [(x, 13, 6) for x in sdiagnostic(g, s.th)])
