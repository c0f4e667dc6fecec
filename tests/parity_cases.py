"""Backend-agnostic parity cases: the C ABI (through specter_b200.api) against the oracle on the
same seeded inputs.  test_parity_emu.py runs them on the CPU-thread emulation of the kernel sources
(small grids, no GPU); test_parity_gpu.py runs them on the nvcc-built library on a B200.

Tolerances (FP64, stated by BASELINE.json north_star): spectral fields <= 1e-11 relative per step;
diagnostics <= 1e-9 relative over 100 steps.  Single operators are held to 1e-12."""
import numpy as np

from oracle import specter_oracle as O
from specter_b200 import api

TOL_OP = 1e-12
TOL_FIELD = 1e-11
TOL_DIAG = 1e-9


# The FC-Gram table the cases build their grids with.  `cond` scales every tolerance: the continuation rows are
# sum_j dir(i,j) f(j) with max|dir| = 3.5e3 for A25-5 (the configuration of BASELINE.json, cond = 1), 29 for A15-3,
# 2.6e6 for A34-8 and 1.1e7 for A33-9, so two FP64 evaluations whose inputs differ in the last bit differ by that much
# more in the continuation rows (and, after the transform, in every spectral coefficient).
_FC = {"Cz": 25, "oz": 5, "cond": 1.0}


class fc_table:
    """with fc_table(33, 9, tables): ...  -- run cases on another of the reference's tables (tables/README.info: O = 3..9,
    C = 15..34), tolerances scaled by max|dir| relative to A25-5."""

    def __init__(self, Cz, oz, tables):
        g = O.Grid(8, 8, Cz + 2 * oz + 8, Cz, oz, tdir=tables, ord=2)
        # x4: the boundary operators chain two or three continuations; measured headroom of the operator cases on
        # A34-8 / A33-9 with the bare ratio was 5-10x (A25-5 at 1e-11: ~100x)
        self.new = {"Cz": Cz, "oz": oz, "cond": max(1.0, 4.0 * float(np.abs(g.dir).max()) / 3536.0)}

    def __enter__(self):
        self.old = dict(_FC)
        _FC.update(self.new)
        return self

    def __exit__(self, *exc):
        _FC.update(self.old)


def rel(x, y):
    d = float(np.abs(y).max())
    return float(np.abs(x - y).max()) / (d if d > 0 else 1.0) / _FC["cond"]


def pr_close(got_pr, want_pr, nph, vfields):
    """p' = dt_sub * p (mixed domain, physical rows).  It is a dt-sized difference of O(|v|) quantities (the wall
    values of v_z and the particular solution -i k.v/k^2), and it only re-enters the path as i k p'_wall added to the
    wall rows of v (vboundary.f90:195-196): its error is held to 1e-9 of its own maximum, or -- on thin grids where
    |p'| / |v| ~ 1e-5 -- to 1e-11 of the velocity maximum, the scale it is formed from and acts on."""
    a, b = got_pr[:, :, :nph], want_pr[:, :, :nph]
    err = float(np.abs(a - b).max())
    vscale = max(float(np.abs(q).max()) for q in vfields)
    pscale = float(np.abs(b).max())
    assert err <= 100 * TOL_FIELD * _FC["cond"] * pscale or err <= TOL_FIELD * _FC["cond"] * vscale, (err, pscale, vscale)


def make(lib, tables, nx, ny, nz, Cz=None, oz=None, ord=2, Lx=1.0, Ly=0.5, Lz=1.0):
    if Cz is None:
        Cz, oz = _FC["Cz"], _FC["oz"]
    elif oz is None:
        oz = _FC["oz"] if Cz else 0
    g = O.Grid(nx, ny, nz, Cz, oz, Lx=Lx, Ly=Ly, Lz=Lz, tdir=tables if Cz else "", ord=ord)
    p = api.Plan(nx, ny, nz, Cz, oz, ord=ord, Lx=Lx, Ly=Ly, Lz=Lz, tdir=tables if Cz else "", lib=lib)
    return g, p


def rand_spec(g, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(g.cshape()) + 1j * rng.standard_normal(g.cshape())


def smooth_velocity(g, seed=0):
    """A band-limited, Hermitian-consistent random field (a real field's transform)."""
    rng = np.random.default_rng(seed)
    out = []
    nph = g.nz - g.Cz
    for _ in range(3):
        r = np.zeros(g.rshape())
        r[:nph] = rng.standard_normal((nph, g.ny, g.nx))
        c = O.fftp3d_real_to_complex(g, r)
        # low-pass so the products do not alias wildly
        kmax = 0.3 * min(g.nx, g.ny) / 2
        mask = (np.abs(g.kx / g.Dkx)[:, None, None] <= kmax) & (np.abs(g.ky / g.Dky)[None, :, None] <= kmax)
        out.append(c * mask)
    return out


# ---- transforms -------------------------------------------------------------------------------
def case_fft1d_z(lib, tables, shape, Cz=None):
    g, p = make(lib, tables, *shape, Cz=Cz)
    a = rand_spec(g, 1)
    d = p.spectral(a); p.fftp1d_real_to_complex_z(d)
    assert rel(d.get(), O.fftp1d_real_to_complex_z(g, a.copy())) < TOL_OP
    d = p.spectral(a); p.fftp1d_complex_to_real_z(d)
    assert rel(d.get(), O.fftp1d_complex_to_real_z(g, a.copy())) < TOL_OP
    p.close()


def case_normalisations(lib, tables, shape):
    """normvec / normsca / normalize (pseudospec_hd.f90:1238-1286, pseudospec_phd.f90:324-368, module_dns.f90:13-42)."""
    g, p = make(lib, tables, *shape)
    v = smooth_velocity(g, 4)
    for kin, d in ((1, 2.5), (0, 0.7), (2, 1.3)):
        dev = [p.spectral(a) for a in v]
        p.normvec(*dev, d, kin)
        ref = [a.copy() for a in v]
        O.normvec(g, *ref, d, kin)
        fields_close([q.get() for q in dev], ref, tol=TOL_OP * 10)
        assert abs(p.energy(*dev, kin) / d - 1) < TOL_DIAG
    for kin, b in ((1, 0.5), (0, 3.0)):
        dev = p.spectral(v[0])
        p.normsca(dev, b, kin)
        ref = v[0].copy()
        O.normsca(g, ref, b, kin)
        fields_close([dev.get()], [ref], tol=TOL_OP * 10)
    dev = [p.spectral(a) for a in v]
    p.normalize(*dev, 0.3, 1)
    rmp = 0.3 / np.sqrt(O.energy(g, *v, 1))
    fields_close([q.get() for q in dev], [a * rmp for a in v], tol=TOL_OP * 10)
    p.close()


def case_goto_domain(lib, tables, shape):
    """goto_domain_w_boundaries / goto_3d_fourier (boundary_mod.fpp:72-194) on one and on three fields."""
    g, p = make(lib, tables, *shape)
    nph = g.nz - g.Cz
    f = [rand_spec(g, 20 + i) for i in range(3)]
    d = [p.spectral(a) for a in f]
    p.goto_domain_w_boundaries(d[0])
    p.goto_domain_w_boundaries(d[1], d[2])
    ref = [a.copy() for a in f]
    O.goto_domain_w_boundaries(g, *ref)
    for q, r in zip(d, ref):
        assert rel(q.get()[:, :, :nph], r[:, :, :nph]) < TOL_OP      # rows above are overwritten by the continuation
    p.goto_3d_fourier(d[0], d[1], d[2])
    O.goto_3d_fourier(g, *ref)
    for q, r in zip(d, ref):
        assert rel(q.get(), r) < TOL_OP
    # the scratch pool of the per-operator entries can be released and grows again on demand
    rr, out = p.real(), p.spectral()
    p.fftp3d_complex_to_real(d[0], rr)
    first = rr.get()
    p.release_scratch()
    p.release_scratch()
    p.fftp3d_complex_to_real(d[0], rr)
    assert np.array_equal(first, rr.get())
    p.fftp3d_real_to_complex(rr, out)
    p.close()


def case_fft3d(lib, tables, shape, Cz=None):
    g, p = make(lib, tables, *shape, Cz=Cz)
    rng = np.random.default_rng(2)
    r = rng.standard_normal(g.rshape())
    dr, dc = p.real(r), p.spectral()
    p.fftp3d_real_to_complex(dr, dc)
    ref = O.fftp3d_real_to_complex(g, r.copy())
    assert rel(dc.get(), ref) < TOL_OP
    p.fftp2d_real_to_complex_xy(dr, dc)
    mixed = O.fftp2d_real_to_complex_xy(g, r)
    assert rel(dc.get(), mixed) < TOL_OP
    # c2r of a Hermitian-consistent spectrum and of a generic one (FFTW c2r semantics)
    for spec in (ref, rand_spec(g, 3)):
        dc.put(spec)
        p.fftp3d_complex_to_real(dc, dr)
        assert rel(dr.get(), O.fftp3d_complex_to_real(g, spec)) < TOL_OP
        assert rel(dc.get(), spec) == 0.0  # unlike the reference, the input is preserved
        p.fftp2d_complex_to_real_xy(dc, dr)
        assert rel(dr.get(), O.fftp2d_complex_to_real_xy(g, spec)) < TOL_OP
    # round trip = identity x N on the physical rows (tests/fft.f90:45-67)
    dc.put(ref); p.fftp3d_complex_to_real(dc, dr)
    nph = g.nz - g.Cz
    back = dr.get() / g.N  # rows >= nph hold the (large, |dir| ~ 3e3) continuation of random data
    assert np.abs(back[:nph] - r[:nph]).max() < 1e-12 * np.abs(back).max()
    p.close()


def case_fft_known_answer(lib, tables, n=32):
    """tests/fft.f90:38-43 on a periodic box."""
    g, p = make(lib, tables, n, n, n, Cz=0, oz=0, Ly=1.0)
    r = O.analytic_field(g, "sin")
    dr, dc = p.real(r), p.spectral()
    p.fftp3d_real_to_complex(dr, dc)
    c = dc.get()
    assert abs(abs(c[4, 8, 6]) - n ** 3 / 8) < 1e-9
    p.close()


# ---- spectral operators ---------------------------------------------------------------------
def case_spectral_ops(lib, tables, shape):
    g, p = make(lib, tables, *shape)
    a, b = rand_spec(g, 4), rand_spec(g, 5)
    da, db, dc = p.spectral(a), p.spectral(b), p.spectral()
    for d_ in (1, 2, 3):
        p.derivk(da, dc, d_)
        assert rel(dc.get(), O.derivk(g, a, d_)) < TOL_OP
        p.curlk(da, db, dc, d_)
        assert rel(dc.get(), O.curlk(g, a, b, d_)) < TOL_OP
    p.laplak(da, dc)
    assert rel(dc.get(), O.laplak(g, a)) < TOL_OP
    p.laplak(da, da)  # in-place aliasing as in hd_rkstep2.f90:14
    assert rel(da.get(), O.laplak(g, a)) < TOL_OP
    db.put(b); p.fc_filter(db)
    assert rel(db.get(), O.fc_filter(g, b.copy())) < TOL_OP
    p.close()


def case_nonlinear(lib, tables, shape):
    g, p = make(lib, tables, *shape)
    v = smooth_velocity(g, 6)
    dv = [p.spectral(q) for q in v]
    out = [p.spectral() for _ in range(3)]
    p.gradre(*dv, *out)
    ref = O.gradre(g, *v)
    scale = max(np.abs(q).max() for q in ref)
    for q, r in zip(out, ref):
        assert np.abs(q.get() - r).max() / scale < TOL_FIELD
    p.prodre(*dv, *out)
    ref = O.prodre(g, *v)
    scale = max(np.abs(q).max() for q in ref)
    for q, r in zip(out, ref):
        assert np.abs(q.get() - r).max() / scale < TOL_FIELD
    p.close()


# ---- boundary -------------------------------------------------------------------------------
def case_projection(lib, tables, shape):
    g, p = make(lib, tables, *shape)
    v = smooth_velocity(g, 7)
    for (t, s, e) in ((1, 0, 0), (0, 0, 0), (0, 2, 2), (0, 0, 2)):
        dv = [p.spectral(q) for q in v]
        dd = p.spectral()
        p.sol_project(*dv, dd, t, s, e)
        rv = [q.copy() for q in v]
        rd = O.sol_project(g, *rv, t, s, e)
        scale = max(np.abs(q).max() for q in rv)
        for q, r in zip(dv, rv):
            assert np.abs(q.get() - r).max() / scale < TOL_FIELD * _FC["cond"], (t, s, e)
        nph = g.nz - g.Cz
        assert rel(dd.get()[:, :, :nph], rd[:, :, :nph]) < TOL_FIELD, (t, s, e)
        for q in dv + [dd]:
            q.free()
    # no-slip + projection, both the first (o == ord) and a later substep, moving walls
    pr = rand_spec(g, 8) * 1e-3
    for o, zs, ze in ((2, (0.0, 0.0), (0.0, 0.0)), (1, (0.3, -0.1), (-0.2, 0.4))):
        dv = [p.spectral(q) for q in v]
        dp = p.spectral(pr)
        p.v_imposebc_and_project(*dv, dp, o, zs, ze)
        rv = [q.copy() for q in v]
        rp = O.v_imposebc_and_project(g, *rv, pr.copy(), o, zs, ze)
        scale = max(np.abs(q).max() for q in rv)
        for q, r in zip(dv, rv):
            assert np.abs(q.get() - r).max() / scale < TOL_FIELD * _FC["cond"]
        nph = g.nz - g.Cz
        assert rel(dp.get()[:, :, :nph], rp[:, :, :nph]) < TOL_FIELD
    p.close()


def case_diagnostics(lib, tables, shape):
    g, p = make(lib, tables, *shape)
    v = smooth_velocity(g, 9)
    f = smooth_velocity(g, 10)
    dv = [p.spectral(q) for q in v]
    df = [p.spectral(q) for q in f]
    for kin in (0, 1, 2):
        assert abs(p.energy(*dv, kin) / O.energy(g, *v, kin) - 1) < TOL_DIAG
    for kin in (0, 1):
        assert abs(p.cross(*dv, *df, kin) / O.cross(g, *v, *f, kin) - 1) < TOL_DIAG
    assert abs(p.divergence(*dv) / O.divergence(g, *v) - 1) < TOL_DIAG
    got, ref = p.vdiagnostic(*dv), O.vdiagnostic(g, *v)
    assert np.allclose(got, ref, rtol=TOL_DIAG, atol=0)
    got, ref = p.hdcheck(*dv, *df), O.hdcheck(g, *v, *f)
    assert np.allclose(got, ref, rtol=TOL_DIAG, atol=0)
    p.close()


# ---- the RK substep ---------------------------------------------------------------------------
def hd_fields_close(got, s, g, tol=TOL_FIELD):
    nph = g.nz - g.Cz
    scale = max(np.abs(q).max() for q in (s.vx, s.vy, s.vz))
    for q, r in zip(got[:3], (s.vx, s.vy, s.vz)):
        err = np.abs(q - r).max() / scale
        assert err < tol * _FC["cond"], err
    pr_close(got[3], s.pr, nph, (s.vx, s.vy, s.vz))


def case_hd_substeps(lib, tables, shape, ord=2, nsteps=1, impl=0, dt=1e-3, nu=1e-3, walls=((0., 0.), (0., 0.))):
    """Per-substep spectral fields against the oracle (hd_rkstep2.f90), state device-resident."""
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_hd_state(g)
    p.hd_put_state(s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)
    for _ in range(nsteps):
        p.hd_rkstep1()
        C = [s.vx.copy(), s.vy.copy(), s.vz.copy()]
        for o in range(ord, 0, -1):
            p.hd_rkstep2(o, dt, nu, walls[0], walls[1], impl)
            O.hd_rkstep2(g, s, *C, o, dt, nu, walls[0], walls[1])
            hd_fields_close(p.hd_get_state(), s, g)
    p.close()


def case_hd_step_host(lib, tables, shape, ord=2, pinned=False, nsteps=1, inflight=False):
    """The host-buffer entry (H2D, ord substeps, D2H) used for the end-to-end number.  pinned: page-locked host arrays as in
    bench.py's e2e leg (copies from / into them are truly asynchronous), the forcing uploaded by the first call only and
    kept resident (NULL afterwards)."""
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_hd_state(g)
    h = [q.copy() for q in (s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)]
    if pinned:
        h = [p.pinned_like(q) for q in h]
    if inflight:
        # a substep of some OTHER state is still in flight (nothing waited for it) when the host-buffer step overwrites the
        # plan-owned state: the entry has to wait for it first (sx_hd_step_host: "nothing in flight still reads the state
        # that is overwritten")
        p.hd_put_state(*[0.5 * q[::-1].copy() for q in (s.vx, s.vy, s.vz)], s.pr, s.fx, s.fy, s.fz)
        p.hd_rkstep1()
        p.hd_rkstep2(ord, 1e-3, 1e-3)
    for k in range(nsteps):
        if k == 0:
            p.hd_step_host(*h, 1e-3, 1e-3)
        else:
            p.hd_step_host(*h[:4], None, None, None, 1e-3, 1e-3)
        O.hd_step(g, s, 1e-3, 1e-3)
        hd_fields_close(h[:4], s, g)
    p.close()


def case_hd_diagnostics_100(lib, tables, shape, nsteps=100, ord=2, impl=0, dt=1e-3, nu=1e-3, golden=None):
    """Energy / dissipation / divergence over `nsteps` steps (hd_global.f90) within 1e-9 relative of
    the oracle's (or of committed golden values computed by the oracle)."""
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_hd_state(g)
    p.hd_put_state(s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)
    v = [p.hd_field(i) for i in range(3)]
    f = [p.hd_field(4 + i) for i in range(3)]
    rows = []
    for t in range(nsteps):
        p.hd_step(dt, nu, impl=impl)
        if (t + 1) % 10 == 0 or t == nsteps - 1:
            rows.append((t + 1,) + p.hdcheck(*v, *f) + p.vdiagnostic(*v)[:1])
    if golden is None:
        ref = []
        s2 = O.make_hd_state(g)
        for t in range(nsteps):
            O.hd_step(g, s2, dt, nu)
            if (t + 1) % 10 == 0 or t == nsteps - 1:
                ref.append((t + 1,) + O.hdcheck(g, s2.vx, s2.vy, s2.vz, s2.fx, s2.fy, s2.fz)
                           + (O.divergence(g, s2.vx, s2.vy, s2.vz),))
    else:
        ref = golden
    rows, ref = np.array(rows), np.array(ref)
    assert rows.shape == ref.shape
    # energy, enstrophy-like column, injection: relative 1e-9
    assert np.allclose(rows[:, 1:4], ref[:, 1:4], rtol=TOL_DIAG, atol=0), np.abs(rows[:, 1:4] / ref[:, 1:4] - 1).max()
    # divergence is a residual at the FC-accuracy floor (~1e-10 of <v^2>): compare on the scale of
    # the energy it is a residual of
    assert np.abs(rows[:, 4] - ref[:, 4]).max() <= TOL_DIAG * np.abs(ref[:, 1]).max()
    p.close()
    return rows


# ---- Boussinesq / MHD operators and substeps -------------------------------------------------------
def fields_close(got, ref, tol=TOL_FIELD, rows=None):
    scale = max(np.abs(r).max() for r in ref)
    for q, r in zip(got, ref):
        if rows is not None:
            q, r = q[:, :, :rows], r[:, :, :rows]
        err = np.abs(q - r).max() / (scale if scale > 0 else 1.0)
        assert err < tol * _FC["cond"], err


def phys_close(g, got, ref, tol=TOL_FIELD):
    """Compare in the mixed (z,ky,kx) domain on the physical rows -- the well-conditioned representation.
    The spectral coefficients of a twice re-continued field (BOUSS theta: s_imposebc, fc_filter, then the 3-D
    round trip of bouss_rkstep2.f90:53-59) carry FC-Gram continuation noise amplified by |dir| ~ 3.5e3 per
    continuation: perturbing the ORACLE's own input by 1e-16 relative moves its spectral theta by 1.0e-10
    (measured, 32x16x64) while the physical rows move by < 1e-13."""
    nph = g.nz - g.Cz
    scale = max(np.abs(O.fftp1d_complex_to_real_z(g, r.copy())[:, :, :nph]).max() for r in ref)
    for q, r in zip(got, ref):
        a = O.fftp1d_complex_to_real_z(g, np.array(q, dtype=np.complex128))[:, :, :nph]
        b = O.fftp1d_complex_to_real_z(g, r.copy())[:, :, :nph]
        err = np.abs(a - b).max() / scale
        assert err < tol * _FC["cond"], err


TOL_RECONTINUED = 2e-9   # spectral coefficients of a twice re-continued field, see phys_close


def case_advect_vector(lib, tables, shape):
    """advect (pseudospec_phd.f90:23-113) and vector (pseudospec_mhd.f90:22-105)."""
    g, p = make(lib, tables, *shape)
    v = smooth_velocity(g, 11)
    b = smooth_velocity(g, 12)
    dv = [p.spectral(q) for q in v]
    db = [p.spectral(q) for q in b]
    out = [p.spectral() for _ in range(3)]
    p.advect(*dv, db[0], out[0])
    fields_close([out[0].get()], [O.advect(g, *v, b[0])])
    p.vector(*dv, *db, *out)
    fields_close([q.get() for q in out], O.vector(g, *v, *b))
    a = smooth_velocity(g, 13)[0]
    da = p.spectral(a)
    for kin in (0, 1):
        assert abs(p.variance(da, kin) / O.variance(g, a, kin) - 1) < TOL_DIAG
    p.close()


def case_scalar_vecpot_bc(lib, tables, shape):
    """s_imposebc (sboundary.f90:67-119) and a_imposebc_and_project (bboundary.f90:100-189, conducting)."""
    g, p = make(lib, tables, *shape)
    g.load_neumann()
    th = smooth_velocity(g, 14)[0]
    dth = p.spectral(th)
    p.s_imposebc(dth)
    r = th.copy(); O.s_imposebc(g, r)
    fields_close([dth.get()], [r])
    a = smooth_velocity(g, 15)
    da = [p.spectral(q) for q in a]
    dph = p.spectral()
    p.a_imposebc_and_project(*da, dph)
    ra = [q.copy() for q in a]
    rph = O.a_imposebc_and_project(g, *ra)
    fields_close([q.get() for q in da], ra)
    nph = g.nz - g.Cz
    fields_close([dph.get()], [rph], rows=nph)
    p.close()


def case_bouss_substeps(lib, tables, shape, ord=2, nsteps=1, impl=0, dt=1e-3, nu=1e-3, kappa=1e-3, xmom=1.0, xtemp=1.0):
    """Per-substep spectral fields against the oracle (bouss_rkstep2.f90:3-59)."""
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_bouss_state(g)
    p.bouss_put_state(s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)
    nph = g.nz - g.Cz
    for _ in range(nsteps):
        p.bouss_rkstep1()
        C = [s.vx.copy(), s.vy.copy(), s.vz.copy(), s.th.copy()]
        for o in range(ord, 0, -1):
            p.bouss_rkstep2(o, dt, nu, kappa, xmom, xtemp, impl=impl)
            O.bouss_rkstep2(g, s, *C, o, dt, nu, kappa, xmom, xtemp)
            got = p.bouss_get_state()
            fields_close(got[:3], (s.vx, s.vy, s.vz))
            phys_close(g, [got[4]], [s.th])
            fields_close([got[4]], [s.th], tol=TOL_RECONTINUED)
            pr_close(got[3], s.pr, nph, (s.vx, s.vy, s.vz))
    p.close()


def case_mhd_substeps(lib, tables, shape, ord=2, nsteps=1, impl=0, dt=1e-3, nu=1e-3, mu=5e-3, b0=(0.0, 0.0, 0.0)):
    """Per-substep spectral fields against the oracle (mhd_rkstep2.f90:3-84), conducting walls."""
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_mhd_state(g)
    p.mhd_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.fx, s.fy, s.fz, s.mx, s.my, s.mz)
    nph = g.nz - g.Cz
    for _ in range(nsteps):
        p.mhd_rkstep1()
        C = [q.copy() for q in (s.vx, s.vy, s.vz, s.ax, s.ay, s.az)]
        for o in range(ord, 0, -1):
            p.mhd_rkstep2(o, dt, nu, mu, b0, impl)
            O.mhd_rkstep2(g, s, *C, o, dt, nu, mu, b0)
            got = p.mhd_get_state()
            fields_close(got[:3], (s.vx, s.vy, s.vz))
            fields_close(got[4:7], (s.ax, s.ay, s.az))
            pr_close(got[3], s.pr, nph, (s.vx, s.vy, s.vz))
            fields_close([got[7]], [s.ph], rows=nph, tol=100 * TOL_FIELD)
    p.close()


# ---- field files, output and restart (binary_io.f90, specter.fpp:1005-1053 / 886-912) ----------------------------
def case_io_output_restart(lib, tables, shape, tmpdir, ord=2, dt=1e-3, nu=1e-3):
    import os
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_hd_state(g)
    p.hd_put_state(s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)
    p.hd_step(dt, nu)
    O.hd_step(g, s, dt, nu)
    ours, ref = os.path.join(str(tmpdir), "ours"), os.path.join(str(tmpdir), "ref")
    os.makedirs(ours), os.makedirs(ref)
    p.hd_output(ours, "0001", dt, outs=1)
    O.hd_output(g, s, ref, "0001", dt, outs=1)
    nph = g.nz - g.Cz
    for name in ("vx", "vy", "vz", "wx", "wy", "wz", "pr"):
        a = np.fromfile(O.io_path(ours, name, "0001"))
        b = np.fromfile(O.io_path(ref, name, "0001"))
        assert a.size == b.size == g.nx * g.ny * nph, (name, a.size)       # the physical box, nothing else
        tol = 100 * TOL_FIELD if name == "pr" else TOL_FIELD                  # p = p'/dt: see hd_fields_close
        assert rel(a, b) < tol, (name, rel(a, b))
    # the raw io entry points: a device real array round-trips through a file bit for bit on the physical planes
    rng = np.random.default_rng(5)
    r = rng.standard_normal(g.rshape())
    d = p.real(r)
    p.io_write(d, ours, "rt", "0007")
    assert np.array_equal(np.fromfile(O.io_path(ours, "rt", "0007")).reshape(nph, g.ny, g.nx), r[:nph])
    d2 = p.real()
    p.io_read(d2, ours, "rt", "0007")
    back = d2.get()
    assert np.array_equal(back[:nph], r[:nph]) and not back[nph:].any()
    # restart from the REFERENCE-format files written by the oracle: state equals the oracle's restart
    p.hd_restart(ref, "0001", dt)
    got = p.hd_get_state()
    want = O.hd_restart(g, ref, "0001", dt)
    scale = max(np.abs(q).max() for q in want[:3])
    for q, w in zip(got[:3], want[:3]):
        assert np.abs(q - w).max() / scale < TOL_FIELD
    assert rel(got[3][:, :, :nph], want[3][:, :, :nph]) < 100 * TOL_FIELD
    p.close()


# ---- vacuum walls, reconstructions, ROTBOUSS / MHDBOUSS, remaining diagnostics ----------------
PERIODIC4 = ["periodic"] * 4
B_KIND = {0: "conducting", 1: "vacuum"}


def case_wall_reconstructions(lib, tables, shape):
    """neumann_reconstruct (fcgram_mod.f90:368-511) and robin_reconstruct (:514-644), z branches."""
    g, p = make(lib, tables, *shape)
    g.load_neumann()
    f = smooth_velocity(g, 21)[0]
    O.fftp1d_complex_to_real_z(g, f)                      # any mixed-domain data will do
    for boun in (5, 6):
        for order in (1, 2):
            d = p.spectral(f)
            p.neumann_reconstruct(d, boun, order)
            r = f.copy(); O.neumann_reconstruct(g, r, boun, order)
            assert rel(d.get(), r) < TOL_OP
        d = p.spectral(f)
        p.robin_reconstruct(d, boun)
        r = f.copy(); O.robin_reconstruct(g, r, boun, g.khom)
        assert rel(d.get(), r) < TOL_OP
    p.close()


def case_vacuum_walls(lib, tables, shape):
    """a_imposebc_and_project (bboundary.f90:100-189) with vacuum walls (insulating_z :294-344) and the mixed
    conducting / vacuum channel; robcheck and bdiagnostic on the result."""
    nph = shape[2] - 25
    for (bs, be) in ((1, 1), (0, 1), (0, 0)):
        g, p = make(lib, tables, *shape)
        g.load_neumann()
        p.setup_bc("b", PERIODIC4 + [B_KIND[bs], " " + B_KIND[be].upper() + " "])   # preprocess: trim + lower case
        a = smooth_velocity(g, 15)
        da = [p.spectral(q) for q in a]
        dph = p.spectral()
        p.a_imposebc_and_project(*da, dph)
        ra = [q.copy() for q in a]
        rph = O.a_imposebc_and_project_bc(g, *ra, bs, be)
        fields_close([q.get() for q in da], ra)
        fields_close([dph.get()], [rph], rows=nph)
        got, want = p.robcheck(*da), O.robcheck(g, *ra)
        # the residuals are differences of O(1) terms cancelling to ~1e-5: compare on the scale of the terms
        scale = O.robcheck(g, *a)[0]
        assert max(abs(x - y) for x, y in zip(got, want)) < TOL_DIAG * scale, (got, want)
        gd, wd = p.bdiagnostic(*da), O.bdiagnostic(g, *ra, bs, be)
        assert sorted(gd) == sorted(wd)
        for key in wd:
            ref_scale = max(max(abs(x) for x in wd[key]), scale)
            assert max(abs(x - y) for x, y in zip(gd[key], wd[key])) < 10 * TOL_DIAG * ref_scale, (key, gd[key], wd[key])
        p.close()
    # the combination laplace_z refuses fails here as it does there, and unknown kinds are rejected
    g, p = make(lib, tables, *shape)
    p.setup_bc("b", PERIODIC4 + ["vacuum", "conducting"])
    da = [p.spectral(q) for q in smooth_velocity(g, 15)]
    try:
        p.a_imposebc_and_project(*da, p.spectral())
        raise AssertionError("vacuum bottom / conducting top must be refused")
    except api.SpecterError as e:
        assert "Unsupported BC combination" in str(e)
    for field, kinds in (("b", PERIODIC4 + ["noslip", "vacuum"]), ("v", PERIODIC4 + ["constant", "noslip"]),
                         ("s", ["constant"] * 6), ("q", PERIODIC4 + ["vacuum", "vacuum"])):
        try:
            p.setup_bc(field, kinds)
            raise AssertionError((field, kinds))
        except api.SpecterError:
            pass
    p.close()


def case_more_diagnostics(lib, tables, shape):
    """helicity, product, pscheck, maxabs, mhdcheck, sdiagnostic against the oracle."""
    g, p = make(lib, tables, *shape)
    v, b = smooth_velocity(g, 31), smooth_velocity(g, 32)
    dv, db = [p.spectral(q) for q in v], [p.spectral(q) for q in b]

    def close(x, y, scale=None):
        s = abs(y) if scale is None else scale
        assert abs(x - y) <= TOL_DIAG * max(s, 1e-300), (x, y)

    hscale = O.energy(g, *v, 1) ** 0.5 * O.energy(g, *v, 0) ** 0.5     # helicity can cancel to ~0
    close(p.helicity(*dv), O.helicity(g, *v), hscale)
    close(p.product(dv[0], db[1]), O.product(g, v[0], b[1]), (O.variance(g, v[0], 1) * O.variance(g, b[1], 1)) ** 0.5)
    got, want = p.pscheck(dv[0], db[0]), O.pscheck(g, v[0], b[0])
    close(got[0], want[0]); close(got[1], want[1])
    close(got[2], want[2], (want[0] * O.variance(g, b[0], 1)) ** 0.5)
    for kin in (0, 1, 2):
        close(p.maxabs(*dv, kin), O.maxabs(g, *v, kin))
    got, want = p.mhdcheck(*dv, *db), O.mhdcheck(g, *v, *b)
    for i in (0, 1, 2, 3, 4, 8):
        close(got[i], want[i])
    close(got[5], want[5], hscale)
    close(got[6], want[6], O.energy(g, *b, 1) ** 0.5 * O.energy(g, *b, 0) ** 0.5)
    close(got[7], want[7], (want[3] * O.energy(g, O.derivk(g, b[0], 1), O.derivk(g, b[1], 2), O.derivk(g, b[2], 3), 1)) ** 0.5)
    got, want = p.sdiagnostic(dv[2]), O.sdiagnostic(g, v[2])
    close(got[0], want[0]); close(got[1], want[1])
    p.close()


def case_rotbouss_substeps(lib, tables, shape, ord=2, nsteps=1, impl=0, dt=1e-3, nu=1e-3, kappa=1e-3, xmom=1.0,
                           xtemp=1.0, omega=(0.3, -0.2, 1.5), walls=((0., 0.), (0., 0.))):
    """Per-substep spectral fields against the oracle (rotbouss_rkstep2.f90:3-56)."""
    g, p = make(lib, tables, *shape, ord=ord)
    s = O.make_bouss_state(g)
    p.bouss_put_state(s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)
    nph = g.nz - g.Cz
    for _ in range(nsteps):
        p.bouss_rkstep1()
        C = [s.vx.copy(), s.vy.copy(), s.vz.copy(), s.th.copy()]
        for o in range(ord, 0, -1):
            p.rotbouss_rkstep2(o, dt, nu, kappa, xmom, xtemp, omega, walls[0], walls[1], impl=impl)
            O.rotbouss_rkstep2(g, s, *C, o, dt, nu, kappa, xmom, xtemp, omega, walls[0], walls[1])
            got = p.bouss_get_state()
            fields_close(got[:3], (s.vx, s.vy, s.vz))
            fields_close([got[4]], [s.th])
            pr_close(got[3], s.pr, nph, (s.vx, s.vy, s.vz))
    p.close()


def case_mhdbouss_substeps(lib, tables, shape, ord=2, nsteps=1, dt=1e-3, nu=1e-3, mu=5e-3, kappa=1e-3, xmom=1.0,
                           xtemp=1.0, b0=(0.0, 0.0, 0.1), bc=(0, 0), impl=0):
    """Per-substep spectral fields against the oracle (mhdbouss_rkstep2.f90:3-106)."""
    g, p = make(lib, tables, *shape, ord=ord)
    g.load_neumann()
    p.setup_bc("b", PERIODIC4 + [B_KIND[bc[0]], B_KIND[bc[1]]])
    s = O.make_mhdbouss_state(g)
    p.mhdbouss_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.th, s.fx, s.fy, s.fz, s.mx, s.my, s.mz, s.fs)
    nph = g.nz - g.Cz
    for _ in range(nsteps):
        p.mhdbouss_rkstep1()
        C = [q.copy() for q in (s.vx, s.vy, s.vz, s.th, s.ax, s.ay, s.az)]
        for o in range(ord, 0, -1):
            p.mhdbouss_rkstep2(o, dt, nu, mu, kappa, xmom, xtemp, b0, impl=impl)
            O.mhdbouss_rkstep2(g, s, *C, o, dt, nu, mu, kappa, xmom, xtemp, b0, bc[0], bc[1])
            got = p.mhdbouss_get_state()
            fields_close(got[:3], (s.vx, s.vy, s.vz))
            fields_close(got[4:7], (s.ax, s.ay, s.az))
            phys_close(g, [got[8]], [s.th])
            fields_close([got[8]], [s.th], tol=TOL_RECONTINUED)
            pr_close(got[3], s.pr, nph, (s.vx, s.vy, s.vz))
            fields_close([got[7]], [s.ph], rows=nph, tol=100 * TOL_FIELD)
    p.close()


def case_solver_output_restart(lib, tables, shape, tmpdir, dt=1e-3):
    """sx_output / sx_restart for the MHDBOUSS state (every field family at once) and benchmark.txt."""
    import os
    g, p = make(lib, tables, *shape)
    g.load_neumann()
    s = O.make_mhdbouss_state(g)
    p.stage_timing(True)
    p.mhdbouss_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.th, s.fx, s.fy, s.fz, s.mx, s.my, s.mz, s.fs)
    p.mhdbouss_step(dt, 1e-3, 5e-3, 1e-3)
    O.mhdbouss_step(g, s, dt, 1e-3, 5e-3, 1e-3)
    ours, ref = os.path.join(str(tmpdir), "ours"), os.path.join(str(tmpdir), "ref")
    os.makedirs(ours), os.makedirs(ref)
    p.output("MHDBOUSS", ours, "0003", dt, outs=2)
    O.solver_output(g, s, ref, "0003", dt, outs=2)
    nph = g.nz - g.Cz
    names = ["vx", "vy", "vz", "wx", "wy", "wz", "pr", "th", "ax", "ay", "az", "bx", "by", "bz", "jx", "jy", "jz", "ph"]
    assert sorted(os.listdir(ours)) == sorted(os.listdir(ref)) == sorted(f"{n}.0003.out" for n in names)
    for name in names:
        a = np.fromfile(O.io_path(ours, name, "0003"))
        b = np.fromfile(O.io_path(ref, name, "0003"))
        assert a.size == b.size == g.nx * g.ny * nph, (name, a.size)
        tol = 100 * TOL_FIELD if name in ("pr", "ph") else (TOL_RECONTINUED if name == "th" else 10 * TOL_FIELD)
        assert rel(a, b) < tol, (name, rel(a, b))
    p.restart("MHDBOUSS", ref, "0003", dt)
    got = p.mhdbouss_get_state()
    want = O.solver_restart(g, ref, "0003", dt, scalar=True, magnetic=True)
    for q, n in zip(got, ("vx", "vy", "vz", "pr", "ax", "ay", "az", "ph", "th")):
        w = want[n]
        if n in ("pr", "ph"):
            assert rel(q[:, :, :nph], w[:, :, :nph]) < 100 * TOL_FIELD, n
        else:
            assert rel(q, w) < TOL_FIELD, (n, rel(q, w))
    try:
        p.output("GHOST", ours, "0003", dt)
        raise AssertionError("unknown solver must be refused")
    except api.SpecterError as e:
        assert "unknown solver" in str(e)
    # benchmark.txt: header once, one row per call, the reference's column count (specter.fpp:1207-1218)
    bench = os.path.join(str(tmpdir), "benchmark.txt")
    p.benchmark_write(bench, 1, nth=1, tcpu=2.0, tomp=2.0, twtime=2.0)
    p.benchmark_write(bench, 2, nth=1, tcpu=2.0, tomp=2.0, twtime=2.0)
    lines = open(bench).read().splitlines()
    assert len(lines) == 3 and lines[0].split()[:7] == ["#", "nx", "ny", "nz", "nsteps", "nprocs", "nth"]
    row = lines[2].split()
    assert len(row) == 16 and [int(x) for x in row[:6]] == [g.nx, g.ny, g.nz, 2, 1, 1]
    vals = [float(x) for x in row[6:]]
    assert vals[0] == vals[1] == vals[2] == 1.0                # totals / nsteps
    tfft, ttra, tcom, tcont, tneu, trob, ttot = vals[3:]
    assert tfft > 0 and ttot >= tfft and ttra == tcont == tneu == trob == 0.0 and tcom == 0.0
    p.close()


# ---- committed golden fixtures (tests/golden/make_golden.py) -----------------------------------------------------
def golden_solver_runs(step_fns):
    """Shared by the CPU pin of the oracle and the GPU parity test: step_fns maps a tag of solvers32_step1.npz to a
    callable returning {field name: array (nxl, ny, nz)} and optional diagnostics after one RK2 step."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solvers32_step1.npz"))
    sub = (slice(None, None, 4), slice(None, None, 4), slice(None, None, 4))
    for tag, fn in step_fns.items():
        fields, diags = fn()
        names = sorted(fields)
        scale_v = max(np.abs(gold[f"{tag}_{n}"]).max() for n in names if n.startswith("v"))
        scale_a = max([np.abs(gold[f"{tag}_{n}"]).max() for n in names if n.startswith("a")] or [1.0])
        for n in names:
            want = gold[f"{tag}_{n}"]
            scale = scale_a if n.startswith("a") else (np.abs(want).max() if n == "th" else scale_v)
            tol = TOL_RECONTINUED if n == "th" else TOL_FIELD
            err = np.abs(fields[n][sub] - want).max() / scale
            assert err < tol, (tag, n, err)
        for key, val in (diags or {}).items():
            want = gold[f"{tag}_{key}"]
            ref = np.abs(want).max()
            assert np.abs(np.asarray(val) - want).max() < 10 * TOL_DIAG * ref, (tag, key, val, want)


def case_golden_solvers(lib, tables):
    """The CUDA path against the committed goldens (one RK2 step of BOUSS, ROTBOUSS, MHD, MHDBOUSS with conducting
    and with vacuum walls on 32x32x64)."""
    shape, om, b0 = (32, 32, 64), (0.3, -0.2, 1.5), (0.0, 0.0, 0.1)

    def bouss(rot):
        def run():
            g, p = make(lib, tables, *shape)
            s = O.make_bouss_state(g)
            p.bouss_put_state(s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)
            if rot:
                p.rotbouss_step(1e-3, 1e-3, 1e-3, omega=om)
            else:
                p.bouss_step(1e-3, 1e-3, 1e-3)
            got = p.bouss_get_state()
            d = None if rot else {"pscheck": p.pscheck(p.bouss_field(10), p.bouss_field(2))}
            p.close()
            return dict(vx=got[0], vy=got[1], vz=got[2], th=got[4]), d
        return run

    def mhdbouss(bc):
        def run():
            g, p = make(lib, tables, *shape)
            g.load_neumann()
            p.setup_bc("b", PERIODIC4 + [B_KIND[bc[0]], B_KIND[bc[1]]])
            s = O.make_mhdbouss_state(g)
            p.mhdbouss_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.th, s.fx, s.fy, s.fz, s.mx, s.my, s.mz, s.fs)
            p.mhdbouss_step(1e-3, 1e-3, 5e-3, 1e-3, b0=b0)
            got = p.mhdbouss_get_state()
            f = [p.mhdbouss_field(i) for i in (0, 1, 2, 10, 11, 12)]
            bd = p.bdiagnostic(*f[3:])
            d = {"mhdcheck": p.mhdcheck(*f), "bdiag": bd["conducting" if bc == (0, 0) else "vacuum"]}
            p.close()
            return dict(vx=got[0], vy=got[1], vz=got[2], ax=got[4], ay=got[5], az=got[6], th=got[8]), d
        return run

    def mhd():
        g, p = make(lib, tables, *shape)
        s = O.make_mhd_state(g)
        p.mhd_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.fx, s.fy, s.fz, s.mx, s.my, s.mz)
        p.mhd_step(1e-3, 1e-3, 5e-3)
        got = p.mhd_get_state()
        p.close()
        return dict(vx=got[0], vy=got[1], vz=got[2], ax=got[4], ay=got[5], az=got[6]), None

    golden_solver_runs({"bouss": bouss(False), "rotbouss": bouss(True), "mhd": mhd, "mhdbouss": mhdbouss((0, 0)),
                        "mhdvacbouss": mhdbouss((1, 1))})


def case_golden_hd_step1(lib, tables):
    """Config 1 (HD 64^3 RK2): the spectral fields after the first step against hd64_step1.npz."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hd64_step1.npz"))
    g, p = make(lib, tables, 64, 64, 64)
    s = O.make_hd_state(g)
    p.hd_put_state(s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)
    p.hd_step(1e-3, 1e-3)
    got = p.hd_get_state()
    p.close()
    scale = max(np.abs(gold[n]).max() for n in ("vx", "vy", "vz"))
    for q, n in zip(got[:3], ("vx", "vy", "vz")):
        assert np.abs(q[::4, ::4, ::2] - gold[n]).max() / scale < TOL_FIELD, n
    nph = 64 - 25
    a, b = got[3][::4, ::4, ::2], gold["pr"]
    k = (nph + 1) // 2
    assert rel(a[:, :, :k], b[:, :, :k]) < 100 * TOL_FIELD


# ---- BASELINE.json's full size: size-independent properties (no oracle at 512^3) -------------------------------------
def case_full_size_properties(lib, tables, shape=(512, 512, 512), ord=4, dt=2e-4, nu=1e-3):
    """HD 512^3 RK4 (BASELINE configs[1]) through properties that need no oracle:
    (1) 3-D transform round trip = N x identity on the physical rows (tests/fft.f90:45-67);
    (2) Parseval: energy(kin=1) of the transform equals the mean square of the real field (tests/energy.f90);
    (3) the fused substep and the per-operator composition -- two independent implementations, each held to the
        oracle at small sizes -- agree to the per-step field tolerance;
    (4) after a substep the field is solenoidal and the wall-normal velocity vanishes (vdiagnostic)."""
    import bench                                  # the synthetic initial condition of the bench (product API only)
    nx, ny, nz = shape
    p = api.Plan(nx, ny, nz, 25, 5, ord=ord, Lx=1.0, Ly=1.0, Lz=1.0, tdir=tables, lib=lib)
    nph = nz - 25
    N = float(nx) * ny * nz
    rng = np.random.default_rng(11)
    r = np.zeros(p.rshape)
    r[:nph] = rng.standard_normal((nph, ny, nx))
    # no x-Nyquist content: `energy' weights every kx > 0 plane by 2, the self-conjugate nx/2 plane included
    # (pseudospec_hd.f90:602-628), which is only right for fields without it -- as in the reference's own test
    r[:nph] = 0.5 * (r[:nph] + np.roll(r[:nph], 1, axis=2))
    dr, dc, dr2 = p.real(r), p.spectral(), p.real()
    p.fftp3d_real_to_complex(dr, dc)
    p.fftp3d_complex_to_real(dc, dr2)
    back = dr2.get()
    assert np.abs(back[:nph] / N - r[:nph]).max() < 1e-12 * np.abs(back[:nph] / N).max()
    zero = p.spectral(np.zeros(p.cshape, dtype=np.complex128))
    eng = p.energy(dc, zero, zero, 1)
    assert abs(eng / float(np.mean(r[:nph] ** 2)) - 1) < 1e-10, eng
    del back
    for d in (dr, dr2, dc, zero):
        d.free()
    bench.device_state(p, "hd")       # the bench's synthetic state, built in place on the device
    st = p.hd_get_state() + [p.hd_field(4 + i).get() for i in range(3)]
    outs = []
    for impl in (0, 1):
        p.hd_put_state(*st)
        p.hd_rkstep1()
        p.hd_rkstep2(ord, dt, nu, impl=impl)
        outs.append(p.hd_get_state())
    scale = max(np.abs(q).max() for q in outs[1][:3])
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert np.isfinite(a).all()
        assert np.abs(a - b).max() / scale < TOL_FIELD
    assert rel(outs[0][3][:, :, :nph], outs[1][3][:, :, :nph]) < 100 * TOL_FIELD
    v = [p.hd_field(i) for i in range(3)]
    div, vt0, vtL, vn0, vnL = p.vdiagnostic(*v)
    e1 = p.energy(*v, 1)
    print("full-size properties:", shape, "div", div, "vt", vt0, vtL, "vn", vn0, vnL, "energy", e1)
    assert e1 > 0 and div < 1e-5 * e1, (div, e1)      # limited by the FC(5) accuracy of the z derivative on coarse grids
    assert vn0 < 1e-24 * e1 and vnL < 1e-24 * e1, (vn0, vnL, e1)
    assert vt0 < 1e-3 * e1 and vtL < 1e-3 * e1, (vt0, vtL, e1)          # slip error of the p' prediction, O(dt^2)
    p.close()


# ---- BOOTS regridder (tools/boots.fpp) ----------------------------------------------------------
def boots_field(nz, ny, nx):
    z = np.linspace(0.0, 1.0, nz)[:, None, None]
    y = (2 * np.pi * np.arange(ny) / ny)[None, :, None]
    x = (2 * np.pi * np.arange(nx) / nx)[None, None, :]
    return np.exp(0.4 * z) * np.sin(2 * x) * np.cos(3 * y) + np.sin(3 * z) * np.cos(x)


def boots_golden_inputs():
    """The two seeded old-grid fields of tests/golden/boots_27_46.npz: a smooth wall-bounded part plus white noise."""
    out = []
    for seed, (nzt, nyt, nxt) in ((21, (27, 16, 16)), (22, (21, 32, 16))):
        out.append(boots_field(nzt, nyt, nxt) + 1e-3 * np.random.default_rng(seed).standard_normal((nzt, nyt, nxt)))
    return out


def case_boots_golden(lib, tables, golden_path):
    """The committed fixture (written by tests/golden/make_golden.py from the oracle): lib = None checks the oracle."""
    gold = np.load(golden_path)
    a, b = boots_golden_inputs()
    for vt, key, (nx, ny, nzp, ozt) in ((a, "a", (32, 32, 46, 5)), (b, "b", (32, 32, 41, 0))):
        got = O.boots_regrid(vt, nx, ny, nzp, ozt, tables) if lib is None else api.boots_regrid(vt, nx, ny, nzp, ozt, tables, lib=lib)
        assert rel(got[::3, ::4, ::4], gold[key]) < TOL_FIELD


def case_boots(lib, tables, cases, tmpdir=None):
    """cases: (nxt, nyt, nzt, ozt, nx, ny, nzp).  The regridded field against the oracle's restatement of the
    BOOTS3D loop (FFT-based, as the reference does it) on a random and on a smooth wall-bounded field; the file
    loop with the reference's names."""
    rng = np.random.default_rng(11)
    n0 = api.boots_launch_count(lib)
    for nxt, nyt, nzt, ozt, nx, ny, nzp in cases:
        assert api.boots_points(nzt, nzp, lib) == O.boots_points(nzt, nzp)
        for vt in (rng.standard_normal((nzt, nyt, nxt)), boots_field(nzt, nyt, nxt)):
            got = api.boots_regrid(vt, nx, ny, nzp, ozt, tables, lib=lib)
            ref = O.boots_regrid(vt, nx, ny, nzp, ozt, tables)
            assert got.shape == ref.shape
            # the error is measured against the largest number the two computations add up: the continuation of
            # white noise is up to |dir| ~ 1e4 (A25-5) .. 8e4 (A50-5) times the field, of a smooth field O(1)
            scale = np.abs(ref).max()
            Czt = O.boots_points(nzt, nzp)[0]
            if Czt > 0:
                a = np.zeros((nxt, nyt, nzt + Czt))
                a[:, :, :nzt] = vt.transpose(2, 1, 0)
                O.fc_continue_z(O.Grid(nxt, nyt, nzt + Czt, Czt, ozt, tdir=tables), a)
                scale = max(scale, np.abs(a).max())
            err = np.abs(got - ref).max() / scale
            assert err < TOL_FIELD, (nxt, nyt, nzt, nx, ny, nzp, err)
    assert api.boots_launch_count(lib) > n0
    if tmpdir is not None:
        nxt, nyt, nzt, ozt, nx, ny, nzp = cases[0]
        idir, odir, rdir = tmpdir / "old", tmpdir / "new", tmpdir / "new_oracle"
        for d in (idir, odir, rdir):
            d.mkdir()
        for name in ("vx.0001.out", "th.0001.out"):
            rng.standard_normal((nzt, nyt, nxt)).tofile(str(idir / name))
        fnlist = "vx.0001.out; th.0001.out"
        api.boots_files(idir, odir, tables, fnlist, nxt, nyt, nzt, ozt, nx, ny, nzp, lib=lib)
        for f in O.boots_files(idir, rdir, tables, fnlist, nxt, nyt, nzt, ozt, nx, ny, nzp):
            import os
            mine = np.fromfile(str(odir / os.path.basename(f)))
            assert mine.size == nx * ny * nzp and rel(mine, np.fromfile(f)) < 10 * TOL_FIELD   # white noise, see above
    # the reference's argument checks (boots.fpp:146-160; fcgram_mod.f90:128-141)
    import pytest
    small = np.zeros((20, 16, 16))
    with pytest.raises(api.SpecterError, match="prolongation"):
        api.boots_regrid(small, 16, 16, 19, 0, tables, lib=lib)
    with pytest.raises(api.SpecterError, match="Mismatch"):
        api.boots_regrid(small, 16, 16, 20, 5, tables, lib=lib)
    with pytest.raises(api.SpecterError, match="table"):
        api.boots_regrid(np.zeros((27, 16, 16)), 16, 16, 46, 4, tables, lib=lib)


# ---- the global-quantity text files (include/<solver>/<solver>_global.f90) -------------------------------------------
GLOBAL_WIDTHS = {"noslip_diagnostic.txt": [13] * 6, "conducting_diagnostic.txt": [13] * 7, "vacuum_diagnostic.txt": [13] * 7,
                 "scalar_constant_diagnostic.txt": [13] * 3, "scalar.txt": [13, 22, 22, 23], "energy.txt": [13, 23, 23],
                 "cross.txt": [13, 23, 24]}


def _fortran_float(field):
    """A `1P Ew.d` field back to a number; three-digit exponents are printed without the E (1.000000-120)."""
    import re
    t = field.strip()
    if "E" not in t:
        t = re.sub(r"(?<=\d)([+-]\d{3})$", r"E\1", t)
    return float(t)


def _split_fixed(line, widths):
    assert len(line) == sum(widths), (len(line), widths)
    out, pos = [], 0
    for w in widths:
        out.append(line[pos:pos + w])
        pos += w
    return out


def compare_global_dirs(ours, ref, solver, nrows, eng):
    """The same files with the same fixed-width layout (the reference's FORMATs) and the same numbers: 16 / 14 printed
    digits to 1e-9, six printed digits to 2e-6, residual columns on the scale of the energy."""
    import os
    from pathlib import Path
    ours, ref = Path(str(ours)), Path(str(ref))
    txt = lambda d: sorted(n for n in os.listdir(d) if n.endswith(".txt") and n != "benchmark.txt")
    assert txt(ours) == txt(ref), (solver, txt(ours), txt(ref))
    mag = solver in ("MHD", "MHDBOUSS")
    for name in txt(ref):
        widths = GLOBAL_WIDTHS.get(name)
        if widths is None:       # balance.txt / helicity.txt: the widths depend on the solver family
            widths = ({"balance.txt": [13, 23, 23, 24], "helicity.txt": [13, 24]} if not mag else
                      {"balance.txt": [13, 23, 23, 23], "helicity.txt": [13, 24, 24]})[name]
        la, lb = open(ours / name).read().split("\n"), open(ref / name).read().split("\n")
        assert len(la) == len(lb) == nrows + 1 and la[-1] == "", (name, len(la), len(lb))
        for a, b in zip(la[:-1], lb[:-1]):
            fa, fb = _split_fixed(a, widths), _split_fixed(b, widths)
            assert fa[0] == fb[0]                                   # the time label, character for character
            for w, x, y in zip(widths[1:], fa[1:], fb[1:]):
                xv, yv = _fortran_float(x), _fortran_float(y)
                tol = TOL_DIAG if w > 13 else 2e-6                  # 16 / 14 digits printed, or 6
                assert abs(xv - yv) <= tol * abs(yv) + TOL_DIAG * eng, (solver, name, x, y)


def case_global_files(lib, tables, shape, tmpdir, dt=1e-3):
    """sx_global for HD, BOUSS and MHDBOUSS (conducting bottom / vacuum top, so that both magnetic files appear): the
    same files, the same fixed-width layout (the reference's FORMATs) and the same numbers as the oracle writes."""
    import os
    runs = []
    for solver, bc in (("HD", None), ("BOUSS", None), ("MHDBOUSS", (0, 1))):
        g, p = make(lib, tables, *shape)
        g.load_neumann()
        if solver == "HD":
            s = O.make_hd_state(g)
            p.hd_put_state(s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)
        elif solver == "BOUSS":
            s = O.make_bouss_state(g)
            s.fs = rand_spec(g, 5) * 1e-3        # a scalar source so that the injection column is not identically zero
            p.bouss_put_state(s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)
        else:
            s = O.make_mhdbouss_state(g)
            p.setup_bc("b", PERIODIC4 + [B_KIND[bc[0]], B_KIND[bc[1]]])
            p.mhdbouss_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.th, s.fx, s.fy, s.fz, s.mx, s.my, s.mz, s.fs)
        ours, ref = tmpdir / ("ours_" + solver), tmpdir / ("ref_" + solver)
        ours.mkdir(), ref.mkdir()
        steps = (1, 11) if solver == "HD" else (11,)      # two calls: the rows are appended
        for t in steps:
            p.global_quantities(solver, ours, t, dt)
            O.solver_global(g, s, solver, ref, t, dt, *(bc or (0, 0)))
        eng = O.energy(g, s.vx, s.vy, s.vz, 1)
        compare_global_dirs(ours, ref, solver, len(steps), eng)
        runs.append(solver)
        p.close()
    return runs


# ---- the reference's other FC-Gram tables (tables/README.info: O = 3..9, C = 15..34) ------------------------------
_SOLVER_FIELDS = {"hd": ("vx", "vy", "vz"), "bouss": ("vx", "vy", "vz", "th"), "mhd": ("vx", "vy", "vz", "ax", "ay", "az")}


def _oracle_substeps(g, solver, ord, dt, eps=0.0, seed=0):
    """The oracle's per-substep fields of one RK step; eps > 0 perturbs the initial fields by eps * N(0,1) relative."""
    s = {"hd": O.make_hd_state, "bouss": O.make_bouss_state, "mhd": O.make_mhd_state}[solver](g)
    names = _SOLVER_FIELDS[solver]
    if eps:
        rng = np.random.default_rng(seed)
        for n in names:
            q = getattr(s, n)
            q *= 1.0 + eps * rng.standard_normal(q.shape)
    init = {n: getattr(s, n).copy() for n in names}
    base = [init[n].copy() for n in names]
    out = []
    for o in range(ord, 0, -1):
        if solver == "hd":
            O.hd_rkstep2(g, s, *base, o, dt, 1e-3, (0., 0.), (0., 0.))
        elif solver == "bouss":
            O.bouss_rkstep2(g, s, *base, o, dt, 1e-3, 1e-3, 1.0, 1.0)
        else:
            O.mhd_rkstep2(g, s, *base, o, dt, 1e-3, 5e-3, (0., 0., 0.))
        out.append({n: getattr(s, n).copy() for n in names})
    return s, init, out


def case_substeps_other_table(lib, tables, shape, Cz, oz, solver, ord=2, dt=1e-3, impl=0, draws=3):
    """One RK step of a solver on another continuation table of the reference, per substep against the oracle.

    The continuation rows are sum_j dir(i,j) f(j): rounding differences of the last bit are amplified by max|dir|
    (29 for A15-3, 3.5e3 for A25-5, 2.6e6 for A34-8, 1.1e7 for A33-9) and carried into every spectral coefficient by
    the transform, in the reference as much as here.  Each field is therefore held to the north-star tolerance 1e-11
    OR to 30 x the oracle's own response to a 1e-16 relative perturbation of its inputs (the largest of `draws` draws),
    whichever is larger: the result of the reference's arithmetic is not defined more sharply than that."""
    with fc_table(Cz, oz, tables):
        g, p = make(lib, tables, *shape, ord=ord)
        s, init, ref = _oracle_substeps(g, solver, ord, dt)
        names = _SOLVER_FIELDS[solver]
        groups = [names[:3]] + ([names[3:]] if len(names) > 3 else [])
        sens = [{n: 0.0 for n in names} for _ in ref]
        for seed in range(1, draws + 1):
            _, _, pert = _oracle_substeps(g, solver, ord, dt, eps=1e-16, seed=seed)
            for k, (a, b) in enumerate(zip(ref, pert)):
                for grp in groups:
                    scale = max(float(np.abs(a[n]).max()) for n in grp)
                    for n in grp:
                        sens[k][n] = max(sens[k][n], float(np.abs(a[n] - b[n]).max()) / scale)
        if solver == "hd":
            p.hd_put_state(init["vx"], init["vy"], init["vz"], s.pr * 0, s.fx, s.fy, s.fz)
            p.hd_rkstep1()
            step = lambda o: p.hd_rkstep2(o, dt, 1e-3, (0., 0.), (0., 0.), impl)
            get = lambda: dict(zip(names, p.hd_get_state()[:3]))
        elif solver == "bouss":
            p.bouss_put_state(init["vx"], init["vy"], init["vz"], s.pr * 0, init["th"], s.fx, s.fy, s.fz, s.fs)
            p.bouss_rkstep1()
            step = lambda o: p.bouss_rkstep2(o, dt, 1e-3, 1e-3, 1.0, 1.0, impl=impl)
            get = lambda: (lambda st: dict(zip(names, list(st[:3]) + [st[4]])))(p.bouss_get_state())
        else:
            p.mhd_put_state(init["vx"], init["vy"], init["vz"], s.pr * 0, init["ax"], init["ay"], init["az"],
                            s.fx, s.fy, s.fz, s.mx, s.my, s.mz)
            p.mhd_rkstep1()
            step = lambda o: p.mhd_rkstep2(o, dt, 1e-3, 5e-3, (0., 0., 0.), impl)
            get = lambda: (lambda st: dict(zip(names, list(st[:3]) + list(st[4:7]))))(p.mhd_get_state())
        worst = {}
        for k, o in enumerate(range(ord, 0, -1)):
            step(o)
            got = get()
            for grp in groups:
                scale = max(float(np.abs(ref[k][n]).max()) for n in grp)
                for n in grp:
                    err = float(np.abs(got[n] - ref[k][n]).max()) / scale
                    tol = max(TOL_FIELD, 30.0 * sens[k][n])
                    assert err <= tol, (solver, Cz, oz, o, n, err, sens[k][n])
                    worst[(o, n)] = (err, sens[k][n])
        p.close()
        return worst


def case_operators_other_table(lib, tables, shape, Cz, oz):
    """The transform / boundary operator cases on another table, tolerances scaled by max|dir| / max|dir(A25-5)|."""
    with fc_table(Cz, oz, tables):
        case_fft1d_z(lib, tables, shape)
        case_fft3d(lib, tables, shape)
        case_goto_domain(lib, tables, shape)
        case_projection(lib, tables, shape)
        case_wall_reconstructions(lib, tables, shape)
        case_scalar_vecpot_bc(lib, tables, shape)


# ---- diagnostics over many steps, BOUSS and MHD (bouss_global.f90, mhd_global.f90) --------------------------------
def _solver_diag_rows(solver, step, sample, nsteps, every):
    rows = []
    for t in range(nsteps):
        step()
        if (t + 1) % every == 0 or t == nsteps - 1:
            rows.append((float(t + 1),) + tuple(float(x) for x in sample()))
    return np.array(rows)


# column kinds: "b" bulk positive quantity (relative 1e-9), "s" signed quantity that may pass through zero and "r"
# residual at the FC-accuracy floor -- both held to 1e-9 of the energy column they belong to
_DIAG_COLS = {
    "bouss": ("eng", "ens", "pot", "th2", "gradth2", "thfs", "div", "th_wall0", "th_wallL"),
    "mhd": ("eng", "ens", "cur", "engk", "engm", "helk", "helm", "crh", "asq", "div",
            "diva", "divb", "jt0", "jtL", "bn0", "bnL"),
}
_DIAG_KIND = {
    "bouss": "bbsbbsrrr",
    "mhd": "bbbbbsssb" + "r" + "rrrrrr",
}


def oracle_solver_diagnostics(g, solver, nsteps=100, every=10, dt=1e-3):
    """The oracle's rows (step, columns of _DIAG_COLS[solver]); also what tests/golden/make_golden.py commits."""
    if solver == "bouss":
        s = O.make_bouss_state(g)
        ostep = lambda: O.bouss_step(g, s, dt, 1e-3, 1e-3)
        osample = lambda: (O.hdcheck(g, s.vx, s.vy, s.vz, s.fx, s.fy, s.fz) + O.pscheck(g, s.th, s.fs)
                           + (O.divergence(g, s.vx, s.vy, s.vz),) + tuple(O.sdiagnostic(g, s.th)))
    else:
        s = O.make_mhd_state(g)
        g.load_neumann()
        ostep = lambda: O.mhd_step(g, s, dt, 1e-3, 5e-3)
        osample = lambda: (O.mhdcheck(g, s.vx, s.vy, s.vz, s.ax, s.ay, s.az) + (O.divergence(g, s.vx, s.vy, s.vz),)
                           + tuple(O.bdiagnostic(g, s.ax, s.ay, s.az)["conducting"]))
    return _solver_diag_rows(solver, ostep, osample, nsteps, every)


def case_solver_diagnostics(lib, tables, shape, solver, nsteps=100, every=10, impl=0, dt=1e-3, golden=None):
    """The global quantities of bouss_global.f90 / mhd_global.f90 (hdcheck or mhdcheck, pscheck, vdiagnostic,
    sdiagnostic, bdiagnostic for conducting walls) every `every` steps over `nsteps` RK2 steps: bulk quantities within
    1e-9 relative of the oracle's (north star: "diagnostics over 100 steps to ~1e-9"), signed and residual columns within
    1e-9 of the energy scale."""
    g, p = make(lib, tables, *shape, ord=2)
    if solver == "bouss":
        s = O.make_bouss_state(g)
        p.bouss_put_state(s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)
        v = [p.bouss_field(i) for i in range(3)]; f = [p.bouss_field(4 + i) for i in range(3)]
        th, fs = p.bouss_field(10), p.bouss_field(11)
        step = lambda: p.bouss_step(dt, 1e-3, 1e-3, impl=impl)
        sample = lambda: p.hdcheck(*v, *f) + p.pscheck(th, fs) + p.vdiagnostic(*v)[:1] + p.sdiagnostic(th)
    else:
        s = O.make_mhd_state(g)
        p.mhd_put_state(s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.fx, s.fy, s.fz, s.mx, s.my, s.mz)
        v = [p.mhd_field(i) for i in range(3)]; a = [p.mhd_field(10 + i) for i in range(3)]
        step = lambda: p.mhd_step(dt, 1e-3, 5e-3, impl=impl)
        sample = lambda: p.mhdcheck(*v, *a) + p.vdiagnostic(*v)[:1] + p.bdiagnostic(*a)["conducting"]
    rows = _solver_diag_rows(solver, step, sample, nsteps, every)
    ref = np.array(golden) if golden is not None else oracle_solver_diagnostics(g, solver, nsteps, every, dt)
    assert rows.shape == ref.shape == (ref.shape[0], 1 + len(_DIAG_COLS[solver])), (rows.shape, ref.shape)
    worst = {}
    escale = np.abs(ref[:, 1]).max()
    for c, (name, kind) in enumerate(zip(_DIAG_COLS[solver], _DIAG_KIND[solver]), start=1):
        if kind == "b":
            err = float(np.abs(rows[:, c] / ref[:, c] - 1).max())
        else:
            err = float(np.abs(rows[:, c] - ref[:, c]).max() / escale)
        worst[name] = err
        assert err <= TOL_DIAG, (solver, name, err)
    p.close()
    return rows, ref, worst
