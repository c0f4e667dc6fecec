"""cuFFT as a correctness and speed COMPARATOR (never in the product path): the library's periodic 3-D r2c / c2r
and its z transform against torch.fft on the GPU (which runs cuFFT), at BASELINE.json's 512^3.  The timing lines are
printed (run with -s); only the agreement is asserted."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _time_ours(p, fn, reps=5):
    fn()
    p.time_begin()
    for _ in range(reps):
        fn()
    return p.time_end() / reps


def _time_torch(torch, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


@pytest.mark.parametrize("shape", [(64, 32, 128), (512, 512, 512)])
def test_periodic_transforms_against_cufft(cuda_lib, shape):
    import torch
    from specter_b200 import api
    nx, ny, nz = shape
    p = api.Plan(nx, ny, nz, 0, 0, ord=2, Lx=1.0, Ly=1.0, Lz=1.0, tdir="", lib=cuda_lib)
    rng = np.random.default_rng(5)
    r = rng.standard_normal(p.rshape)
    dr, dc, dr2 = p.real(r), p.spectral(), p.real()
    p.fftp3d_real_to_complex(dr, dc)
    ours = dc.get()                                             # (kx, ky, kz)
    R = torch.from_numpy(r).cuda()                              # (z, y, x)
    F = torch.fft.rfftn(R, dim=(0, 1, 2))                       # (kz, ky, kx), forward sign -1, unnormalised: FFTW's
    ref = F.permute(2, 1, 0).contiguous().cpu().numpy()
    scale = np.abs(ref).max()
    assert np.abs(ours - ref).max() / scale < 1e-13
    p.fftp3d_complex_to_real(dc, dr2)
    back = torch.fft.irfftn(F, s=(nz, ny, nx), dim=(0, 1, 2), norm="forward").cpu().numpy()    # unnormalised backward
    assert np.abs(dr2.get() - back).max() / np.abs(back).max() < 1e-13
    # z transform alone (fftp1d_complex_to_real_z: backward, unnormalised) against cuFFT's c2c
    p.fftp1d_complex_to_real_z(dc)
    zref = torch.fft.ifft(F, dim=0, norm="forward").permute(2, 1, 0).contiguous().cpu().numpy()
    assert np.abs(dc.get() - zref).max() / np.abs(zref).max() < 1e-13
    t_ours = _time_ours(p, lambda: p.fftp3d_real_to_complex(dr, dc))
    t_cufft = _time_torch(torch, lambda: torch.fft.rfftn(R, dim=(0, 1, 2)))
    t_ours_b = _time_ours(p, lambda: p.fftp3d_complex_to_real(dc, dr2))
    t_cufft_b = _time_torch(torch, lambda: torch.fft.irfftn(F, s=(nz, ny, nx), dim=(0, 1, 2), norm="forward"))
    gb = 2 * 8.0 * nx * ny * nz / 1e9                            # one read + one write of the field per axis pass
    print(f"\ncomparator {shape}: r2c ours {t_ours:.3f} ms (stand-alone operators, 3 passes = {3e3 * gb / t_ours:.0f} GB/s "
          f"algorithmic) vs cuFFT {t_cufft:.3f} ms; c2r ours {t_ours_b:.3f} ms vs cuFFT {t_cufft_b:.3f} ms")
    p.close()
