#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python layer.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py small      # seconds
    python tests/golden/make_golden.py counts     # converts the shipped counts.* files to integers
    python tests/golden/make_golden.py box360     # ~20 min, ~18 GB: Pk+Bk on dat/test_box.hdf5
    python tests/golden/make_golden.py kmupy      # seconds: _Pk_periodic_rsd(code="python"), pure numpy in the reference
    python tests/golden/make_golden.py survey     # seconds: FFT_survey_mono / _B0_survey (B0_survey body) + util.py

How: the reference package is copied to a scratch dir (its counts cold path writes into
its own dat/ directory, pyspectrum.py:1026-1028), one token is fixed (pyspectrum.py:713,
`delta[inkbin]` -> `delta_fft[inkbin]`, without which Pk_periodic raises IndexError on any
numpy >= 1.13) and it is imported with three modules shimmed in sys.modules:

    pyfftw            -> scipy.fft (pocketfft) keeping complex64 / complex128
    estimator         -> oracle/estimator_oracle.c through ctypes (no gfortran here)
    astropy.cosmology -> FlatLambdaCDM restated in oracle/pyspec_oracle.py (flat matter+Lambda, scipy quad); astropy is
                         un-pinned third-party code that is not installed here -- used by the survey path only

So the goldens pin the *Python layer* of the reference exactly (binning, shells, triangle
order, units, shot noise), on top of the C restatement of estimator.f.
"""
import os
import shutil
import sys
import types

import numpy as np
import scipy.fft as sfft

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'
SCRATCH = '/tmp/pyspectrum_ref_copy'


def import_reference(estimator_module=None):
    """estimator_module: what `import estimator` resolves to inside the reference (default: the oracle's C restatement;
    tests/test_gpu_boundary.py passes pyspectrum_b200.estimator, the CUDA drop-in)."""
    from oracle import pyspec_oracle as O
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(os.path.join(REF, 'pyspectrum'), os.path.join(SCRATCH, 'pyspectrum'))
    src = os.path.join(SCRATCH, 'pyspectrum', 'pyspectrum.py')
    txt = open(src).read()
    bad = "p0k[i-1] = np.sum(np.absolute(delta[inkbin])**2)/float(Nk)/kf**3\n            counts[i-1]"
    assert txt.count(bad) == 1
    txt = txt.replace(bad, bad.replace('delta[inkbin]', 'delta_fft[inkbin]'))
    open(src, 'w').write(txt)

    # ---- pyfftw shim -------------------------------------------------------------------
    pyfftw = types.ModuleType('pyfftw')

    def n_byte_align_empty(shape, n, dtype='complex64'):
        return np.empty(shape, dtype=dtype)
    pyfftw.n_byte_align_empty = n_byte_align_empty

    class _Plan(object):
        def __init__(self, a, inverse):
            self.a, self.inverse = a, inverse

        def __call__(self, x=None, normalise_idft=True):
            a = self.a if x is None else x
            if self.inverse:
                return sfft.ifftn(a, norm='backward' if normalise_idft else 'forward')
            return sfft.fftn(a)
    builders = types.ModuleType('pyfftw.builders')
    builders.fftn = lambda a, **kw: _Plan(a, False)
    builders.ifftn = lambda a, **kw: _Plan(a, True)
    interfaces = types.ModuleType('pyfftw.interfaces')
    cache = types.ModuleType('pyfftw.interfaces.cache')
    cache.enable = lambda: None
    interfaces.cache = cache
    pyfftw.builders, pyfftw.interfaces = builders, interfaces
    sys.modules.update({'pyfftw': pyfftw, 'pyfftw.builders': builders,
                        'pyfftw.interfaces': interfaces, 'pyfftw.interfaces.cache': cache})

    # ---- estimator shim (f2py signatures, SURVEY 8b level 2) -----------------------------
    est = types.ModuleType('estimator')
    est.assign_quad = lambda r, w, dtl, kf_ks, offset, ia, ib, ic, id, np_=None, ngrid=None: \
        O.assign_quad(r, w, dtl, kf_ks, offset, ia, ib, ic, id)
    est.fcomb_periodic = lambda dcl, n, ngrid=None: O.fcomb_periodic(dcl, n)
    est.fcomb_survey = lambda dcl, ngrid=None: O.fcomb_survey(dcl)
    est.pk_pbox_rsd = lambda dtl, irsd, lbox, nbin, nmu, ngrid=None: O.pk_pbox_rsd(dtl, irsd, int(lbox), nbin, nmu)
    sys.modules['estimator'] = est if estimator_module is None else estimator_module

    # ---- astropy stub ----------------------------------------------------------------------
    astropy = types.ModuleType('astropy')
    cosmo = types.ModuleType('astropy.cosmology')
    cosmo.FlatLambdaCDM = O.FlatLambdaCDM
    astropy.cosmology = cosmo
    sys.modules.update({'astropy': astropy, 'astropy.cosmology': cosmo})

    sys.path.insert(0, SCRATCH)
    for m in [m for m in sys.modules if m == 'pyspectrum' or m.startswith('pyspectrum.')]:
        del sys.modules[m]                                   # a fresh import binds the estimator module given above
    from pyspectrum import pyspectrum as pySpec
    assert pySpec.__file__.startswith(SCRATCH)
    return pySpec


def clustered_catalogue(seed, Np, Lbox):
    """Deterministic clustered toy catalogue: Gaussian blobs around uniform parents + uniform background."""
    rng = np.random.default_rng(seed)
    npar = max(Np // 40, 1)
    parents = rng.uniform(0, Lbox, (3, npar))
    which = rng.integers(0, npar, Np // 2)
    kids = parents[:, which] + rng.normal(0, 0.03 * Lbox, (3, Np // 2))
    bg = rng.uniform(0, Lbox, (3, Np - Np // 2))
    xyz = np.concatenate([kids, bg], axis=1) % Lbox
    return np.ascontiguousarray(xyz)


def small(pySpec):
    out = {}
    # --- case A: Ngrid=32, clustered, unweighted -------------------------------------------
    for tag, N, L, Np, seed, weighted in [('A', 32, 200., 4000, 11, False),
                                         ('B', 24, 100., 1500, 12, True),
                                         ('C', 36, 500., 6000, 13, False)]:     # 36 = 2^2 3^2 (non power of two)
        xyz = clustered_catalogue(seed, Np, L)
        # a few particles outside the box on both sides: exercises the clip (py:939-941)
        xyz[:, :3] = np.array([[-1.0, L + 3.0, L], [0.0, L * (1 - 1e-7), -0.5], [L * 0.5, L + 1e-3, 0.0]])
        w = np.random.default_rng(seed + 100).uniform(0.5, 2.0, Np) if weighted else None
        pk = pySpec.Pk_periodic(xyz, w=w, Lbox=L, Ngrid=N)
        d = {'xyz': xyz, 'Lbox': L, 'Ngrid': N}
        if w is not None:
            d['w'] = w
        for key in ['k', 'p0k', 'counts']:
            d['pk_' + key] = pk[key]
        d['pk_p0k_sn'] = pk['p0k_sn']
        for rsd in (0, 1, 2):
            for nmu in (5, 10):
                # Lbox is truncated to an integer inside pk_pbox_rsd (estimator.f:158)
                pr = pySpec.Pk_periodic_rsd(xyz, w=w, Lbox=L, Ngrid=N, rsd=rsd, Nmubin=nmu)
                for key in ['k', 'p0k', 'p2k', 'p4k', 'p_sn', 'counts', 'k_kmu', 'mu_kmu', 'p_kmu', 'counts_kmu']:
                    d['rsd%d_mu%d_%s' % (rsd, nmu, key)] = np.asarray(pr[key])
        # delta(k) half field itself
        delta = pySpec.FFT_periodic(xyz, w=w, Lbox=L, Ngrid=N)
        d['delta_half'] = np.ascontiguousarray(delta)          # (N/2+1,N,N) complex64, [kx,ky,kz]
        # bispectrum: Bk_periodic asserts Ngrid==360 (py:332), so replay its body around _Bk_periodic
        for (step, Ncut, Nmax) in [(3, 3, 4), (2, 3, 6), (1, 1, 8)]:
            if step * (Nmax + 0.5) > N:      # shells beyond the grid corner are empty anyway
                continue
            delta_fft = pySpec.reflect_delta(delta, Ngrid=N)
            bk = pySpec._Bk_periodic(delta_fft, step=step, Ncut=Ncut, Nmax=Nmax)
            ww = np.ones(Np) if w is None else w
            nbar = np.sum(ww) / L ** 3
            kf = 2 * np.pi / L
            bk['p0k1'] = bk['p0k1'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
            bk['p0k2'] = bk['p0k2'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
            bk['p0k3'] = bk['p0k3'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
            b_sn = (bk['p0k1'] + bk['p0k2'] + bk['p0k3']) / nbar + 1. / nbar ** 2
            bk['b123'] = bk['b123'] * (2 * np.pi) ** 6 / kf ** 6 - b_sn
            bk['b123_sn'] = b_sn
            bk['q123'] = bk['b123'] / (bk['p0k1'] * bk['p0k2'] + bk['p0k1'] * bk['p0k3'] + bk['p0k2'] * bk['p0k3'])
            pre = 'bk_s%d_c%d_m%d_' % (step, Ncut, Nmax)
            for key in ['i_k1', 'i_k2', 'i_k3', 'p0k1', 'p0k2', 'p0k3', 'b123', 'q123', 'counts', 'b123_sn']:
                d[pre + key] = np.asarray(bk[key])
            raw = pySpec._counts_Bk123(Ngrid=N, Nmax=Nmax, Ncut=Ncut, step=step)
            d[pre + 'rawcounts'] = np.rint(raw / N ** 3).astype(np.int64)
            assert np.abs(raw / N ** 3 - d[pre + 'rawcounts']).max() < 1e-6
        np.savez_compressed(os.path.join(HERE, 'small_%s.npz' % tag), **d)
        print('wrote small_%s.npz' % tag, len(d), 'arrays')
    return out


def kmupy(pySpec):
    """code='python' branch of _Pk_periodic_rsd (py:545-626): pure numpy in the reference, so these goldens involve NO restatement
    at all beyond the stored delta(k) it is fed (small_*.npz, reflected by the reference's own reflect_delta)."""
    import contextlib
    import io
    d = {}
    for tag in 'ABC':
        g = np.load(os.path.join(HERE, 'small_%s.npz' % tag))
        N, L = int(g['Ngrid']), float(g['Lbox'])
        full = pySpec.reflect_delta(g['delta_half'], Ngrid=N)
        for rsd, nmu, Lb in [(2, 5, L), (0, 5, L), (1, 4, None), (2, 10, None)]:
            with contextlib.redirect_stdout(io.StringIO()):           # py:601 prints every (i,j)
                out = pySpec._Pk_periodic_rsd(full, Lbox=Lb, rsd=rsd, Nmubin=nmu, code='python')
            pre = '%s_rsd%d_mu%d_L%s_' % (tag, rsd, nmu, 'none' if Lb is None else 'box')
            for key, val in zip(['k', 'p0k', 'p2k', 'p4k', 'nk', 'k_kmu', 'mu_kmu', 'p_kmu', 'n_kmu'], out):
                d[pre + key] = np.asarray(val)
    np.savez_compressed(os.path.join(HERE, 'kmu_python.npz'), **d)
    print('wrote kmu_python.npz', len(d), 'arrays')


def counts():
    """Shipped caches -> exact integers: counts/N^3 rounds to an integer to <3e-14."""
    from scipy.io import FortranFile
    N = 360
    for nmax in (10, 40, 50):
        f = FortranFile(os.path.join(REF, 'pyspectrum', 'dat', 'counts.Ngrid360.Nmax%d.Ncut3.step3.pyfftw' % nmax), 'r')
        c = f.read_reals().reshape(nmax, nmax, nmax)
        f.close()
        ci = np.rint(c / N ** 3).astype(np.int64)
        rel = np.abs(c - ci.astype(float) * N ** 3)[c > 0] / c[c > 0]
        print('Nmax', nmax, 'nonzero', np.count_nonzero(ci), 'max rel dev from integer', rel.max())
        idx = np.argwhere(ci > 0)
        np.savez_compressed(os.path.join(HERE, 'counts_N360_Nmax%d_Ncut3_step3.npz' % nmax),
                            ijl=(idx + 1).astype(np.int16), n=ci[ci > 0], raw=c[c > 0])


def read_test_box():
    raw = open(os.path.join(REF, 'pyspectrum', 'dat', 'test_box.hdf5'), 'rb').read()
    xyz = np.frombuffer(raw[2048:2048 + 1762608], '<f8').reshape(73442, 3)
    vxyz = np.frombuffer(raw[1764656:1764656 + 1762608], '<f8').reshape(73442, 3)
    return xyz, vxyz


def box360(pySpec):
    xyz, _ = read_test_box()
    xyz = np.ascontiguousarray(xyz.T)
    d = {'xyz': xyz.astype(np.float64)}
    pk = pySpec.Pk_periodic(xyz, Lbox=2600., Ngrid=360)
    for key in ['k', 'p0k', 'counts']:
        d['pk_' + key] = pk[key]
    d['pk_p0k_sn'] = pk['p0k_sn']
    print('Pk done', pk['p0k'][:4])
    pr = pySpec.Pk_periodic_rsd(xyz, Lbox=2600., Ngrid=360)
    for key in ['k', 'p0k', 'p2k', 'p4k', 'counts', 'k_kmu', 'mu_kmu', 'p_kmu', 'counts_kmu']:
        d['rsd2_mu10_' + key] = np.asarray(pr[key])
    print('Pk rsd done')
    bk = pySpec.Bk_periodic(xyz, Lbox=2600., Ngrid=360, step=3, Ncut=3, Nmax=40)
    for key in ['i_k1', 'i_k2', 'i_k3', 'p0k1', 'p0k2', 'p0k3', 'b123', 'q123', 'counts', 'b123_sn']:
        d['bk_' + key] = np.asarray(bk[key])
    np.savez_compressed(os.path.join(HERE, 'box360.npz'), **d)
    print('wrote box360.npz; Ntri =', len(bk['b123']))


def survey_catalogue(seed, Nd, Nr):
    """Deterministic toy survey: a cone in (RA, Dec, z) with clustered data (blobs) and uniform-in-volume randoms, a
    redshift-dependent nbar(z) and (for the data) systematic weights."""
    rng = np.random.default_rng(seed)

    def cone(n):
        ra = rng.uniform(110., 250., n)
        dec = np.degrees(np.arcsin(rng.uniform(np.sin(np.radians(-5.)), np.sin(np.radians(60.)), n)))
        z = (rng.uniform(0.2 ** 3, 0.55 ** 3, n)) ** (1. / 3.)
        return np.array([ra, dec, z])
    rand = cone(Nr)
    par = cone(max(Nd // 30, 1))
    kids = par[:, rng.integers(0, par.shape[1], Nd // 2)] + rng.normal(0, 1., (3, Nd // 2)) * np.array([[2.], [2.], [0.01]])
    kids[2] = np.clip(kids[2], 0.2, 0.55)
    kids[1] = np.clip(kids[1], -5., 60.)
    data = np.concatenate([kids, cone(Nd - Nd // 2)], axis=1)
    nz = lambda z: 3e-4 * np.exp(-((z - 0.35) / 0.2) ** 2)
    w = rng.uniform(0.8, 1.3, Nd)
    return data, nz(data[2]), w, rand, nz(rand[2])


def survey(pySpec):
    from pyspectrum import util as UT
    from oracle import pyspec_oracle as O
    for tag, N, L, Nd, Nr, seed, weighted, cfgs in [('A', 32, 3200., 3000, 12000, 21, False, [(3, 3, 4), (2, 3, 6)]),
                                                     ('B', 36, 3000., 2500, 9000, 22, True, [(2, 2, 7), (1, 1, 8)])]:
        radecz, nb, w, radecz_r, nb_r = survey_catalogue(seed, Nd, Nr)
        d = {'radecz': radecz, 'nbar': nb, 'radecz_r': radecz_r, 'nbar_r': nb_r, 'Lbox': L, 'Ngrid': N, 'P0_fkp': 1e4}
        if weighted:
            d['w'] = w
        P0 = 1e4
        # the reference mutates radecz (degrees -> radians) and w (FKP) in place: always hand it copies
        xyz = UT.radecz_to_cartesian(radecz.copy(), cosmo=O.FlatLambdaCDM(H0=67.6, Om0=0.31))
        d['xyz'] = np.asarray(xyz)
        wd = w.copy() if weighted else None
        delta_d, Ngtot, I12d, I13d, I22d, I23d, I33d = pySpec.FFT_survey_mono(radecz.copy(), nb, w=wd, P0_fkp=P0, Lbox=L, Ngrid=N)
        delta_r, Nrtot, I12r, I13r, I22r, I23r, I33r = pySpec.FFT_survey_mono(radecz_r.copy(), nb_r, w=None, P0_fkp=P0, Lbox=L, Ngrid=N)
        d['delta_d'] = np.ascontiguousarray(delta_d)
        d['delta_r'] = np.ascontiguousarray(delta_r)
        d['sums_d'] = np.array([Ngtot, I12d, I13d, I22d, I23d, I33d])
        d['sums_r'] = np.array([Nrtot, I12r, I13r, I22r, I23r, I33r])
        # body of B0_survey (py:101-128; its Ngrid==360 assert, py:90, keeps the function itself from running here)
        deltak_d = pySpec.reflect_delta(delta_d, Ngrid=N)
        deltak_r = pySpec.reflect_delta(delta_r, Ngrid=N)
        alpha = Ngtot / Nrtot
        deltak = deltak_d - alpha * deltak_r
        for (step, Ncut, Nmax) in cfgs:
            bk = pySpec._B0_survey(deltak, alpha, alpha * I12r, alpha * I13r, alpha * I22r, alpha * I23r, alpha * I33r,
                                   Nmax=Nmax, Ncut=Ncut, step=step)
            pre = 'b0_s%d_c%d_m%d_' % (step, Ncut, Nmax)
            for key in ['i_k1', 'i_k2', 'i_k3', 'p0k1', 'p0k2', 'p0k3', 'b123', 'q123', 'counts']:
                d[pre + key] = np.asarray(bk[key])
        np.savez_compressed(os.path.join(HERE, 'survey_%s.npz' % tag), **d)
        print('wrote survey_%s.npz' % tag, len(d), 'arrays; alpha', alpha)
    # ---- util.py -------------------------------------------------------------------------------
    u = {}
    rng = np.random.default_rng(31)
    tri = np.array([(i, j, l) for i in range(1, 9) for j in range(1, i + 1) for l in range(max(i - j, 1), j + 1)])
    perm = rng.permutation(len(tri))
    u['ijl_in'] = tri[perm]
    u['ijl_out'] = UT.ijl_order(tri[perm, 0], tri[perm, 1], tri[perm, 2], typ='GM')
    UT.FlatLambdaCDM = O.FlatLambdaCDM              # util.py uses the name without importing it (NameError otherwise)
    xyz = rng.uniform(0, 500., (3, 200))
    vxyz = rng.normal(0, 600., (3, 200))
    u['rsd_xyz'], u['rsd_vxyz'] = xyz, vxyz
    for los in 'xyz':
        u['rsd_out_' + los] = UT.applyRSD(xyz, vxyz, 0.5, h=0.7, omega0_m=0.3, LOS=los, Lbox=500.)
    np.savez_compressed(os.path.join(HERE, 'util.npz'), **u)
    print('wrote util.npz')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'small'
    if what == 'counts':
        counts()
    else:
        ps = import_reference()
        {'small': small, 'box360': box360, 'survey': survey, 'kmupy': kmupy}[what](ps)
