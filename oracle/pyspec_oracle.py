"""oracle/pyspec_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + the C file next to it) of the reference's periodic-box
estimator path, used only as the parity checker and as the timed CPU baseline:

  FFT_periodic        <- pyspectrum/pyspectrum.py:909-959
  _FFT                <- pyspectrum/pyspectrum.py:1060-1080   (pyfftw -> scipy.fft/pocketfft, complex64)
  reflect_delta       <- pyspectrum/pyspectrum.py:1134-1157
  Pk_periodic         <- pyspectrum/pyspectrum.py:644-728     (with the one-token fix at :713, see below)
  Pk_periodic_rsd     <- pyspectrum/pyspectrum.py:460-538, 628-641
  Bk_periodic         <- pyspectrum/pyspectrum.py:285-356
  _Bk_periodic        <- pyspectrum/pyspectrum.py:359-457
  _counts_Bk123       <- pyspectrum/pyspectrum.py:962-1030    (cold path, vectorised; no file I/O)
  counts_bruteforce   <- definition of the counts (closed triangles mod N), toy grids only
  B0_survey           <- pyspectrum/pyspectrum.py:13-132      (without the Ngrid==360 assert)
  _B0_survey          <- pyspectrum/pyspectrum.py:135-282
  FFT_survey_mono     <- pyspectrum/pyspectrum.py:731-826
  radecz_to_cartesian <- pyspectrum/util.py:27-51, ijl_order <- util.py:8-24, applyRSD <- util.py:54-75
  FlatLambdaCDM       <- astropy.cosmology.FlatLambdaCDM(H0, Om0) (third party, un-pinned in setup.py:64, absent here):
                         flat matter+Lambda, Tcmb0=0; comoving_distance by adaptive quadrature as astropy does

Native pieces (assign_quad / fcomb_periodic / pk_pbox_rsd) come from
oracle/estimator_oracle.c, a line-by-line C restatement of pyspectrum/estimator.f.

Deviations from the reference, all documented in DESIGN.md:
  * FFTW (pyfftw, unpinned version, not installed) -> pocketfft (scipy.fft), same precision
    (complex64; complex128 on the counts cold path, as py:1004 ends up doing).
  * pyspectrum.py:713 indexes `delta` where `delta_fft` is meant (IndexError on numpy>=1.13);
    the oracle bins |delta_fft|**2, the evident intent.
  * `Ngrid == 360` assert (py:332) is dropped; `_Bk_periodic` itself is grid-agnostic.
  * shell fields are stored as float32 (they are np.real of a complex64 FFT, py:400) and
    widened to float64 inside the triple sum, instead of a (Nmax+1)*N^3 float64 store.
  * counts are computed, never read from / written to the package `dat/` directory.

PARITY STATUS: the Python layer is pinned against the *unmodified* reference module run
under import shims (tests/golden/make_golden.py) and against the reference's shipped
triangle-count files; the Fortran layer could not be compiled here (no gfortran) and is
pinned by known-answer tests only.

Nothing in pyspectrum_b200/ imports this module.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.fft as sfft

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle.so')
_LIB = None


def build(force=False):
    """gcc the C restatement into oracle/_build/liboracle.so (same flags as oracle/Makefile)."""
    src = os.path.join(_HERE, 'estimator_oracle.c')
    if (not force) and os.path.isfile(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared',
           '-o', _SO, src, '-lm']
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_assign_quad.argtypes = [fp, fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_assign_quad.restype = None
        L.oracle_fcomb.argtypes = [fp, ctypes.c_float, ctypes.c_int, ctypes.c_int]
        L.oracle_fcomb.restype = None
        L.oracle_pk_pbox_rsd.argtypes = [fp] + [dp] * 9 + [ctypes.c_int] * 5
        L.oracle_pk_pbox_rsd.restype = None
        L.oracle_quad_fields.argtypes = [ctypes.c_int, fp, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int]
        L.oracle_quad_fields.restype = None
        L.oracle_triple_sum_f32.argtypes = [fp, fp, fp, ctypes.c_int64]
        L.oracle_triple_sum_f32.restype = ctypes.c_double
        _LIB = L
    return _LIB


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------------------
# f2py-shaped wrappers (same names / argument meaning as the `estimator` extension module)
# ----------------------------------------------------------------------------------------
def assign_quad(r, w, dtl, kf_ks, offset, ia, ib, ic, id, np_=None, ngrid=None):
    """estimator.assign_quad(r,w,dtl,kf_ks,offset,ia,ib,ic,id,[np,ngrid]); dtl is intent(inout)."""
    r = np.asfortranarray(r, dtype=np.float32)          # intent(in): f2py copies/casts silently
    w = np.ascontiguousarray(w, dtype=np.float32)
    if not (dtl.dtype == np.float32 and dtl.flags.f_contiguous):
        raise ValueError('dtl must be a float32 Fortran-contiguous array (intent inout)')
    ngrid = dtl.shape[1]
    assert dtl.shape == (2 * ngrid, ngrid, ngrid) and r.shape[0] == 3
    lib().oracle_assign_quad(_fptr(r), _fptr(w), _fptr(dtl), r.shape[1], ngrid,
                             np.float32(kf_ks), np.float32(offset), ia, ib, ic, id)


def fcomb_periodic(dcl, n, ngrid=None):
    """estimator.fcomb_periodic(dcl,n,[ngrid]); dcl complex64 F-order, in place."""
    if not (dcl.dtype == np.complex64 and dcl.flags.f_contiguous):
        raise ValueError('dcl must be complex64 Fortran-contiguous (intent inout)')
    lib().oracle_fcomb(_fptr(dcl), np.float32(n), dcl.shape[0], 1)


def fcomb_survey(dcl, ngrid=None):
    if not (dcl.dtype == np.complex64 and dcl.flags.f_contiguous):
        raise ValueError('dcl must be complex64 Fortran-contiguous (intent inout)')
    lib().oracle_fcomb(_fptr(dcl), np.float32(1.0), dcl.shape[0], 0)


def _quad(mode, ins, out, irsd=0):
    if not (out.dtype == np.complex64 and out.flags.f_contiguous):
        raise ValueError('the intent(inout) array must be complex64 Fortran-contiguous')
    N = out.shape[1]
    ins = [np.asfortranarray(a, dtype=np.complex64) for a in ins]
    assert all(a.shape == out.shape == (N // 2 + 1, N, N) for a in ins)
    ptrs = [_fptr(a) for a in ins] + [None] * (4 - len(ins))
    lib().oracle_quad_fields(mode, ptrs[0], ptrs[1], ptrs[2], ptrs[3], _fptr(out), N, int(irsd))


def fivedelta2g_1(dcgxx, dcgyy, dcgzz, ngrid=None):
    """estimator.f:514-539."""
    _quad(1, [dcgyy, dcgzz], dcgxx)


def fivedelta2g_2(dcg, dcgxx, dcgxy, dcgyz, dcgzx, ngrid=None):
    """estimator.f:541-572."""
    _quad(2, [dcg, dcgxy, dcgyz, dcgzx], dcgxx)


def build_quad(dclr1, dclr2, irsd, ngrid=None):
    """estimator.f:574-603."""
    _quad(3, [dclr1], dclr2, irsd)


def pk_pbox_rsd(dtl, irsd, lbox, nbin, nmu, ngrid=None):
    """estimator.pk_pbox_rsd -> (k,p0,p2,p4,nk,km,mk,pkm,nkm); lbox is an INTEGER dummy (f:158)."""
    dtl = np.asfortranarray(dtl, dtype=np.complex64)
    ngrid = dtl.shape[1]
    assert dtl.shape == (ngrid // 2 + 1, ngrid, ngrid)
    o1 = [np.zeros(nbin, dtype=np.float64) for _ in range(5)]
    o2 = [np.zeros((nbin, nmu), dtype=np.float64, order='F') for _ in range(4)]
    k, p0, p2, p4, nk = o1
    km, mk, pkm, nkm = o2
    lib().oracle_pk_pbox_rsd(_fptr(dtl), _dptr(k), _dptr(p0), _dptr(p2), _dptr(p4),
                             _dptr(nk), _dptr(km), _dptr(mk), _dptr(pkm), _dptr(nkm),
                             int(irsd), int(lbox), int(nbin), int(nmu), int(ngrid))
    return k, p0, p2, p4, nk, km, mk, pkm, nkm


# ----------------------------------------------------------------------------------------
# Python layer
# ----------------------------------------------------------------------------------------
def _FFT(_delta, Ngrid, workers=1):
    """py:1060-1080: de-interleave to complex64, unnormalised backward (sign +) c2c FFT, F-order copy."""
    delta = np.empty((Ngrid, Ngrid, Ngrid), dtype=np.complex64)
    delta.real = _delta[::2, :, :]
    delta.imag = _delta[1::2, :, :]
    out = sfft.ifftn(delta, norm='forward', workers=workers)      # == FFTW_BACKWARD, no 1/N^3
    assert out.dtype == np.complex64
    ifft_delta = np.zeros((Ngrid, Ngrid, Ngrid), dtype=np.complex64, order='F')
    ifft_delta[:, :, :] = out
    return ifft_delta


def assign_mesh(xyz, w=None, Lbox=2600., Ngrid=360):
    """py:931-951: clip (float64) -> float32 -> assign_quad.  Returns the (2N,N,N) float32 mesh."""
    kf_ks = np.float32(float(Ngrid) / Lbox)
    N = xyz.shape[1]
    if w is None:
        w = np.ones(N)
    xyzs = np.zeros([3, N], dtype=np.float32, order='F')
    for a in range(3):
        xyzs[a, :] = np.clip(xyz[a, :], 0., Lbox * (1. - 1e-6))
    _delta = np.zeros([2 * Ngrid, Ngrid, Ngrid], dtype=np.float32, order='F')
    assign_quad(xyzs, w, _delta, kf_ks, 0, 0, 0, 0, 0)
    return _delta


def FFT_periodic(xyz, w=None, Lbox=2600., Ngrid=360, workers=1, timings=None):
    """py:909-959.  Returns the half field, shape (Ngrid//2+1, Ngrid, Ngrid), F-order view."""
    import time
    N = xyz.shape[1]
    if w is None:
        w = np.ones(N)
    t0 = time.perf_counter()
    _delta = assign_mesh(xyz, w, Lbox, Ngrid)
    t1 = time.perf_counter()
    ifft_delta = _FFT(_delta, Ngrid, workers=workers)
    t2 = time.perf_counter()
    fcomb_periodic(ifft_delta, np.sum(w))
    t3 = time.perf_counter()
    if timings is not None:
        timings['assign'] = timings.get('assign', 0.) + (t1 - t0)
        timings['fft'] = timings.get('fft', 0.) + (t2 - t1)
        timings['fcomb'] = timings.get('fcomb', 0.) + (t3 - t2)
    return ifft_delta[:Ngrid // 2 + 1, :, :]


def reflect_delta(delt, Ngrid):
    """py:1134-1157: half field (kx in [0,N/2]) -> full Hermitian field, C-order complex64."""
    h = Ngrid // 2
    delta = np.zeros((Ngrid, Ngrid, Ngrid), dtype=np.complex64)
    delta[:h + 1] = delt
    # rows N-1 .. h+1 receive conj of rows 1 .. h-1 with (j,k) -> (-j,-k)
    src = np.conj(delt[1:h])
    delta[:h:-1, Ngrid:0:-1, Ngrid:0:-1] = src[:, 1:, 1:]
    delta[:h:-1, Ngrid:0:-1, 0] = src[:, 1:, 0]
    delta[:h:-1, 0, Ngrid:0:-1] = src[:, 0, 1:]
    delta[:h:-1, 0, 0] = src[:, 0, 0]
    for idx in [(h, 0, 0), (0, h, 0), (0, 0, h), (0, h, h), (h, 0, h), (h, h, 0), (h, h, h)]:
        delta[idx] = np.real(delt[idx])
    return delta


def _kgrid(Ngrid):
    a = np.array([min(i, Ngrid - i) for i in range(Ngrid)])
    return a


def Pk_periodic(xyz, w=None, Lbox=2600, Ngrid=360, workers=1, timings=None):
    """py:644-728 with delta -> delta_fft at :713."""
    import time
    N = xyz.shape[1]
    if w is None:
        w = np.ones(N)
    nbar = np.sum(w) / Lbox ** 3
    kf = 2 * np.pi / Lbox
    delta = FFT_periodic(xyz, w=w, Lbox=Lbox, Ngrid=Ngrid, workers=workers, timings=timings)
    t0 = time.perf_counter()
    delta_fft = reflect_delta(delta, Ngrid)
    Nbins = Ngrid // 2
    kf = 2 * np.pi / float(Lbox)
    phys_nyq = kf * float(Ngrid) / 2.
    _i = _kgrid(Ngrid)
    rk = kf * np.sqrt(_i[:, None, None] ** 2 + _i[None, :, None] ** 2 + _i[None, None, :] ** 2)
    irk = (Nbins * rk / phys_nyq + 0.5).astype(int)
    k = np.zeros(Nbins)
    p0k = np.zeros(Nbins)
    counts = np.zeros(Nbins)
    # same per-bin arithmetic as the reference's loop, but one argsort instead of N/2 mask passes
    order = np.argsort(irk.ravel(), kind='stable')
    irk_s = irk.ravel()[order]
    edges = np.searchsorted(irk_s, np.arange(1, Nbins + 2))
    rk_f = rk.ravel()
    d_f = delta_fft.ravel()
    for i in range(1, Nbins + 1):
        sel = order[edges[i - 1]:edges[i]]
        Nk = sel.size
        if Nk > 0:
            k[i - 1] = np.sum(rk_f[sel]) / float(Nk)
            p0k[i - 1] = np.sum(np.absolute(d_f[sel]) ** 2) / float(Nk) / kf ** 3
            counts[i - 1] = float(Nk)
    p0k *= (2. * np.pi) ** 3
    if timings is not None:
        timings['binning'] = timings.get('binning', 0.) + (time.perf_counter() - t0)
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'N': N, 'nbar': nbar, 'kf': kf}
    return {'meta': meta, 'k': k, 'p0k': p0k - 1. / nbar, 'counts': counts, 'p0k_sn': 1. / nbar}


def Pk_periodic_rsd(xyz, w=None, Lbox=2600, Ngrid=360, rsd=2, Nmubin=10, workers=1, timings=None):
    """py:460-538 with code='fortran' (py:628-641)."""
    import time
    N = xyz.shape[1]
    nbar = float(N) / Lbox ** 3
    kf = 2 * np.pi / Lbox
    delta = FFT_periodic(xyz, w=w, Lbox=Lbox, Ngrid=Ngrid, workers=workers, timings=timings)
    t0 = time.perf_counter()
    Nbins = Ngrid // 2
    dtl = np.zeros((Ngrid // 2 + 1, Ngrid, Ngrid), dtype=np.complex64, order='F')
    dtl[:, :, :] = delta[:, :, :]
    k, p0k, p2k, p4k, n_k, k_kmu, mu_kmu, p_kmu, n_kmu = pk_pbox_rsd(dtl, rsd, Lbox, Nbins, Nmubin)
    pk_norm = (2. * np.pi) ** 3
    p0k *= pk_norm
    p2k *= pk_norm
    p4k *= pk_norm
    p_kmu *= pk_norm
    if timings is not None:
        timings['multipoles'] = timings.get('multipoles', 0.) + (time.perf_counter() - t0)
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'N': N, 'nbar': nbar, 'kf': kf}
    return {'meta': meta, 'k': k, 'p0k': p0k - 1. / nbar, 'p2k': p2k, 'p4k': p4k,
            'p_sn': np.repeat(1. / nbar, len(k)), 'counts': n_k, 'k_kmu': k_kmu, 'mu_kmu': mu_kmu,
            'p_kmu': p_kmu - 1. / nbar, 'counts_kmu': n_kmu}


def shell_index(Ngrid, step):
    """py:373-378: irk = int(|k|/step + 0.5) on the full grid (float64)."""
    a = _kgrid(Ngrid)
    rk = np.sqrt(a[:, None, None] ** 2 + a[None, :, None] ** 2 + a[None, None, :] ** 2)
    return (rk / step + 0.5).astype(int)


def triangle_list(Nmax, Ncut, step):
    """Loop nest of py:415-417 -> int array (Ntri,3) of shell indices (i,j,l)."""
    s = Ncut // step
    out = []
    for i in range(s, Nmax + 1):
        for j in range(s, i + 1):
            for l in range(max(i - j, s), j + 1):
                out.append((i, j, l))
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def _fac(i, j, l):
    """py:418-422."""
    fac = 1.
    if (j == l) and (i == j): fac = 6.
    if (i == j) and (j != l): fac = 2.
    if (i == l) and (l != j): fac = 2.
    if (j == l) and (l != i): fac = 2.
    return fac


def _triple(a, b, c):
    return lib().oracle_triple_sum_f32(_fptr(a), _fptr(b), _fptr(c), a.size)


def shell_fields(delta_fft, irk, shells, workers=1, dtype=np.float32):
    """py:387-400: I_j(x) = Re FFT[ delta(k) 1_{irk==j} ] (forward, sign -), one array per shell."""
    out = {}
    for j in shells:
        tempK = np.zeros(delta_fft.shape, dtype=delta_fft.dtype)
        m = (irk == j)
        tempK[m] = delta_fft[m]
        out[j] = np.ascontiguousarray(np.real(sfft.fftn(tempK, workers=workers)).astype(dtype).ravel())
    return out


def _counts_Bk123(Ngrid=360, Nmax=40, Ncut=3, step=3, workers=1, triangles=None):
    """py:975-1023 cold path (delta==1; double-precision FFT as py:1004 ends up using).
    Returns the (Nmax,Nmax,Nmax) float64 array of raw sums  N^3 * (#closed triangles)."""
    irk = shell_index(Ngrid, step)
    s = Ncut // step
    ones = np.ones((Ngrid,) * 3, dtype=np.complex128)
    fields = shell_fields(ones, irk, range(s, Nmax + 1), workers=workers, dtype=np.float64)
    counts = np.zeros((Nmax, Nmax, Nmax), dtype=float)
    tri = triangle_list(Nmax, Ncut, step) if triangles is None else triangles
    for (i, j, l) in tri:
        counts[i - 1, j - 1, l - 1] = np.einsum('i,i,i', fields[i], fields[j], fields[l])
    return counts


def counts_bruteforce(Ngrid, Nmax, Ncut, step):
    """Definition: #{(q1,q2,q3) in shell_i x shell_j x shell_l : q1+q2+q3 = 0 mod N}.
    O(|S_i||S_j|) per pair -- toy grids only.  Returns integer array (Nmax,Nmax,Nmax)."""
    N = Ngrid
    irk = shell_index(N, step)
    s = Ncut // step
    coords = {j: np.argwhere(irk == j) for j in range(s, Nmax + 1)}
    out = np.zeros((Nmax, Nmax, Nmax), dtype=np.int64)
    for (i, j, l) in triangle_list(Nmax, Ncut, step):
        a, b = coords[i], coords[j]
        if a.size == 0 or b.size == 0:
            continue
        q3 = (-(a[:, None, :] + b[None, :, :])) % N
        out[i - 1, j - 1, l - 1] = np.count_nonzero(irk[q3[..., 0], q3[..., 1], q3[..., 2]] == l)
    return out


def _Bk_periodic(delta, Nmax=40, Ncut=3, step=3, workers=1, counts=None, triangles=None,
                 timings=None, pool_threads=1):
    """py:359-457.  `delta` is the full complex64 field.  `triangles` (subset, optional) bounds the
    work for the timed CPU baseline; default is the full loop nest."""
    import time
    Ngrid = delta.shape[0]
    irk = shell_index(Ngrid, step)
    Nk = np.array([np.sum(irk == i) for i in np.arange(Nmax + 1)])
    s = Ncut // step
    tri = triangle_list(Nmax, Ncut, step) if triangles is None else np.asarray(triangles).reshape(-1, 3)
    need = sorted(set(tri.ravel().tolist())) if triangles is not None else list(range(s, Nmax + 1))
    t0 = time.perf_counter()
    fields = shell_fields(delta, irk, need, workers=workers)
    p0k = np.zeros(Nmax)
    for j in need:
        f64 = fields[j].astype(np.float64)
        p0k[j - 1] = np.einsum('i,i', f64, f64) / Ngrid ** 3 / Nk[j]
    t1 = time.perf_counter()
    if counts is None:
        counts = _counts_Bk123(Ngrid=Ngrid, Nmax=Nmax, Ncut=Ncut, step=step, workers=workers, triangles=tri)
    t2 = time.perf_counter()

    def one(t):
        i, j, l = t
        return _triple(fields[i], fields[j], fields[l])
    if pool_threads > 1:
        with ThreadPoolExecutor(pool_threads) as ex:
            sums = list(ex.map(one, [tuple(t) for t in tri]))
    else:
        sums = [one(tuple(t)) for t in tri]
    t3 = time.perf_counter()
    if timings is not None:
        timings['shells'] = timings.get('shells', 0.) + (t1 - t0)
        timings['counts'] = timings.get('counts', 0.) + (t2 - t1)
        timings['triangles'] = timings.get('triangles', 0.) + (t3 - t2)

    i_arr, j_arr, l_arr = [], [], []
    p0k_i, p0k_j, p0k_l = [], [], []
    b123_arr, q123_arr, cnts_arr = [], [], []
    for (i, j, l), bisp_ijl in zip(tri, sums):
        fac = _fac(i, j, l)
        c = counts[i - 1, j - 1, l - 1]
        if c > 0:
            i_arr.append(i); j_arr.append(j); l_arr.append(l)
            p0k_i.append(p0k[i - 1]); p0k_j.append(p0k[j - 1]); p0k_l.append(p0k[l - 1])
            b123_arr.append(bisp_ijl / c)
            q123_arr.append(bisp_ijl / c / (p0k[i - 1] * p0k[j - 1] + p0k[j - 1] * p0k[l - 1] + p0k[l - 1] * p0k[i - 1]))
            cnts_arr.append(c / (fac * float(Ngrid ** 3)))
        else:       # py:438-445 (value lists grow, index lists do not: SURVEY Q9)
            p0k_i.append(0.); p0k_j.append(0.); p0k_l.append(0.)
            b123_arr.append(0.); q123_arr.append(0.); cnts_arr.append(0.)
    output = {}
    output['i_k1'] = np.array(i_arr) * step
    output['i_k2'] = np.array(j_arr) * step
    output['i_k3'] = np.array(l_arr) * step
    output['p0k1'] = np.array(p0k_i)
    output['p0k2'] = np.array(p0k_j)
    output['p0k3'] = np.array(p0k_l)
    output['b123'] = np.array(b123_arr)
    output['q123'] = np.array(q123_arr)
    output['counts'] = np.array(cnts_arr)
    return output


def Bk_periodic(xyz, w=None, Lbox=2600, Ngrid=360, step=3, Ncut=3, Nmax=40, workers=1, counts=None,
                triangles=None, timings=None, pool_threads=1):
    """py:285-356 (without the Ngrid==360 assert)."""
    N = xyz.shape[1]
    if w is None:
        w = np.ones(N)
    nbar = np.sum(w) / Lbox ** 3
    kf = 2 * np.pi / Lbox
    delta = FFT_periodic(xyz, w=w, Lbox=Lbox, Ngrid=Ngrid, workers=workers, timings=timings)
    delta_fft = reflect_delta(delta, Ngrid)
    bispec = _Bk_periodic(delta_fft, step=step, Ncut=Ncut, Nmax=Nmax, workers=workers, counts=counts,
                          triangles=triangles, timings=timings, pool_threads=pool_threads)
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'step': step, 'Ncut': Ncut, 'Nmax': Nmax, 'N': N, 'nbar': nbar, 'kf': kf}
    bispec['meta'] = meta
    bispec['p0k1'] = bispec['p0k1'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k2'] = bispec['p0k2'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k3'] = bispec['p0k3'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k_sn'] = 1. / nbar
    b_shotnoise = (bispec['p0k1'] + bispec['p0k2'] + bispec['p0k3']) / nbar + 1. / nbar ** 2
    bispec['b123'] = bispec['b123'] * (2 * np.pi) ** 6 / kf ** 6 - b_shotnoise
    bispec['b123_sn'] = b_shotnoise
    bispec['q123'] = bispec['b123'] / (bispec['p0k1'] * bispec['p0k2'] + bispec['p0k1'] * bispec['p0k3'] + bispec['p0k2'] * bispec['p0k3'])
    return bispec


# ----------------------------------------------------------------------------------------
# survey geometry and the util.py helpers either side of the path
# ----------------------------------------------------------------------------------------
class FlatLambdaCDM(object):
    """astropy.cosmology.FlatLambdaCDM(H0, Om0) restated for the two members the reference uses (py:88, 769; ut:45, 68-71):
    E(z) = sqrt(Om0 (1+z)^3 + 1 - Om0), comoving_distance = c/H0 * quad(1/E, 0, z) in Mpc."""

    def __init__(self, H0=67.6, Om0=0.31):
        self.H0, self.Om0, self.h = float(H0), float(Om0), float(H0) / 100.

    def efunc(self, z):
        return np.sqrt(self.Om0 * (1. + np.asarray(z, dtype=float)) ** 3 + (1. - self.Om0))

    def comoving_distance(self, z):
        from scipy.integrate import quad
        zz = np.atleast_1d(np.asarray(z, dtype=float))
        out = np.array([quad(lambda x: 1. / np.sqrt(self.Om0 * (1. + x) ** 3 + 1. - self.Om0), 0., zi,
                             epsabs=0., epsrel=1e-13)[0] for zi in zz])
        return 299792.458 / self.H0 * out


def radecz_to_cartesian(radecz, cosmo=None):
    """ut:27-51 (on a copy: the reference scales the caller's RA/Dec rows in place)."""
    ra, dec, z = np.array(radecz, dtype=float)
    ra *= np.pi / 180.
    dec *= np.pi / 180.
    rad = cosmo.comoving_distance(z) * cosmo.h
    return np.array([rad * np.cos(dec) * np.cos(ra), rad * np.cos(dec) * np.sin(ra), rad * np.sin(dec)])


def ijl_order(i_k, j_k, l_k, typ='GM'):
    """ut:8-24."""
    i_bq = np.arange(len(i_k))
    i_bq_new = []
    for l in np.sort(np.unique(l_k)):
        for j in np.sort(np.unique(j_k[l_k == l])):
            for i in np.sort(np.unique(i_k[(l_k == l) & (j_k == j)])):
                i_bq_new.append(i_bq[(i_k == i) & (j_k == j) & (l_k == l)])
    return np.array(i_bq_new)


def applyRSD(xyz, vxyz, redshift, h=0.7, omega0_m=0.3, LOS=None, Lbox=None):
    """ut:54-75 (with the FlatLambdaCDM the reference forgets to import)."""
    i_los = {'x': 0, 'y': 1, 'z': 2}[LOS]
    cosmo = FlatLambdaCDM(H0=100. * h, Om0=omega0_m)
    rsd_factor = (1 + redshift) / (100 * cosmo.efunc(redshift))
    xyz_rsd = xyz.copy()
    xyz_rsd[i_los] += rsd_factor * vxyz[i_los] + Lbox
    xyz_rsd[i_los] = (xyz_rsd[i_los] % Lbox)
    return xyz_rsd


def FFT_survey_mono(radecz, nb, w=None, P0_fkp=1e6, Lbox=2600., Ngrid=360, cosmo=None, workers=1):
    """py:731-826.  The caller's w is not modified (py:799 does `w *= w_fkp`)."""
    kf_ks = np.float32(float(Ngrid) / Lbox)
    N = radecz.shape[1]
    if cosmo is None:
        cosmo = FlatLambdaCDM(H0=67.6, Om0=0.31)
    w = np.ones(N) if w is None else np.array(w, dtype=float)
    xyz = radecz_to_cartesian(radecz, cosmo=cosmo)
    xyz_max, xyz_min = np.max(xyz, axis=1), np.min(xyz, axis=1)
    assert np.sum(xyz_max >= 0.5 * Lbox) + np.sum(np.abs(xyz_min) >= 0.5 * Lbox) == 0, 'box not big enough!'
    xyzs = np.zeros([3, N], dtype=np.float32, order='F')
    for a in range(3):
        xyzs[a, :] = xyz[a, :]
    Ntot = np.sum(w)
    w *= 1. / (1. + nb * P0_fkp)
    I12, I13 = np.sum(w ** 2), np.sum(w ** 3)
    I22, I23, I33 = np.sum(nb * w ** 2), np.sum(nb * w ** 3), np.sum(nb ** 2 * w ** 3)
    _delta = np.zeros([2 * Ngrid, Ngrid, Ngrid], dtype=np.float32, order='F')
    assign_quad(xyzs, w, _delta, kf_ks, 0.5 * Ngrid, 0, 0, 0, 0)
    ifft_delta = _FFT(_delta, Ngrid, workers=workers)
    fcomb_survey(ifft_delta, Ngrid)
    return ifft_delta[:Ngrid // 2 + 1, :, :], Ntot, I12, I13, I22, I23, I33


def _B0_survey(delta, alpha, I12, I13, I22, I23, I33, Nmax=40, Ncut=3, step=3, workers=1, counts=None):
    """py:135-282.  `delta` = full complex field (data - alpha * randoms)."""
    Ngrid = delta.shape[0]
    irk = shell_index(Ngrid, step)
    Nk = np.array([np.sum(irk == i) for i in np.arange(Nmax + 1)])
    s = Ncut // step
    tempdelta = delta.astype(np.complex64)                   # py:204 assigns into a complex64 tempK
    fields = shell_fields(tempdelta, irk, range(s, Nmax + 1), workers=workers)
    p0k = np.zeros(Nmax)
    for j in range(s, Nmax + 1):
        f64 = fields[j].astype(np.float64)
        p0k[j - 1] = np.einsum('i,i', f64, f64) / Ngrid ** 3 / Nk[j]
    p0k /= I22
    p0k -= (1. + alpha) * I12 / I22
    if counts is None:
        counts = _counts_Bk123(Ngrid=Ngrid, Nmax=Nmax, Ncut=Ncut, step=step, workers=workers)
    i_arr, j_arr, l_arr = [], [], []
    p0k_i, p0k_j, p0k_l = [], [], []
    b123_arr, q123_arr, cnts_arr = [], [], []
    for (i, j, l) in triangle_list(Nmax, Ncut, step):
        fac = _fac(i, j, l)
        c = counts[i - 1, j - 1, l - 1]
        if c > 0:
            i_arr.append(i); j_arr.append(j); l_arr.append(l)
            b = _triple(fields[i], fields[j], fields[l]) / c
            b -= (p0k[i - 1] + p0k[j - 1] + p0k[l - 1]) * I23 + (1. - alpha ** 2) * I13
            b /= I33
            p0k_i.append(p0k[i - 1]); p0k_j.append(p0k[j - 1]); p0k_l.append(p0k[l - 1])
            b123_arr.append(b)
            q123_arr.append(b / (p0k[i - 1] * p0k[j - 1] + p0k[j - 1] * p0k[l - 1] + p0k[l - 1] * p0k[i - 1]))
            cnts_arr.append(c / (fac * float(Ngrid ** 3)))
        else:
            p0k_i.append(0.); p0k_j.append(0.); p0k_l.append(0.)
            b123_arr.append(0.); q123_arr.append(0.); cnts_arr.append(0.)
    output = {}
    output['i_k1'] = np.array(i_arr) * step
    output['i_k2'] = np.array(j_arr) * step
    output['i_k3'] = np.array(l_arr) * step
    output['p0k1'] = np.array(p0k_i)
    output['p0k2'] = np.array(p0k_j)
    output['p0k3'] = np.array(p0k_l)
    output['b123'] = np.array(b123_arr)
    output['q123'] = np.array(q123_arr)
    output['counts'] = np.array(cnts_arr)
    return output


def B0_survey(radecz, nbar, w=None, radecz_r=None, nbar_r=None, w_r=None, P0_fkp=1e6, Lbox=2600, Ngrid=360, step=3,
              Ncut=3, Nmax=40, cosmo=None, workers=1, counts=None):
    """py:13-132 (without the Ngrid==360 assert)."""
    if cosmo is None:
        cosmo = FlatLambdaCDM(H0=67.6, Om0=0.31)
    kf = 2 * np.pi / Lbox
    N = radecz.shape[1]
    delta_d, Ngtot, I12d, I13d, I22d, I23d, I33d = FFT_survey_mono(radecz, nbar, w=w, P0_fkp=P0_fkp, Lbox=Lbox, Ngrid=Ngrid,
                                                                    cosmo=cosmo, workers=workers)
    deltak_d = reflect_delta(delta_d, Ngrid)
    delta_r, Nrtot, I12r, I13r, I22r, I23r, I33r = FFT_survey_mono(radecz_r, nbar_r, w=w_r, P0_fkp=P0_fkp, Lbox=Lbox,
                                                                    Ngrid=Ngrid, cosmo=cosmo, workers=workers)
    deltak_r = reflect_delta(delta_r, Ngrid)
    alpha = Ngtot / Nrtot
    deltak = deltak_d - alpha * deltak_r                     # complex128 under numpy >= 2 (alpha is a float64 scalar)
    bispec = _B0_survey(deltak, alpha, alpha * I12r, alpha * I13r, alpha * I22r, alpha * I23r, alpha * I33r, Nmax=Nmax,
                        Ncut=Ncut, step=step, workers=workers, counts=counts)
    bispec['meta'] = {'Lbox': Lbox, 'Ngrid': Ngrid, 'step': step, 'Ncut': Ncut, 'Nmax': Nmax, 'N': N, 'nbar': nbar, 'kf': kf}
    return bispec
