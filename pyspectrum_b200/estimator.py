"""pyspectrum_b200.estimator -- drop-in for the f2py extension module `estimator` (pyspectrum.py:8,
built from pyspectrum/estimator.f by setup.py:11-52).  Same function names, argument order, dtype /
Fortran-order requirements and in-place semantics as the f2py wrappers (`f2py -h` signatures, SURVEY
8b level 2); the work is done by the psb_host_* entry points of libpsb200.so on the GPU.

    assign_quad(r,w,dtl,kf_ks,offset,ia,ib,ic,id,[np,ngrid])          estimator.f:284
    fcomb_periodic(dcl,n,[ngrid])                                     estimator.f:605
    fcomb_survey(dcl,[ngrid])                                         estimator.f:677
    ffting(dtl,n,[ngrid])                                             estimator.f:266
    k,p0,p2,p4,nk,km,mk,pkm,nkm = pk_pbox_rsd(dtl,irsd,lbox,nbin,nmu,[ngrid])   estimator.f:155
    bk_counts(coun,nside,step,ncut,[nmax])                            estimator.f:2
    fivedelta2g_1(dcgxx,dcgyy,dcgzz,[ngrid])                          estimator.f:514   (f2py lower-cases the names; the
    fivedelta2g_2(dcg,dcgxx,dcgxy,dcgyz,dcgzx,[ngrid])                estimator.f:541    CamelCase spellings of py:1106, 1130
    build_quad(dclr1,dclr2,irsd,[ngrid])                              estimator.f:574    are kept as aliases)
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _inout(a, dtype, name):
    # f2py raises for intent(inout) arrays that are not exactly typed and Fortran-contiguous
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.f_contiguous:
        raise ValueError('failed in converting argument `%s` to C/Fortran array: need %s, Fortran-contiguous (intent inout)'
                         % (name, np.dtype(dtype).name))
    return a


def assign_quad(r, w, dtl, kf_ks, offset, ia, ib, ic, id, np_=None, ngrid=None):
    r = np.asfortranarray(r, dtype=np.float32)               # intent(in): silently cast/copied, as f2py does
    w = np.ascontiguousarray(w, dtype=np.float32)
    _inout(dtl, np.float32, 'dtl')
    n = dtl.shape[1]
    if r.ndim != 2 or r.shape[0] != 3 or w.shape != (r.shape[1],) or dtl.shape != (2 * n, n, n):
        raise ValueError('assign_quad: shape mismatch (r (3,np), w (np), dtl (2*ngrid,ngrid,ngrid))')
    check(_lib.lib().psb_host_assign_quad(_p(r), _p(w), _p(dtl), r.shape[1], n, np.float32(kf_ks), np.float32(offset),
                                          int(ia), int(ib), int(ic), int(id)), 'assign_quad')


def fcomb_periodic(dcl, n, ngrid=None):
    _inout(dcl, np.complex64, 'dcl')
    N = dcl.shape[0]
    if dcl.shape != (N, N, N):
        raise ValueError('fcomb_periodic: dcl must be (ngrid,ngrid,ngrid)')
    check(_lib.lib().psb_host_fcomb_periodic(_p(dcl), np.float32(n), N), 'fcomb_periodic')


def fcomb_survey(dcl, ngrid=None):
    _inout(dcl, np.complex64, 'dcl')
    N = dcl.shape[0]
    if dcl.shape != (N, N, N):
        raise ValueError('fcomb_survey: dcl must be (ngrid,ngrid,ngrid)')
    check(_lib.lib().psb_host_fcomb_survey(_p(dcl), N), 'fcomb_survey')


def ffting(dtl, n=None, ngrid=None):
    _inout(dtl, np.complex64, 'dtl')
    N = dtl.shape[0]
    check(_lib.lib().psb_host_ffting(_p(dtl), N), 'ffting')


def pk_pbox_rsd(dtl, irsd, lbox, nbin, nmu, ngrid=None):
    dtl = np.asfortranarray(dtl, dtype=np.complex64)
    N = dtl.shape[1]
    if dtl.shape != (N // 2 + 1, N, N):
        raise ValueError('pk_pbox_rsd: dtl must be (ngrid/2+1,ngrid,ngrid)')
    k, p0, p2, p4, nk = [np.zeros(nbin, dtype=np.float64) for _ in range(5)]
    km, mk, pkm, nkm = [np.zeros((nbin, nmu), dtype=np.float64, order='F') for _ in range(4)]
    check(_lib.lib().psb_host_pk_pbox_rsd(_p(dtl), _p(k), _p(p0), _p(p2), _p(p4), _p(nk), _p(km), _p(mk), _p(pkm), _p(nkm),
                                          int(irsd), int(lbox), int(nbin), int(nmu), N), 'pk_pbox_rsd')
    return k, p0, p2, p4, nk, km, mk, pkm, nkm


def bk_counts(coun, nside, step, ncut, nmax=None):
    _inout(coun, np.float64, 'coun')
    nmax = coun.shape[0]
    if coun.shape != (nmax, nmax, nmax):
        raise ValueError('bk_counts: coun must be (nmax,nmax,nmax)')
    check(_lib.lib().psb_host_bk_counts(_p(coun), int(nside), np.float32(step), int(ncut), nmax), 'bk_counts')


def _half(a, name, inout=False):
    if inout:
        _inout(a, np.complex64, name)
    else:
        a = np.asfortranarray(a, dtype=np.complex64)
    n = a.shape[1] if a.ndim == 3 else 0
    if a.ndim != 3 or a.shape != (n // 2 + 1, n, n):
        raise ValueError('%s must be (ngrid/2+1,ngrid,ngrid) complex64' % name)
    return a


def fivedelta2g_1(dcgxx, dcgyy, dcgzz, ngrid=None):
    """dcgxx <- 7.5 (dcgxx kxh^2 + dcgyy kyh^2 + dcgzz kzh^2) for k != 0, with the Fortran's implicitly INTEGER kxh, kyh, kzh."""
    dcgxx = _half(dcgxx, 'dcgxx', True)
    dcgyy, dcgzz = _half(dcgyy, 'dcgyy'), _half(dcgzz, 'dcgzz')
    if dcgyy.shape != dcgxx.shape or dcgzz.shape != dcgxx.shape:
        raise ValueError('fivedelta2g_1: shape mismatch')
    check(_lib.lib().psb_host_fivedelta2g_1(_p(dcgxx), _p(dcgyy), _p(dcgzz), dcgxx.shape[1]), 'fivedelta2g_1')


def fivedelta2g_2(dcg, dcgxx, dcgxy, dcgyz, dcgzx, ngrid=None):
    """dcgxx <- dcgxx + 7.5 (2 dcgxy kxh kyh + 2 dcgyz kyh kzh + 2 dcgzx kzh kxh) - 2.5 dcg for k != 0 (integer kxh, kyh, kzh)."""
    dcgxx = _half(dcgxx, 'dcgxx', True)
    ins = [_half(a, n) for a, n in ((dcg, 'dcg'), (dcgxy, 'dcgxy'), (dcgyz, 'dcgyz'), (dcgzx, 'dcgzx'))]
    if any(a.shape != dcgxx.shape for a in ins):
        raise ValueError('fivedelta2g_2: shape mismatch')
    check(_lib.lib().psb_host_fivedelta2g_2(_p(ins[0]), _p(dcgxx), _p(ins[1]), _p(ins[2]), _p(ins[3]), dcgxx.shape[1]), 'fivedelta2g_2')


def build_quad(dclr1, dclr2, irsd, ngrid=None):
    """dclr2 <- (7.5 mu^2 - 2.5) dclr1 for k != 0, mu along axis irsd (1 = x, 2 = y, 3 = z)."""
    dclr2 = _half(dclr2, 'dclr2', True)
    dclr1 = _half(dclr1, 'dclr1')
    if dclr1.shape != dclr2.shape:
        raise ValueError('build_quad: shape mismatch')
    check(_lib.lib().psb_host_build_quad(_p(dclr1), _p(dclr2), int(irsd), dclr2.shape[1]), 'build_quad')


FiveDelta2g_1, FiveDelta2g_2 = fivedelta2g_1, fivedelta2g_2
