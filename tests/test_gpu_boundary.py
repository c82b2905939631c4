"""Boundary proof: the reference's Python layer running ON TOP OF pyspectrum_b200.estimator, the CUDA
drop-in for the f2py module `estimator` (pyspectrum.py:8) -- f2py argument order, intent(inout) buffers, Fortran order -- and
compared with the goldens the unmodified reference produced (tests/golden/small_*.npz).

Two layers, chosen by what the machine has:
  * where /root/reference exists (the build container): the UNMODIFIED reference pyspectrum.py (scratch copy with the one-token
    fix at py:713, pyfftw -> scipy shim) imported with sys.modules['estimator'] = pyspectrum_b200.estimator;
  * on the GPU box (/root/reference is not shipped there and its sources may not be copied into this repo): the oracle's
    restatement of that Python layer (oracle/pyspec_oracle.py, itself pinned bit-for-bit to the same goldens on CPU by
    tests/test_oracle_golden.py) with its three native calls re-pointed at pyspectrum_b200.estimator.  The call sites are the
    reference's: assign_quad(xyzs, w, _delta, kf_ks, 0, 0,0,0,0) py:951, fcomb_periodic(ifft_delta, np.sum(w)) py:957,
    pk_pbox_rsd(dtl, rsd, Lbox, Nbins, Nmubin, Ngrid) py:632.
Either way every native call of the periodic path goes through the C ABI (psb_host_*) with HOST numpy buffers."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5


class _OracleLayer(object):
    """oracle.pyspec_oracle's Python layer with estimator's natives swapped in (same names as the reference module)."""

    def __init__(self, est, monkeypatch):
        from oracle import pyspec_oracle as O
        monkeypatch.setattr(O, 'assign_quad', lambda r, w, dtl, kf_ks, offset, ia, ib, ic, id: est.assign_quad(r, w, dtl, kf_ks, offset, ia, ib, ic, id))
        monkeypatch.setattr(O, 'fcomb_periodic', lambda dcl, n: est.fcomb_periodic(dcl, n))
        monkeypatch.setattr(O, 'pk_pbox_rsd', lambda dtl, irsd, lbox, nbin, nmu: est.pk_pbox_rsd(dtl, irsd, int(lbox), nbin, nmu))
        self.O = O
        self.Pk_periodic = O.Pk_periodic
        self.Pk_periodic_rsd = O.Pk_periodic_rsd
        self.FFT_periodic = O.FFT_periodic

    def reflect_delta(self, delta, Ngrid):
        return self.O.reflect_delta(delta, Ngrid)

    def _Bk_periodic(self, delta_fft, step, Ncut, Nmax):
        return self.O._Bk_periodic(delta_fft, step=step, Ncut=Ncut, Nmax=Nmax)


@pytest.fixture()
def layer(monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pyspectrum_b200 import estimator as est
    if os.path.isdir('/root/reference/pyspectrum'):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
        import make_golden
        saved = {k: sys.modules.get(k) for k in ('estimator', 'pyfftw', 'astropy', 'astropy.cosmology')}
        ref = make_golden.import_reference(estimator_module=est)
        yield ref, 'reference'
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    else:
        yield _OracleLayer(est, monkeypatch), 'oracle-layer'


@pytest.mark.parametrize('tag', ['A', 'B', 'C'])
def test_reference_python_layer_on_cuda_estimator(layer, golden_dir, tag):
    ref, kind = layer
    g = dict(np.load(os.path.join(golden_dir, 'small_%s.npz' % tag)))
    N, L, w = int(g['Ngrid']), float(g['Lbox']), g.get('w')
    xyz = g['xyz']
    pk = ref.Pk_periodic(xyz, w=w, Lbox=L, Ngrid=N)
    assert np.array_equal(pk['counts'], g['pk_counts'])
    np.testing.assert_allclose(pk['k'], g['pk_k'], rtol=1e-12)
    np.testing.assert_allclose(pk['p0k'] + pk['p0k_sn'], g['pk_p0k'] + g['pk_p0k_sn'], rtol=RTOL)
    for rsd in (0, 1, 2):
        pr = ref.Pk_periodic_rsd(xyz, w=w, Lbox=L, Ngrid=N, rsd=rsd, Nmubin=5)
        pre = 'rsd%d_mu5_' % rsd
        assert np.array_equal(pr['counts'], g[pre + 'counts'])                    # mode counts through pk_pbox_rsd: bit exact
        assert np.array_equal(pr['counts_kmu'], g[pre + 'counts_kmu'])
        sn = pr['p_sn'][0]
        np.testing.assert_allclose(pr['p0k'] + sn, g[pre + 'p0k'] + sn, rtol=RTOL)
        scale = np.abs(g[pre + 'p0k'] + sn)
        assert np.all(np.abs(pr['p2k'] - g[pre + 'p2k']) <= 5 * RTOL * scale)
    delta = ref.FFT_periodic(xyz, w=w, Lbox=L, Ngrid=N)
    assert np.abs(np.asarray(delta) - g['delta_half']).max() <= 3e-6 * np.abs(g['delta_half']).max()
    step, Ncut, Nmax = 3, 3, 4
    pre = 'bk_s%d_c%d_m%d_' % (step, Ncut, Nmax)
    delta_fft = ref.reflect_delta(delta, Ngrid=N)
    bk = ref._Bk_periodic(delta_fft, step=step, Ncut=Ncut, Nmax=Nmax)
    for key in ['i_k1', 'i_k2', 'i_k3']:
        assert np.array_equal(bk[key], g[pre + key])
    ww = np.ones(xyz.shape[1]) if w is None else w
    nbar, kf = np.sum(ww) / L ** 3, 2 * np.pi / L
    p1 = bk['p0k1'] * (2 * np.pi) ** 3 / kf ** 3
    np.testing.assert_allclose(p1, g[pre + 'p0k1'] + 1. / nbar, rtol=RTOL)
    b = bk['b123'] * (2 * np.pi) ** 6 / kf ** 6
    scale = np.abs(g[pre + 'b123'] + g[pre + 'b123_sn'])
    assert np.all(np.abs(b - (g[pre + 'b123'] + g[pre + 'b123_sn'])) <= RTOL * scale + 1e-7 * scale.max())
