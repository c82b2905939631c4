"""pyspectrum_b200.util -- host utilities either side of the estimator hot path (pySpectrum's pyspectrum/util.py).

    ijl_order            (util.py:8-24)    triangle re-ordering to the 'GM' convention
    radecz_to_cartesian  (util.py:27-51)   (RA, Dec, z) -> comoving Cartesian Mpc/h
    applyRSD             (util.py:54-75)   redshift-space positions in a periodic box
    read_fortFFT         (util.py:78-107)  Fortran-unformatted half field -> full Hermitian field

Pure numpy host code (a few passes over the catalogue; nothing here is on the GPU path).  The reference takes
its cosmology from astropy (`FlatLambdaCDM`, un-pinned, not in this image); `FlatLambdaCDM` below restates the
two members the reference touches -- `comoving_distance` and `efunc` for a flat matter + Lambda universe, which is
what `astropy.cosmology.FlatLambdaCDM(H0, Om0)` is with its default Tcmb0=0 (no radiation, no neutrinos) -- and any
object with the same members (an astropy cosmology) is accepted in its place.

Deliberate differences from the reference:
  * `radecz_to_cartesian` does NOT convert the caller's RA/Dec rows to radians in place (util.py:41-42 does, so a
    second call on the same array silently gives different positions);
  * `applyRSD` works: the reference uses `FlatLambdaCDM` without importing it (NameError on every call);
  * `read_fortFFT` uses integer division for the record shape (util.py:84 passes a float to reshape).
"""
import numpy as np

__all__ = ['ijl_order', 'radecz_to_cartesian', 'applyRSD', 'read_fortFFT', 'FlatLambdaCDM']

_C_KMS = 299792.458                                   # speed of light [km/s]
_GL_X, _GL_W = np.polynomial.legendre.leggauss(24)    # per-panel Gauss-Legendre rule


class FlatLambdaCDM(object):
    """Flat LambdaCDM background: E(z)^2 = Om0 (1+z)^3 + (1 - Om0).  Members follow astropy's names."""

    def __init__(self, H0=67.6, Om0=0.31):
        self.H0 = float(H0)
        self.Om0 = float(Om0)
        self.Ode0 = 1. - self.Om0
        self.h = self.H0 / 100.

    def efunc(self, z):
        z = np.asarray(z, dtype=np.float64)
        return np.sqrt(self.Om0 * (1. + z) ** 3 + self.Ode0)

    def _integral(self, z):
        """int_0^z dz'/E(z') by composite 24-point Gauss-Legendre, panels of at most 0.5 in z (analytic integrand:
        relative error < 1e-13)."""
        out = np.zeros_like(z)
        if z.size == 0:
            return out
        npanel = max(int(np.ceil(np.max(np.abs(z)) / 0.5)), 1)
        for p in range(npanel):
            a = z * (p / float(npanel))
            b = z * ((p + 1) / float(npanel))
            half = 0.5 * (b - a)
            mid = 0.5 * (b + a)
            zz = mid[:, None] + half[:, None] * _GL_X[None, :]
            out += half * np.sum(_GL_W[None, :] / self.efunc(zz), axis=1)
        return out

    def comoving_distance(self, z):
        """Line-of-sight comoving distance in Mpc: (c/H0) int_0^z dz'/E(z').  Catalogue-sized inputs (random catalogues
        hold 1e7-1e9 redshifts) go through a cubic Hermite table on 4097 nodes over [0, zmax] whose node values come from
        the quadrature and whose slopes are the exact integrand 1/E (interpolation error ~ dz^4/384 times the fourth
        derivative: below 1e-15)."""
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        if z.size <= 8192 or np.min(z) < 0.:
            return (_C_KMS / self.H0) * self._integral(z)
        nn = 4096
        zmax = float(np.max(z))
        if zmax == 0.:
            return np.zeros_like(z)
        dz = zmax / nn
        zn = np.arange(nn + 1) * dz
        fn = self._integral(zn)
        gn = 1. / self.efunc(zn)
        k = np.minimum((z / dz).astype(np.int64), nn - 1)
        t = z / dz - k
        f0, f1, g0, g1 = fn[k], fn[k + 1], gn[k] * dz, gn[k + 1] * dz
        t2 = t * t
        t3 = t2 * t
        out = (2. * t3 - 3. * t2 + 1.) * f0 + (t3 - 2. * t2 + t) * g0 + (-2. * t3 + 3. * t2) * f1 + (t3 - t2) * g1
        return (_C_KMS / self.H0) * out


def ijl_order(i_k, j_k, l_k, typ='GM'):
    """util.py:8-24: indices that re-order triangles (i>=j>=l) with l slowest, then j, then i (Gil-Marin's order).
    Returns an array of shape (Ntri, m) like the reference: one row per distinct (i,j,l), holding the indices of the
    input entries with that triple (m=1 when the triples are unique)."""
    i_k, j_k, l_k = np.asarray(i_k), np.asarray(j_k), np.asarray(l_k)
    if typ != 'GM':
        raise NotImplementedError
    n = len(i_k)
    # lexicographic sort on (l, j, i); stable, so duplicates keep their input order like the reference's masks
    order = np.lexsort((i_k, j_k, l_k))
    keys = np.stack([l_k[order], j_k[order], i_k[order]], axis=1)
    if n == 0:
        return np.array([])
    new = np.ones(n, dtype=bool)
    new[1:] = np.any(keys[1:] != keys[:-1], axis=1)
    starts = np.flatnonzero(new)
    groups = np.split(order, starts[1:])
    return np.array(groups)


def radecz_to_cartesian(radecz, cosmo=None):
    """util.py:27-51: [3,N] (RA deg, Dec deg, z) -> [3,N] comoving Cartesian coordinates in Mpc/h."""
    radecz = np.asarray(radecz)
    assert radecz.shape[0] == 3, "radecz has to be have shape [3,N]"
    if cosmo is None:
        cosmo = FlatLambdaCDM(H0=67.6, Om0=0.31)      # the default of pyspectrum.py:88
    ra = radecz[0] * (np.pi / 180.)
    dec = radecz[1] * (np.pi / 180.)
    rad = cosmo.comoving_distance(radecz[2])
    rad = np.asarray(getattr(rad, 'value', rad), dtype=np.float64) * cosmo.h       # astropy returns a Quantity [Mpc]
    return np.array([rad * np.cos(dec) * np.cos(ra),
                     rad * np.cos(dec) * np.sin(ra),
                     rad * np.sin(dec)])


def applyRSD(xyz, vxyz, redshift, h=0.7, omega0_m=0.3, LOS=None, Lbox=None):
    """util.py:54-75: x_s = (x + v (1+z)/(100 E(z)) + L) mod L along the line of sight, velocities in km/s.
    numpy arrays in -> numpy out (host, as the reference); torch CUDA tensors in -> the same arithmetic on the device (psb_apply_rsd)."""
    dev_in = _is_cuda_tensor(xyz)                        # catalogue already on the GPU: psb_apply_rsd, result stays there
    if not dev_in:
        xyz, vxyz = np.asarray(xyz), np.asarray(vxyz)
    assert xyz.shape[0] == 3
    assert vxyz.shape[0] == 3
    if LOS is None:
        raise ValueError("specify line of sight")
    if Lbox is None:
        raise ValueError("specify box size")
    i_los = {'x': 0, 'y': 1, 'z': 2}[LOS]
    cosmo = FlatLambdaCDM(H0=100. * h, Om0=omega0_m)
    rsd_factor = (1 + redshift) / (100 * cosmo.efunc(redshift))
    if dev_in:
        return _apply_rsd_device(xyz, vxyz, i_los, float(rsd_factor), float(Lbox))
    xyz_rsd = xyz.copy()
    xyz_rsd[i_los] += rsd_factor * vxyz[i_los] + Lbox
    xyz_rsd[i_los] = (xyz_rsd[i_los] % Lbox)
    return xyz_rsd


def _is_cuda_tensor(a):
    try:
        import torch
        return isinstance(a, torch.Tensor) and a.is_cuda
    except ImportError:
        return False


def _apply_rsd_device(xyz, vxyz, i_los, rsd_factor, Lbox):
    """applyRSD for torch CUDA tensors (3 x N): float64 arithmetic of the numpy path bit for bit, no host round trip."""
    import ctypes
    import torch
    from . import _lib
    x = xyz.double().contiguous()
    v = vxyz.to(x.device)[i_los].double().contiguous()
    out = torch.empty_like(x)
    st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().psb_apply_rsd(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(v.data_ptr()), int(x.shape[1]), int(i_los),
                                            rsd_factor, Lbox, ctypes.c_void_p(out.data_ptr()), st), 'psb_apply_rsd')
    return out


def fortran_records(path):
    """Logical records of a Fortran sequential unformatted file with 4-byte markers (gfortran / ifort default) as a list of
    byte strings.  Every (sub)record is `marker | payload | marker`; gfortran splits records of 2 GiB or more into subrecords:
    a negative leading marker says another subrecord follows, the trailing marker repeats the length (negative if a subrecord
    precedes).  Raises ValueError on a truncated file or markers that do not match."""
    raw = np.fromfile(path, dtype=np.uint8)
    out, parts, o = [], [], 0
    while o < raw.size:
        if o + 4 > raw.size:
            raise ValueError('%s: truncated record marker at byte %d' % (path, o))
        head = int(raw[o:o + 4].view('<i4')[0])
        n = abs(head)
        if o + 8 + n > raw.size:
            raise ValueError('%s: record of %d bytes at byte %d runs past the end of the file' % (path, n, o))
        tail = int(raw[o + 4 + n:o + 8 + n].view('<i4')[0])
        if abs(tail) != n:
            raise ValueError('%s: record markers disagree (%d vs %d) at byte %d' % (path, head, tail, o))
        parts.append(raw[o + 4:o + 4 + n])
        o += 8 + n
        if head >= 0:                                    # last subrecord of this logical record
            out.append(parts[0].tobytes() if len(parts) == 1 else np.concatenate(parts).tobytes())
            parts = []
    if parts:
        raise ValueError('%s: file ends inside a split record' % path)
    return out


def read_fortFFT(file=None):
    """util.py:78-107: read `Ngrid` + the (Ngrid/2+1, Ngrid, Ngrid) complex64 half field from a Fortran-unformatted file
    and return the full Hermitian field (same completion as pyspectrum.reflect_delta)."""
    from .pyspectrum import reflect_delta
    recs = fortran_records(file)
    if len(recs) < 2 or len(recs[0]) < 4:
        raise ValueError('%s: expected a record with Ngrid and a record with the half field' % file)
    Ngrid = int(np.frombuffer(recs[0][:4], '<i4')[0])
    want = 8 * (Ngrid // 2 + 1) * Ngrid * Ngrid
    if len(recs[1]) != want:
        raise ValueError('%s: the field record holds %d bytes, Ngrid=%d needs %d' % (file, len(recs[1]), Ngrid, want))
    delt = np.frombuffer(recs[1], '<c8')
    delt = np.reshape(delt, (Ngrid // 2 + 1, Ngrid, Ngrid), order='F')
    return reflect_delta(delt, Ngrid=Ngrid)
