"""CPU tests: the C-ABI library loads and exports every symbol include/psb200.h declares (no compute calls),
host-side logic of the Python layer, and the 2-rank gloo path of the multi-GPU plumbing."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'psb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(psb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from pyspectrum_b200 import _lib
    so = _lib.build()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), 'missing symbol ' + n
    # the ctypes prototype table mirrors the header one to one
    assert sorted(_lib.PROTOTYPES) == names
    assert _lib.lib().psb_version() >= 100
    assert _lib.lib().psb_error_string(-2).decode().startswith('unsupported grid')


def test_host_table_builders_match_reference_expressions():
    from pyspectrum_b200 import _lib
    L = _lib.lib()
    N = 36
    tw = np.empty(2 * N, np.float32)
    assert L.psb_twiddles_f32(N, tw.ctypes.data_as(ctypes.c_void_p)) == 0
    k = np.arange(N)
    np.testing.assert_allclose(tw[0::2] + 1j * tw[1::2], np.exp(2j * np.pi * k / N), atol=1e-7)
    # fcomb tables: |rec| drifts below 1 because the base is rounded to single; window = sinc^4
    rec = np.empty(2 * (N // 2 + 1), np.float64)
    wk = np.empty(N // 2 + 1, np.float32)
    assert L.psb_fcomb_tables(N, rec.ctypes.data_as(ctypes.c_void_p), wk.ctypes.data_as(ctypes.c_void_p)) == 0
    r = rec[0::2] + 1j * rec[1::2]
    np.testing.assert_allclose(np.angle(r), -np.pi * np.arange(N // 2 + 1) / N, atol=1e-6)
    assert abs(abs(r[-1]) - 1) < 1e-5 and r[0] == 1
    x = np.pi * np.arange(1, N // 2 + 1) / N
    np.testing.assert_allclose(wk[1:], (np.sin(x) / x) ** 4, rtol=2e-6)
    # rsd bins: nint(Nbin*sqrt(m)/(N/2)) in float32; irk: int(sqrt(m)/step+0.5) in float32
    mmax = 3 * (N // 2) ** 2
    b = np.empty(mmax + 1, np.uint16)
    assert L.psb_rsd_bin_table(N, N // 2, mmax, b.ctypes.data_as(ctypes.c_void_p)) == 0
    m = np.arange(mmax + 1)
    assert np.array_equal(b, np.floor(np.sqrt(m) + 0.5).astype(np.uint16))
    t = np.empty(mmax + 1, np.uint16)
    assert L.psb_irk_table_f32(ctypes.c_float(3.0), mmax, t.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(t, (np.sqrt(m) / 3 + 0.5).astype(int).astype(np.uint16))
    trig = np.empty(4, np.float32)
    assert L.psb_rsd_trig(2, trig.ctypes.data_as(ctypes.c_void_p)) == 0 and list(trig) == [1, 0, 1, 0]
    assert L.psb_rsd_trig(7, trig.ctypes.data_as(ctypes.c_void_p)) == -1


def test_triangle_list_and_tiles_match_reference_loop():
    from pyspectrum_b200 import _lib
    from pyspectrum_b200.pyspectrum import triangle_list
    from oracle import pyspec_oracle as O
    for (Nmax, Ncut, step) in [(40, 3, 3), (10, 3, 3), (8, 1, 1), (6, 3, 2), (80, 3, 2)]:
        tri = triangle_list(Nmax, Ncut, step)
        assert np.array_equal(tri, O.triangle_list(Nmax, Ncut, step))
    assert len(triangle_list(40, 3, 3)) == 6350 and len(triangle_list(80, 3, 2)) == 46700
    tri = np.ascontiguousarray(triangle_list(40, 3, 3))
    nt = ctypes.c_int(0)
    L = _lib.lib()
    assert L.psb_bk_build_tiles(tri.ctypes.data_as(ctypes.c_void_p), len(tri), 1, None, ctypes.byref(nt)) == 0
    tiles = np.empty((nt.value, 68), np.int32)
    assert L.psb_bk_build_tiles(tri.ctypes.data_as(ctypes.c_void_p), len(tri), 1, tiles.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nt)) == 0
    slots = tiles[:, 4:]
    assert sorted(slots[slots >= 0].tolist()) == list(range(6350))          # every triangle exactly once
    for ti in range(len(tiles)):
        i0, j0, l0 = tiles[ti, :3]
        for q in np.nonzero(slots[ti] >= 0)[0]:
            a, b, c = q // 16, (q // 4) % 4, q % 4
            assert tuple(tri[slots[ti, q]] - 1) == (i0 + a, j0 + b, l0 + c)


def test_fortran_record_roundtrip_and_reflect_delta(tmp_path):
    from pyspectrum_b200 import pyspectrum as P
    from oracle import pyspec_oracle as O
    c = np.random.default_rng(0).uniform(size=(5, 5, 5))
    f = str(tmp_path / 'counts.test')
    P._write_fortran_record(f, c)
    assert np.array_equal(P._read_fortran_record(f, 5), c)
    from scipy.io import FortranFile                      # the reference's reader (py:970-972)
    ff = FortranFile(f, 'r')
    assert np.array_equal(ff.read_reals().reshape(5, 5, 5), c)
    ff.close()
    N = 8
    rng = np.random.default_rng(1)
    d = (rng.normal(size=(N // 2 + 1, N, N)) + 1j * rng.normal(size=(N // 2 + 1, N, N))).astype(np.complex64)
    assert np.array_equal(P.reflect_delta(d, Ngrid=N), O.reflect_delta(d, N))


def test_fortran_records_are_validated(tmp_path):
    """Record markers are checked (a truncated or foreign counts cache is refused, not reshaped), and gfortran's subrecords
    (records of 2 GiB or more: negative leading marker = continued) are joined."""
    from pyspectrum_b200 import pyspectrum as P, util as UT
    c = np.arange(27, dtype='<f8').reshape(3, 3, 3)
    f = str(tmp_path / 'counts.ok')
    P._write_fortran_record(f, c)
    raw = open(f, 'rb').read()
    for bad in (raw[:-5], raw[:-4] + b'\x01\x00\x00\x00', raw[:40]):          # truncated, trailing marker differs, cut short
        g = str(tmp_path / 'counts.bad')
        open(g, 'wb').write(bad)
        with pytest.raises(ValueError):
            P._read_fortran_record(g, 3)
    with pytest.raises(ValueError):
        P._read_fortran_record(f, 4)                                             # a file for another Nmax
    b = c.tobytes()
    m = lambda v: np.array([v], '<i4').tobytes()
    split = m(-104) + b[:104] + m(104) + m(112) + b[104:] + m(-112)            # two subrecords of one logical record
    g = str(tmp_path / 'counts.split')
    open(g, 'wb').write(split)
    assert UT.fortran_records(g) == [b]
    assert np.array_equal(P._read_fortran_record(g, 3), c)


def test_api_signatures_match_reference():
    import inspect
    from pyspectrum_b200 import pyspectrum as P
    assert str(inspect.signature(P.Pk_periodic)) == "(xyz, w=None, Lbox=2600, Ngrid=360, fft='pyfftw', silent=True)"
    assert str(inspect.signature(P.Pk_periodic_rsd)) == "(xyz, w=None, Lbox=2600, Ngrid=360, rsd=2, Nmubin=10, fft='pyfftw', code='fortran', silent=True)"
    assert str(inspect.signature(P.Bk_periodic)) == "(xyz, w=None, Lbox=2600, Ngrid=360, step=3, Ncut=3, Nmax=40, fft='pyfftw', nthreads=1, silent=True)"
    assert str(inspect.signature(P.FFT_periodic)) == "(xyz, w=None, Lbox=2600.0, Ngrid=360, fft='pyfftw', silent=True)"


def test_no_cpu_fallback_and_f2py_style_argument_checks():
    import torch
    from pyspectrum_b200 import pyspectrum as P, estimator as E
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            P.Pk_periodic(np.zeros((3, 4)), Lbox=10., Ngrid=24)
    with pytest.raises(ValueError):      # intent(inout) array must be float32 + Fortran order, as f2py insists
        E.assign_quad(np.zeros((3, 2), np.float32), np.ones(2, np.float32), np.zeros((8, 4, 4), np.float32), 1., 0, 0, 0, 0, 0)
    with pytest.raises(ValueError):
        E.fcomb_periodic(np.zeros((4, 4, 4), np.complex64), 1.)
    with pytest.raises(ValueError):
        E.bk_counts(np.zeros((3, 3, 3), np.float32, order='F'), 24, 3., 3)


def test_product_code_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, 'pyspectrum_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert 'oracle' not in open(os.path.join(d, f)).read().lower().replace('psb_host', ''), f


def _gloo_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    from pyspectrum_b200 import dist as D
    r, w = D.init('gloo')
    assert (r, w) == (rank, world)
    D.barrier()
    mx = D.max_over_ranks([10.0 + rank, 5.0 - rank])
    sm = D.sum_over_ranks([1.0, float(rank)])
    out.put((rank, mx, sm, D.catalogue_seed(2, rank)))
    dist.destroy_process_group()


def test_two_rank_gloo_timing_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, mx, sm, seed in res:
        assert mx == [11.0, 5.0] and sm == [2.0, 1.0] and seed == 2 + rank
    from pyspectrum_b200 import dist as D
    assert D.seconds_per_catalogue(3.0, 5, 2) == 0.3


@pytest.mark.parametrize('layout', [0, 1])
@pytest.mark.parametrize('Nmax,Ncut,step', [(40, 3, 3), (12, 1, 1), (10, 3, 2), (11, 3, 2), (30, 3, 1), (7, 3, 3), (64, 3, 1), (72, 3, 1), (80, 3, 2), (5, 4, 4)])
def test_tensor_core_plan_covers_every_triangle_once(Nmax, Ncut, step, layout):
    """build_tc_plan (host plan of psb_bk_triangle_sums_tc): every triangle (i,j,l) is owned by exactly one pass, its row decodes
    back to (i,j) through the lane's field slots in the declared layout, its column is l; rows are unique; lanes stay in range."""
    from pyspectrum_b200 import pyspectrum as P
    tri = P.triangle_list(Nmax, Ncut, step)
    s0 = Ncut // step
    S = Nmax - s0 + 1
    NT, MT, used, passes = P.build_tc_plan(tri, s0, Nmax, layout)
    MT_max = 4 if NT <= 64 else 256 // NT
    assert NT % 16 == 0 and NT >= S and 1 <= MT <= MT_max and MT * NT <= 256
    npairs = len({(int(a), int(b)) for a, b, _ in tri})
    # 2x2 blocks need four tiles; <= 128 pair rows take one tile of plain rows instead (cheaper, measured)
    assert used == (layout if (MT_max == 4 and npairs > 128) else 0) and (MT == 4 or used == 0)
    if used == 0:                                       # the fewest tiles that do not add a pass
        _, _, _, full = P.build_tc_plan(tri, s0, Nmax, 0, MT_max)
        assert len(passes) == len(full)
        if MT > 1:
            assert len(P.build_tc_plan(tri, s0, Nmax, 0, MT - 1)[3]) > len(full)
    owner = np.zeros(len(tri), int)
    for lij, rc in passes:
        assert lij.shape == (128, 5) and lij.dtype == np.int32 and rc.shape == (len(tri), 2) and rc.dtype == np.int32
        assert lij.max() < S and lij.min() >= -1
        mine = rc[:, 0] >= 0
        owner += mine
        rows, cols = rc[mine, 0], rc[mine, 1]
        assert rows.max() < MT * 128
        m, ln = rows // 128, rows % 128
        sl = lij[ln]
        if used == 0:                                   # {i; j_0..j_3}
            fi, fj = sl[:, 0], sl[np.arange(len(ln)), 1 + m]
        else:                                           # {i0, i1, j0, j1}: tile m = 2*(which i) + (which j)
            fi, fj = sl[np.arange(len(ln)), m // 2], sl[np.arange(len(ln)), 2 + m % 2]
        assert np.array_equal(fi + s0, tri[mine, 0]) and np.array_equal(fj + s0, tri[mine, 1]) and np.array_equal(cols + s0, tri[mine, 2])
        pairs = {(int(a), int(b)) for a, b in zip(tri[mine, 0], tri[mine, 1])}
        assert len({int(r) for r in rows}) == len(pairs)            # one row per (i,j) pair
    assert np.all(owner == 1)
