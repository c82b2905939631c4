"""pyspectrum_b200.pyspectrum -- drop-in for the periodic-box API of pySpectrum's pyspectrum/pyspectrum.py.

Same public functions, arguments and output dictionaries as the reference:

    Pk_periodic      (pyspectrum.py:644-728)
    Pk_periodic_rsd  (pyspectrum.py:460-538, code='fortran' branch 628-641)
    Bk_periodic      (pyspectrum.py:285-356 + _Bk_periodic 359-457)
    FFT_periodic     (pyspectrum.py:909-959), reflect_delta (1134-1157), _counts_Bk123 (962-1030)
    B0_survey        (pyspectrum.py:13-132 + _B0_survey 135-282 + FFT_survey_mono 731-826): the survey-geometry monopole
                     bispectrum, same kernels with the assignment offset 0.5*Ngrid and fcomb_survey

Behind them the f2py module `estimator` and the pyfftw calls are replaced by hand-written sm_100a
kernels reached through the C ABI of include/psb200.h (ctypes; torch tensors are only device buffers
and the stream).  There is no CPU fallback: without libpsb200.so or a CUDA device the calls raise.

Deliberate differences from the reference (all in DESIGN.md):
  * `Ngrid == 360` assert of Bk_periodic (py:332) is relaxed to "even N = 2^a 3^b 5^c";
  * Pk_periodic bins |delta_fft|^2 (py:713 indexes the wrong array and raises on numpy >= 1.13);
  * `fft`, `nthreads` are accepted and ignored; `code='python'` selects the float64 (k,mu) estimator of py:545-626;
  * triangle counts are exact integers computed in float64 on the GPU and cached (memory + a file in the
    reference's own Fortran-record format) instead of the reference's pure-Python cold path.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import check

__all__ = ['Pk_periodic', 'Pk_periodic_rsd', 'Bk_periodic', 'FFT_periodic', 'reflect_delta',
           'B0_survey', 'FFT_survey_mono', '_B0_survey', '_Bk_periodic', '_Pk_periodic_rsd', '_counts_Bk123_f77',
           '_counts_Bk123', 'dat_dir', 'PeriodicPipeline']

_DAT_DIR = os.environ.get('PYSPECTRUM_B200_DAT', os.path.join(os.path.dirname(os.path.realpath(__file__)), 'dat'))


def dat_dir():
    """Where triangle-count caches live (the reference's dat_dir(), pyspectrum/__init__.py:6)."""
    return _DAT_DIR


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Range(object):
    """Profiling hooks: with PSB_NVTX=1 every stage is an NVTX range (visible in nsys / ncu --nvtx); with
    PSB_TIMERS=1 the Bk/Pk entry points add meta['stage_ms'] (CUDA events on the launching stream).  Both are off by default."""
    nvtx = os.environ.get('PSB_NVTX', '0') == '1'
    timers = os.environ.get('PSB_TIMERS', '0') == '1'

    def __init__(self, name, sink=None):
        self.name, self.sink = name, sink

    def __enter__(self):
        if _Range.nvtx:
            torch.cuda.nvtx.range_push('psb:' + self.name)
        if self.sink is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.sink is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.sink.append((self.name, self.e0, e1))
        if _Range.nvtx:
            torch.cuda.nvtx.range_pop()
        return False


def _copy_threads():
    """Worker threads of the host staging copies: up to 8, sharing the cores with the other ranks of the node."""
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 2
    return max(1, min(8, (cores - 1) // ranks))


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('pyspectrum_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def triangle_list(Nmax, Ncut, step):
    """Shell-index triples (i,j,l) in the loop order of pyspectrum.py:415-417."""
    s = Ncut // step
    i = np.arange(s, Nmax + 1)
    I, J, L = np.meshgrid(i, i, i, indexing='ij')
    m = (J <= I) & (L <= J) & (L >= np.maximum(I - J, s))
    return np.stack([I[m], J[m], L[m]], axis=1).astype(np.int32)      # C-order ravel == nested loop order


def build_tc_plan(tri, s0, Nmax, layout=1, MT=None):
    """Host plan of the tensor-core triangle kernel (pure numpy): which pair row (i,j) sits in which TMEM lane and M tile.
    layout 0: a lane holds one i and up to MT of its partners j (one per M tile): a thread loads I_i once for MT rows.
    layout 1 (MT == 4 only): a lane holds a 2x2 block (i0,i1) x (j0,j1), rows (i0,j0), (i0,j1), (i1,j0), (i1,j1) in tiles 0..3:
             four field vectors give four rows of products.
    Lanes are cut into passes of 128.  Returns (NT, MT, layout_used, [(lane_ij int32 [128][5], tri_rc int32 [ntri][2]), ...]):
    lane_ij holds field slots (shell - s0, -1 = unused), tri_rc the (row = tile*128 + lane, column = l - s0) of every triangle
    owned by the pass, (-1,-1) otherwise.
    MT: M tiles per pass for layout 0 (default: the largest that fits TMEM, lowered as long as the number of passes stays the
    same -- the kernel's cost per cell grows with the tiles it forms, whether their rows are used or not)."""
    S = Nmax - s0 + 1
    NT = (S + 15) // 16 * 16
    MT_max = 4 if NT <= 64 else 256 // NT          # accumulator tiles that fit in 256 TMEM columns
    nj = {}
    for i, j in {(int(a), int(b)) for a, b, _ in tri}:
        nj[i] = nj.get(i, 0) + 1
    npass = lambda m: (sum((n + m - 1) // m for n in nj.values()) + 127) // 128
    if layout == 1 and (MT_max != 4 or (MT is not None and MT != 4)):
        layout = 0
    if layout == 1 and MT is None and npass(1) == 1:
        layout = 0                                  # <= 128 pair rows: one tile of plain rows is cheaper than four tiles of 2x2 blocks
    if layout == 0:
        if MT is None:
            MT = MT_max
            while MT > 1 and npass(MT - 1) == npass(MT_max):
                MT -= 1
        MT = max(1, min(int(MT), MT_max))
    else:
        MT = 4
    partners = {}
    for i, j, _ in tri:
        partners.setdefault(int(i), set()).add(int(j))
    lanes = []                                     # 5 field slots per lane (shell indices, -1 = unused) + its rows
    if layout == 0:
        for i in sorted(partners):
            js = sorted(partners[i])
            nl = (len(js) + MT - 1) // MT
            for k in range(nl):
                jl = [js[k + m * nl] if k + m * nl < len(js) else -1 for m in range(MT)]
                lanes.append(([i] + jl + [-1] * (4 - MT), [(i, j) for j in jl]))
    else:
        ivals = sorted(partners)
        for a in range(0, len(ivals), 2):
            grp = ivals[a:a + 2]
            i0, i1 = grp[0], (grp[1] if len(grp) > 1 else -1)
            js = sorted(set().union(*[partners[i] for i in grp]))
            nl = (len(js) + 1) // 2               # consecutive lanes take consecutive j: rows one apart -> conflict-free LDS.128
            for k in range(nl):
                j0, j1 = js[k], (js[k + nl] if k + nl < len(js) else -1)
                lanes.append(([i0, i1, j0, j1, -1], [(i0, j0), (i0, j1), (i1, j0), (i1, j1)]))
    passes = []
    ti, tj, tl = tri[:, 0], tri[:, 1], tri[:, 2]
    for l0 in range(0, len(lanes), 128):
        sub = lanes[l0:l0 + 128]
        lij = np.full((128, 5), -1, np.int32)
        row_of = {}
        for ln, (slots, rows) in enumerate(sub):
            lij[ln] = [v - s0 if v >= 0 else -1 for v in slots]
            for m, (i, j) in enumerate(rows):
                if i >= 0 and j >= 0:
                    row_of[(i, j)] = m * 128 + ln
        rc = np.full((len(tri), 2), -1, np.int32)
        for t in range(len(tri)):
            r = row_of.get((int(ti[t]), int(tj[t])))
            if r is not None:
                rc[t] = (r, tl[t] - s0)
        passes.append((lij, rc))
    return NT, MT, layout, passes


def shell_reach(shell, step):
    """Largest |k_a| of any mode of a shell: irk = int(|k|/step + 0.5) == shell  =>  |k| < step*(shell + 1/2) (py:376-378)."""
    return np.floor(step * (np.asarray(shell) + 0.5)).astype(np.int64)


COARSE_GRIDS = (256, 320, 400)           # grids with compiled two-stage FFT plans that the shell stage may drop to


def coarse_levels(N, step, tri, sizes):
    """Split the triangle list over transform grids.  sum_x I_i I_j I_l / Ngrid^3 = sum over closed mode triples
    (k1 + k2 + k3 = 0 mod Ngrid) of delta delta delta: as long as R_i + R_j + R_l < Nc no triple can close through a wrap on a
    grid of Nc or more points per side, so the sum -- and the triangle count -- is the SAME number on every such grid, and the
    band-limited shell fields (2 R < Nc) are the same functions sampled at fewer points.  A triangle goes to the coarsest of
    `sizes` (ascending, all < N) that holds it; the rest (at Ngrid=360, Nmax=40 those whose aliases the reference includes) stay
    on N.  Returns [(Nc, index array into tri, largest shell used)], empty levels dropped."""
    tri = np.asarray(tri)
    R = shell_reach(tri, step)
    rs, rmax = R.sum(axis=1), R.max(axis=1)
    left = np.ones(len(tri), dtype=bool)
    out = []
    for Nc in sorted(int(x) for x in sizes):
        if Nc >= N:
            continue
        take = left & (rs < Nc) & (2 * rmax < Nc)
        if take.any():
            out.append((Nc, np.nonzero(take)[0], int(tri[take].max())))
            left &= ~take
    if left.any():
        out.append((int(N), np.nonzero(left)[0], int(tri[left].max())))
    return out


def tc_plan_shape(tri, s0, smax, layout=1):
    """(layout, MT, NT, passes) `build_tc_plan` would choose for a triangle subset, without building the plan."""
    S = smax - s0 + 1
    NT = (S + 15) // 16 * 16
    MT_max = 4 if NT <= 64 else 256 // NT
    pairs = {(int(a), int(b)) for a, b, _ in np.asarray(tri)[:, :3]}
    nj = {}
    for i, j in pairs:
        nj.setdefault(i, set()).add(j)
    npass = lambda m: (sum((len(v) + m - 1) // m for v in nj.values()) + 127) // 128
    if layout == 1 and MT_max == 4 and npass(1) > 1:
        iv = sorted(nj)
        lanes = sum((len(nj[iv[a]] | (nj[iv[a + 1]] if a + 1 < len(iv) else set())) + 1) // 2 for a in range(0, len(iv), 2))
        return 1, 4, NT, (lanes + 127) // 128
    MT = MT_max
    while MT > 1 and npass(MT - 1) == npass(MT_max):
        MT -= 1
    return 0, MT, NT, npass(MT)


def level_cost(N, step, s0, tri, sizes):
    """Modelled device seconds of the shell + triangle stage for a set of coarse grids (constants measured on B200,
    profiles/r2_summary.md): K5 ~ 6e-12 s per (shell pair x cell); K6 per pass and cell 2.1e-10 s in the 2x2 layout,
    (1.6 + 0.29 MT) e-10 s with MT tiles of plain rows, +2 % per column beyond 48."""
    cost = 0.
    for Nc, idx, smax in coarse_levels(N, step, tri, sizes):
        cells = float(Nc) ** 3
        cost += 6e-12 * ((smax - s0 + 2) // 2) * cells
        layout, MT, NT, npass = tc_plan_shape(tri[idx], s0, smax)
        per_cell = 2.1e-10 if layout == 1 else (1.6e-10 + 0.29e-10 * MT)
        cost += npass * cells * per_cell * (1. + 0.02 * max(NT - 48, 0))
    return cost


def default_level_sizes(N, step, tri, s0=None):
    """Which coarse grids the shell stage uses by default (PSB_BK_LEVELS overrides: 'off', or a comma list of grid sizes):
    the subset of COARSE_GRIDS below N with the smallest modelled cost (`level_cost`).  At the reference's Ngrid=360 that is none
    (the extra shell transforms cost more than the triangle stage saves); at 512 with 80 shells it is {256, 400}; at 1024 with 40
    shells {400} -- there the fine grid is not needed at all."""
    import itertools
    import os
    spec = os.environ.get('PSB_BK_LEVELS', 'auto').strip().lower()
    if spec in ('off', 'none', '0'):
        return []
    if spec != 'auto':
        return [int(x) for x in spec.split(',') if x.strip()]
    tri = np.asarray(tri)
    if s0 is None:
        s0 = int(tri.min())
    cand = [c for c in COARSE_GRIDS if c < N]
    best, best_cost = [], level_cost(N, step, s0, tri, [])
    for r in range(1, len(cand) + 1):
        for sub in itertools.combinations(cand, r):
            c = level_cost(N, step, s0, tri, sub)
            if c < 0.97 * best_cost:
                best, best_cost = list(sub), c
    return best


class PeriodicPipeline(object):
    """Device-resident pipeline for one (Ngrid) on the current CUDA device.  Holds the host-built tables
    (twiddles, fcomb phase/window tables, shell / bin index tables) and caches per-configuration data
    (mode counts, triangle tiles, exact triangle counts)."""

    _cache = {}

    @classmethod
    def get(cls, Ngrid):
        dev = _device()
        key = (dev.index, int(Ngrid))
        if key not in cls._cache:
            cls._cache[key] = cls(int(Ngrid), dev)
        return cls._cache[key]

    def __init__(self, Ngrid, dev):
        if Ngrid < 4 or Ngrid % 2:
            raise ValueError('Ngrid must be even')
        self.N, self.h, self.dev = Ngrid, Ngrid // 2, dev
        self.L = _lib.lib()
        N, h = self.N, self.h
        tw = np.empty(2 * N, np.float32)
        check(self.L.psb_twiddles_f32(N, _np_ptr(tw)), 'psb_twiddles_f32')
        rec = np.empty(2 * (h + 1), np.float64)
        wk = np.empty(h + 1, np.float32)
        check(self.L.psb_fcomb_tables(N, _np_ptr(rec), _np_ptr(wk)), 'psb_fcomb_tables')
        self.tw32 = torch.from_numpy(tw).to(dev)
        self.rec = torch.from_numpy(rec).to(dev)
        self.wk = torch.from_numpy(wk).to(dev)
        self._tw64 = None
        self.mmax = 3 * h * h
        self._irk, self._bins, self._nk, self._tiles, self._counts = {}, {}, {}, {}, {}
        self._pin_pool = {}                               # free pinned result buffers by size (bispectrum_launch/finish)
        self._copy_stream = None                          # upload stream of the *_many generators
        self._side_stream = None                          # second stream of the routed shell stage (multi-GPU)
        self._stage = {}                                  # pinned staging buffers + copy threads of the streamed upload

    # ------------------------------------------------------------------ tables
    @property
    def tw64(self):
        if self._tw64 is None:
            tw = np.empty(2 * self.N, np.float64)
            check(self.L.psb_twiddles_f64(self.N, _np_ptr(tw)), 'psb_twiddles_f64')
            self._tw64 = torch.from_numpy(tw).to(self.dev)
        return self._tw64

    def irk_table(self, step):
        """irk = int(|k|/step + 0.5) as a function of m = |k|^2: the reference's float64 expression
        (pyspectrum.py:376-378) evaluated on the host, so shell membership is bit-exact."""
        if step not in self._irk:
            m = np.arange(self.mmax + 1)
            irk = (np.sqrt(m) / step + 0.5).astype(int)
            self._irk[step] = torch.from_numpy(irk.astype(np.uint16)).to(self.dev)
        return self._irk[step]

    def pk_bin_table(self, Lbox, kf=None):
        """Pk_periodic's bin index (pyspectrum.py:696-703; same expression at py:560) as a function of m."""
        if kf is None:
            kf = 2 * np.pi / float(Lbox)
        key = ('pk', float(kf))
        if key not in self._bins:
            Nbins = self.N // 2
            phys_nyq = kf * float(self.N) / 2.
            rk = kf * np.sqrt(np.arange(self.mmax + 1))
            irk = (Nbins * rk / phys_nyq + 0.5).astype(int)
            self._bins[key] = torch.from_numpy(np.minimum(irk, 65535).astype(np.uint16)).to(self.dev)
        return self._bins[key]

    def rsd_bin_table(self, Nbins):
        key = ('rsd', int(Nbins))
        if key not in self._bins:
            t = np.empty(self.mmax + 1, np.uint16)
            check(self.L.psb_rsd_bin_table(self.N, int(Nbins), self.mmax, _np_ptr(t)), 'psb_rsd_bin_table')
            self._bins[key] = torch.from_numpy(t).to(self.dev)
        return self._bins[key]

    # ------------------------------------------------------------------ K1..K3
    STAGE_BYTES = 48 << 20                           # pinned staging buffer size of `upload`

    def upload(self, t):
        """Host tensor -> device.  Large PAGEABLE arrays (numpy callers) go through pinned staging buffers filled by worker threads
        under the DMA of the previous piece (the driver's own staged copy runs at ~9 GB/s); everything else is a plain async copy."""
        if t.is_cuda:
            return t
        nbytes = t.numel() * t.element_size()
        if t.is_pinned() or nbytes < 2 * self.STAGE_BYTES or not t.is_contiguous() or os.environ.get('PSB_HOST_STAGING', '1') == '0':
            return t.to(self.dev, non_blocking=True)
        st = self._staging_flat()
        out = torch.empty(t.shape, dtype=t.dtype, device=self.dev)
        src = t.view(-1).view(torch.uint8).numpy()
        dst = out.view(-1).view(torch.uint8)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        cs, main = self._copy_stream, torch.cuda.current_stream(self.dev)
        cs.wait_stream(main)
        nthr = st['nthr']
        for k, a in enumerate(range(0, nbytes, self.STAGE_BYTES)):
            b = min(a + self.STAGE_BYTES, nbytes)
            slot = st['slots'][k % len(st['slots'])]
            if slot['ev'] is not None:
                slot['ev'].synchronize()
            cuts = [a + (b - a) * i // nthr for i in range(nthr + 1)]
            list(st['pool'].map(lambda c: np.copyto(slot['np'][c[0] - a:c[1] - a], src[c[0]:c[1]]), zip(cuts, cuts[1:])))
            with torch.cuda.stream(cs):
                dst[a:b].copy_(slot['t'][:b - a], non_blocking=True)
                slot['ev'] = torch.cuda.Event()
                slot['ev'].record(cs)
        main.wait_stream(cs)
        return out

    def _staging_flat(self):
        st = self._stage.get('flat')
        if st is None:
            from concurrent.futures import ThreadPoolExecutor
            nthr = _copy_threads()
            slots = []
            for _ in range(3):
                t = torch.empty(self.STAGE_BYTES, dtype=torch.uint8, pin_memory=True)
                slots.append({'t': t, 'np': t.numpy(), 'ev': None})
            st = {'slots': slots, 'pool': ThreadPoolExecutor(max_workers=nthr), 'nthr': nthr}
            self._stage['flat'] = st
        return st

    def to_device(self, xyz, w=None):
        """Positions (3xN, numpy or torch, float32/float64) and weights -> device tensors + layout flags."""
        if isinstance(xyz, torch.Tensor):
            pos = xyz
            if pos.dtype not in (torch.float32, torch.float64):
                pos = pos.double()
            if pos.shape[0] != 3:
                raise ValueError('xyz must be 3 x N')
            aos = 0
            if not pos.is_contiguous():
                if pos.t().is_contiguous():
                    aos = 1
                    pos = pos.t()                        # (N,3) contiguous, like the numpy branch
                else:
                    pos = pos.contiguous()
            pos = self.upload(pos)
        else:
            xyz = np.asarray(xyz)
            if xyz.ndim != 2 or xyz.shape[0] != 3:
                raise ValueError('xyz must be 3 x N')
            if xyz.dtype not in (np.float32, np.float64):
                xyz = xyz.astype(np.float64)
            aos = 0
            if xyz.flags.f_contiguous and not xyz.flags.c_contiguous:
                aos = 1
                pos = self.upload(torch.from_numpy(xyz.T))                           # (N,3) contiguous
            else:
                pos = self.upload(torch.from_numpy(np.ascontiguousarray(xyz)))
        wt = None
        if w is not None:
            if isinstance(w, torch.Tensor):
                wt = w if w.dtype in (torch.float32, torch.float64) else w.double()
                wt = self.upload(wt.contiguous())
            else:
                w = np.asarray(w)
                if w.dtype not in (np.float32, np.float64):
                    w = w.astype(np.float64)
                wt = self.upload(torch.from_numpy(np.ascontiguousarray(w)))
        return pos, aos, wt

    def assign(self, pos, aos, wt, Lbox, offset=0., clip=True):
        """K1: pyspectrum.py:931-951 + estimator.f:284-512.  Returns (mesh [N,N,N,2] float32, sumw tensor)."""
        N = self.N
        Np = pos.shape[0] if aos else pos.shape[1]
        kf_ks = np.float32(float(N) / Lbox)
        mesh = torch.empty((N, N, N, 2), dtype=torch.float32, device=self.dev)
        wsb = self.L.psb_assign_workspace_bytes(Np, N)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        sumw = torch.empty(1, dtype=torch.float64, device=self.dev)
        check(self.L.psb_assign_pcs_interlaced(_ptr(pos), int(pos.dtype == torch.float64), aos, _ptr(wt),
                                               int(wt is not None and wt.dtype == torch.float64), Np, N,
                                               float(Lbox) if clip else 0.0, kf_ks, np.float32(offset),
                                               _ptr(mesh), 1, _ptr(ws), wsb, _ptr(sumw), _stream()),
              'psb_assign_pcs_interlaced')
        return mesh, sumw

    def mesh_to_delta(self, mesh, sumw, periodic=1):
        """K2+K3: pyspectrum.py:953-959.  Returns the half field as float32 [N,N,N/2+1,2] (kz,ky,kx)."""
        N, h = self.N, self.h
        half = torch.empty((N, N, h + 1, 2), dtype=torch.float32, device=self.dev)
        check(self.L.psb_fft_mesh_to_delta(_ptr(mesh), _ptr(half), N, _ptr(self.tw32), _ptr(self.rec), _ptr(self.wk),
                                           _ptr(sumw), periodic, _stream()), 'psb_fft_mesh_to_delta')
        return half

    CHUNK = 1 << 21                                  # particles per upload chunk of the streamed assignment

    def fft_periodic(self, xyz, w, Lbox):
        """FFT_periodic (py:909-959) -> (device half field, sum of weights).  Catalogues that still live on the host are uploaded in
        chunks on a copy stream while K1 assigns the previous chunk (the mesh accumulates): the drop-in call then costs the
        upload plus one chunk of K1 instead of upload + the whole K1."""
        host = (not isinstance(xyz, torch.Tensor)) or (not xyz.is_cuda)
        n = int(xyz.shape[1]) if hasattr(xyz, 'shape') and len(xyz.shape) == 2 else 0
        if host and n >= 2 * self.CHUNK and os.environ.get('PSB_STREAMED_ASSIGN', '1') != '0':
            with _Range('K1 assign (streamed upload)'):
                mesh, sumw = self.assign_streamed(xyz, w, Lbox)
        else:
            pos, aos, wt = self.to_device(xyz, w)
            with _Range('K1 assign'):
                mesh, sumw = self.assign(pos, aos, wt, Lbox)
        with _Range('K2+K3 fft+fcomb'):
            half = self.mesh_to_delta(mesh, sumw)
        return half, sumw

    def assign_streamed(self, xyz, w, Lbox):
        """K1 over a HOST catalogue (numpy or CPU torch, 3 x N) in chunks: upload of chunk k+1 (copy stream) under K1 of chunk k."""
        N = self.N
        xt = xyz if isinstance(xyz, torch.Tensor) else torch.from_numpy(np.asarray(xyz))
        if xt.dtype not in (torch.float32, torch.float64):
            xt = xt.double()
        if xt.dim() != 2 or xt.shape[0] != 3:
            raise ValueError('xyz must be 3 x N')
        aos = 0
        if not xt.is_contiguous():
            if xt.t().is_contiguous():
                aos = 1                                  # Fortran-ordered (3,N): rows of xt.t() are particles
            else:
                xt = xt.contiguous()
        wt_h = None
        if w is not None:
            wt_h = w if isinstance(w, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(w))
            if wt_h.dtype not in (torch.float32, torch.float64):
                wt_h = wt_h.double()
            wt_h = wt_h.contiguous()
        Np = int(xt.shape[1])
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        cs, main = self._copy_stream, torch.cuda.current_stream(self.dev)
        mesh = torch.empty((N, N, N, 2), dtype=torch.float32, device=self.dev)
        wsb = self.L.psb_assign_workspace_bytes(self.CHUNK, N)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        sumw = torch.zeros(1, dtype=torch.float64, device=self.dev)
        part = torch.empty(1, dtype=torch.float64, device=self.dev)
        kf_ks = np.float32(float(N) / Lbox)
        cs.wait_stream(main)
        bounds = [(a, min(a + self.CHUNK, Np)) for a in range(0, Np, self.CHUNK)]
        # pageable host memory (what a numpy caller passes): the driver's own staged copy runs at ~9 GB/s (C2: 42 ms per call against
        # 22 ms from pinned memory).  Worker threads copy each chunk into pinned staging buffers instead (numpy releases the GIL), the
        # DMA of chunk k and K1 of chunk k-1 run under the host copy of chunk k+1.
        staged_copy = not xt.is_pinned() and os.environ.get('PSB_HOST_STAGING', '1') != '0'
        stage = self._staging(xt.dtype, aos, None if wt_h is None else wt_h.dtype) if staged_copy else None
        src_np = xt.numpy() if staged_copy else None
        w_np = wt_h.numpy() if (staged_copy and wt_h is not None) else None
        for k, (a, b) in enumerate(bounds):
            n = b - a
            if staged_copy:
                slot = stage['slots'][k % len(stage['slots'])]
                if slot['ev'] is not None:
                    slot['ev'].synchronize()             # the DMA that last read this staging buffer is done
                jobs = []
                nsp = stage['nsplit']
                cuts = [n * i // nsp for i in range(nsp + 1)]
                if aos:
                    dst = slot['pos_np'][:n]
                    src = src_np.T[a:b]
                    jobs += [(dst[lo:hi], src[lo:hi]) for lo, hi in zip(cuts, cuts[1:])]
                else:
                    for c in range(3):
                        jobs += [(slot['pos_np'][c, lo:hi], src_np[c, a + lo:a + hi]) for lo, hi in zip(cuts, cuts[1:])]
                if w_np is not None:
                    jobs.append((slot['w_np'][:n], w_np[a:b]))
                list(stage['pool'].map(lambda j: np.copyto(j[0], j[1]), jobs))
                hp = slot['pos'][:n] if aos else slot['pos'][:, :n]
                hw = slot['w'][:n] if w_np is not None else None
            else:
                hp = xt.t()[a:b] if aos else xt[:, a:b]
                hw = wt_h[a:b] if wt_h is not None else None
            with torch.cuda.stream(cs):
                if aos:
                    pc = hp.to(self.dev, non_blocking=True)
                else:
                    pc = torch.empty((3, n), dtype=xt.dtype, device=self.dev)
                    if hp.is_contiguous():
                        pc.copy_(hp, non_blocking=True)
                    else:                                # a column range of a wider array: one copy per coordinate row
                        for c in range(3):
                            pc[c].copy_(hp[c], non_blocking=True)
                wc = hw.to(self.dev, non_blocking=True) if hw is not None else None
                ev = torch.cuda.Event()
                ev.record(cs)
            if staged_copy:
                slot['ev'] = ev
            main.wait_event(ev)
            pc.record_stream(main)
            if wc is not None:
                wc.record_stream(main)
            check(self.L.psb_assign_pcs_interlaced(_ptr(pc), int(pc.dtype == torch.float64), aos, _ptr(wc),
                                                   int(wc is not None and wc.dtype == torch.float64), n, N,
                                                   float(Lbox), kf_ks, np.float32(0.), _ptr(mesh), 1 if k == 0 else 0, _ptr(ws), wsb,
                                                   _ptr(part), _stream()), 'psb_assign_pcs_interlaced')
            sumw += part
        return mesh, sumw

    def _staging(self, dtype, aos, wdtype):
        """Pinned staging buffers (three chunks deep) + the copy thread pool of the streamed upload, created once per layout."""
        key = (dtype, aos, wdtype)
        st = self._stage.get(key)
        if st is None:
            from concurrent.futures import ThreadPoolExecutor
            nthr = _copy_threads()
            slots = []
            for _ in range(3):
                pos = torch.empty((self.CHUNK, 3) if aos else (3, self.CHUNK), dtype=dtype, pin_memory=True)
                w = torch.empty(self.CHUNK, dtype=wdtype, pin_memory=True) if wdtype is not None else None
                slots.append({'pos': pos, 'pos_np': pos.numpy(), 'w': w, 'w_np': None if w is None else w.numpy(), 'ev': None})
            st = {'slots': slots, 'pool': ThreadPoolExecutor(max_workers=nthr), 'nsplit': max(1, nthr // (1 if aos else 3) or 1)}
            self._stage[key] = st
        return st

    @staticmethod
    def survey_distance_table(cosmo, zmax, nnodes=4097):
        """Cubic Hermite nodes {D(z_k), D'(z_k) dz} of the line-of-sight comoving distance in Mpc/h over [0, zmax] for the
        caller's cosmology object (anything with `comoving_distance(z)` [Mpc, plain array or astropy Quantity] and `h`).
        The slopes come from a not-a-knot cubic spline through the nodes (error ~ dz^3: far below float32 positions)."""
        from scipy.interpolate import CubicSpline
        zn = np.linspace(0., float(zmax), nnodes)
        D = cosmo.comoving_distance(zn)
        D = np.asarray(getattr(D, 'value', D), dtype=np.float64) * float(getattr(cosmo.h, 'value', cosmo.h))
        slope = CubicSpline(zn, D)(zn, 1) * (float(zmax) / (nnodes - 1))
        return np.ascontiguousarray(np.stack([D, slope], axis=1))

    def survey_prepare(self, radecz, nb, w, P0_fkp, cosmo):
        """Catalogue pre-step of the survey path on the device (psb_survey_prepare): returns (xyz float32 [3,Np] device,
        FKP weights float32 [Np] device, host float64[12] = Ntot, I12, I13, I22, I23, I33, min xyz, max xyz)."""
        radecz = np.ascontiguousarray(radecz, dtype=np.float64)
        if radecz.ndim != 2 or radecz.shape[0] != 3:
            raise ValueError('radecz has to be have shape [3,N]')
        Np = int(radecz.shape[1])
        zmin, zmax = float(radecz[2].min()), float(radecz[2].max())
        if not (zmin >= 0.) or not np.isfinite(zmax):
            raise ValueError('redshifts must be finite and >= 0')
        zmax = max(zmax, 1e-6)
        tab = torch.from_numpy(self.survey_distance_table(cosmo, zmax)).to(self.dev)
        rdz = self.upload(torch.from_numpy(radecz))
        nbd = self.upload(torch.from_numpy(np.ascontiguousarray(nb, dtype=np.float64)))
        if nbd.numel() != Np:
            raise ValueError('nbar must have one entry per object')
        wd = None
        if w is not None:
            wd = self.upload(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)))
            if wd.numel() != Np:
                raise ValueError('w must have one entry per object')
        xyz = torch.empty((3, Np), dtype=torch.float32, device=self.dev)
        wf = torch.empty(Np, dtype=torch.float32, device=self.dev)
        out = torch.empty(12, dtype=torch.float64, device=self.dev)
        check(self.L.psb_survey_prepare(_ptr(rdz), _ptr(nbd), _ptr(wd), Np, _ptr(tab), tab.shape[0], zmax, float(P0_fkp),
                                        _ptr(xyz), _ptr(wf), _ptr(out), _stream()), 'psb_survey_prepare')
        return xyz, wf, out.cpu().numpy()

    def fft_survey(self, xyz, w, Lbox):
        """Survey-geometry delta_0(k) (pyspectrum.py:817-823): positions in (-L/2, L/2) shifted by half a box through the
        assignment offset (0.5*Ngrid), no clipping, and fcomb_survey (no division by sum w; estimator.f:686)."""
        pos, aos, wt = self.to_device(xyz, w)
        mesh, sumw = self.assign(pos, aos, wt, Lbox, offset=0.5 * self.N, clip=False)
        return self.mesh_to_delta(mesh, sumw, periodic=0)

    def half_from_full(self, delta):
        """Host full field (Ngrid,Ngrid,Ngrid) complex, indexed [kx,ky,kz] -> device half field of its Hermitian part
        d_h(k) = (d(k) + conj(d(-k)))/2.  The reference's shell stage keeps Re(FFT(masked field)) (pyspectrum.py:202-215,
        389-399), which is the transform of the Hermitian part, so this is exact for any input and the identity for the
        output of reflect_delta."""
        N, h = self.N, self.h
        delta = np.asarray(delta)
        if delta.shape != (N, N, N):
            raise ValueError('delta must be (Ngrid,Ngrid,Ngrid) for this pipeline')
        d = delta.astype(np.complex64, copy=False)
        idx = (-np.arange(N)) % N
        mirror = np.conj(d[idx[:h + 1]][:, idx][:, :, idx])
        dh = np.complex64(0.5) * (d[:h + 1] + mirror)
        arr = np.ascontiguousarray(dh.transpose(2, 1, 0))              # [kz][ky][kx]
        return torch.from_numpy(arr.view(np.float32).reshape(N, N, h + 1, 2)).to(self.dev)

    # ------------------------------------------------------------------ K4
    def pk_monopole(self, half, Lbox):
        Nbins = self.N // 2
        out = torch.empty(3 * Nbins, dtype=torch.float64, device=self.dev)
        kf = 2 * np.pi / float(Lbox)
        check(self.L.psb_pk_monopole(_ptr(half), self.N, _ptr(self.pk_bin_table(Lbox)), Nbins, kf, _ptr(out), _stream()),
              'psb_pk_monopole')
        return out

    def pk_multipoles(self, half, Lbox_int, rsd, Nmubin):
        Nbins = self.N // 2
        out = torch.empty((5 + 4 * Nmubin) * Nbins, dtype=torch.float64, device=self.dev)
        trig = np.empty(4, np.float32)
        check(self.L.psb_rsd_trig(int(rsd), _np_ptr(trig)), 'psb_rsd_trig')
        pi = np.float32(3.141592654)
        kf32 = np.float32(np.float32(2.) * pi) / np.float32(Lbox_int)              # estimator.f:164,169
        check(self.L.psb_pk_multipoles(_ptr(half), self.N, _ptr(self.rsd_bin_table(Nbins)), Nbins, int(Nmubin),
                                       kf32, _np_ptr(trig), _ptr(out), _stream()), 'psb_pk_multipoles')
        return out, kf32

    def pk_kmu_python(self, full, kf, rsd, Nmubin):
        """code='python' (k,mu) estimator (pyspectrum.py:545-626) on a device FULL field [kx,ky,kz,2] float32: raw float64 sums."""
        Nbins = self.N // 2
        out = torch.empty((5 + 4 * Nmubin) * Nbins, dtype=torch.float64, device=self.dev)
        theta_obs = [0.5 * np.pi, 0.5 * np.pi, 0.][rsd]                    # py:563-564
        phi_obs = [0., 0.5 * np.pi, 0.][rsd]
        trig = np.array([np.cos(theta_obs), np.sin(theta_obs), np.cos(phi_obs), np.sin(phi_obs)], dtype=np.float64)
        check(self.L.psb_pk_kmu_python(_ptr(full), self.N, _ptr(self.pk_bin_table(None, kf=kf)), Nbins, int(Nmubin), float(kf),
                                       _np_ptr(trig), _ptr(out), _stream()), 'psb_pk_kmu_python')
        return out

    # ------------------------------------------------------------------ K5
    def shell_mode_counts(self, step, Nmax):
        key = (step, Nmax)
        if key not in self._nk:
            nk = torch.empty(Nmax + 1, dtype=torch.int64, device=self.dev)
            check(self.L.psb_shell_mode_counts(self.N, _ptr(self.irk_table(step)), Nmax + 1, _ptr(nk), _stream()),
                  'psb_shell_mode_counts')
            self._nk[key] = nk.cpu().numpy()
        return self._nk[key]

    def shell_scales(self, half, step, s0, Nmax):
        """Exact power-of-two normalisation per shell (rms of the stored field ~ 2) from the Parseval shell power."""
        S = Nmax - s0 + 1
        S_alloc = S + (S % 2)
        st = _stream()
        pw = torch.empty(Nmax, dtype=torch.float64, device=self.dev)
        check(self.L.psb_bk_shell_power(_ptr(half), self.N, _ptr(self.irk_table(step)), Nmax, _ptr(pw), st), 'psb_bk_shell_power')
        scales = torch.ones(S_alloc, dtype=torch.float32, device=self.dev)
        check(self.L.psb_bk_shell_scales(ctypes.c_void_p(pw.data_ptr() + 8 * (s0 - 1)), S, np.float32(2.0), _ptr(scales), st),
              'psb_bk_shell_scales')
        return scales

    def shell_fields(self, half, step, s0, Nmax, dtype=torch.float32, scaled=False, pairs=None, scales=None, src=None, routed=None):
        """K5: shells s0..Nmax as real fields [S_alloc, N^3] + sum_x I_j^2 per shell.
        half=None -> delta == 1 (counts).  Two shells ride on one complex transform.
        scaled=True stores I_j * scale_j with scale_j an exact power of two putting the rms near 2 (from the
        Parseval shell power, or the `scales` tensor if given) and stored pre-split for the tensor-core triangle kernel:
        each aligned cell pair (x, x+1) holds the words {half2 hi(x,x+1), half2 lo(x,x+1)}, hi = fp16(v), lo = fp16(v - hi)
        (still 4 bytes per cell; `unpack_fields` gives the float32 values back).
        returns (fields, sumsq, scales, maxabs) -- sumsq and maxabs refer to the scaled values before the split.
        pairs: optional list of pair indices p (shells s0+2p, s0+2p+1) to compute -- the multi-GPU path shards shells this way;
        the returned fields then hold only those pairs, in the given order (2 rows per pair), sumsq/maxabs likewise.
        src: the pipeline (finer grid) `half` was measured on, if not this one: the shells are then transformed on THIS coarser
        grid -- the same band-limited fields at fewer points (needs 2*floor(step*(Nmax+1/2)) < N; see `coarse_levels`).
        routed: (device int64 tensor [len(pairs), 2, nranks], planes per rank, nranks) -- multi-GPU: the z pass writes every output
        plane straight into the slab buffer of the rank that owns it (psb_bk_shell_pair_f32_routed; the table holds, per local
        pair, shell of the pair and rank, the address the field would start at in that rank's buffer, see
        multigpu.SlabBuffers.route); no field tensor is allocated or returned (fields = None)."""
        N = self.N
        Ns = N if src is None else src.N
        if Ns != N and (Ns < N or 2 * int(np.floor(step * (Nmax + 0.5))) >= N):
            raise ValueError('a coarser shell grid must hold every shell without wrap-around')
        ncell = N * N * N
        S = Nmax - s0 + 1
        S_alloc = S + (S % 2)
        f64 = dtype == torch.float64
        plist = list(range(S_alloc // 2)) if pairs is None else list(pairs)
        nrow = 2 * len(plist)
        if routed is not None and (f64 or not scaled):
            raise ValueError('routed output exists for the scaled float32 fields only')
        fields = None if routed is not None else torch.empty((nrow, ncell), dtype=dtype, device=self.dev)
        sumsq = torch.zeros(nrow, dtype=torch.float64, device=self.dev)
        maxabs = None
        st = _stream()
        irk = self.irk_table(step)
        if scaled:
            assert not f64 and half is not None
            if scales is None:
                scales = (src or self).shell_scales(half, step, s0, Nmax)
            maxabs = torch.zeros(nrow, dtype=torch.int32, device=self.dev)
            if pairs is None:
                sc_local = scales
            else:                                        # one gather: rows of this rank's pairs (padding pairs -> scale 1)
                rows = [min(2 * q + e, S_alloc - 1) for q in plist for e in (0, 1)]
                sc_local = scales[torch.tensor(rows, dtype=torch.long, device=self.dev)].contiguous()
        Rmax = int(np.floor(step * (Nmax + 0.5)))
        W = min(2 * Rmax + 1, N)
        cdt = torch.float64 if f64 else torch.float32
        t1 = torch.empty(W * W * N * 2, dtype=cdt, device=self.dev)
        t2 = torch.empty(W * N * N * 2, dtype=cdt, device=self.dev)
        tw = self.tw64 if f64 else self.tw32
        # routed output: the z pass of a pair is bound by its NVLink stores, so consecutive pairs alternate between two streams (own
        # scratch each) and the x / y passes of pair k+1 run under the z pass of pair k
        nstream = int(os.environ.get('PSB_ROUTED_STREAMS', '2')) if (routed is not None and len(plist) > 1) else 1
        main = torch.cuda.current_stream(self.dev)
        lanes = [(main, t1, t2)]
        if nstream > 1:
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.dev)
            side = self._side_stream
            side.wait_stream(main)
            lanes.append((side, torch.empty_like(t1), torch.empty_like(t2)))
            for t in lanes[1][1:]:
                t.record_stream(side)
        for r, pidx in enumerate(plist):
            s = 2 * pidx
            if s >= S:                                   # padding pair (shard equalisation): empty shells -> zero fields
                if fields is not None:
                    fields[2 * r:2 * r + 2].zero_()
                continue
            sa = s0 + s
            sb = sa + 1 if s + 1 < S else -1
            R = int(np.floor(step * (max(sa, sb) + 0.5)))
            sq = ctypes.c_void_p(sumsq.data_ptr() + 8 * 2 * r)
            if f64:
                check(self.L.psb_bk_shell_pair_f64(_ptr(half), _ptr(irk), N, Ns, sa, sb, R, _ptr(t1), _ptr(t2), _ptr(fields[2 * r]),
                                                   _ptr(fields[2 * r + 1]), sq, _ptr(tw), st), 'psb_bk_shell_pair_f64')
            elif routed is not None:
                table, rplanes, rranks = routed
                ln_stream, ln_t1, ln_t2 = lanes[r % len(lanes)]
                check(self.L.psb_bk_shell_pair_f32_routed(_ptr(half), _ptr(irk), N, Ns, sa, sb, R, _ptr(ln_t1), _ptr(ln_t2),
                                                          ctypes.c_void_p(table.data_ptr() + 8 * 2 * rranks * r), rplanes, rranks, sq,
                                                          ctypes.c_void_p(sc_local.data_ptr() + 4 * 2 * r),
                                                          ctypes.c_void_p(maxabs.data_ptr() + 4 * 2 * r), 1, _ptr(tw),
                                                          ctypes.c_void_p(ln_stream.cuda_stream)),
                      'psb_bk_shell_pair_f32_routed')
            else:
                sc = mx = None
                if scaled:
                    sc = ctypes.c_void_p(sc_local.data_ptr() + 4 * 2 * r)
                    mx = ctypes.c_void_p(maxabs.data_ptr() + 4 * 2 * r)
                check(self.L.psb_bk_shell_pair_f32(_ptr(half), _ptr(irk), N, Ns, sa, sb, R, _ptr(t1), _ptr(t2), _ptr(fields[2 * r]),
                                                   _ptr(fields[2 * r + 1]), sq, sc, mx, 1 if scaled else 0, _ptr(tw), st),
                      'psb_bk_shell_pair_f32')
        if len(lanes) > 1:
            main.wait_stream(lanes[1][0])
        if scaled:
            if fields is not None:
                fields.psb_packed = True                 # rows hold {half2 hi, half2 lo} per cell pair (see unpack_fields)
            return fields, sumsq, sc_local, maxabs
        return fields, sumsq

    @staticmethod
    def unpack_fields(fields):
        """float32 values of packed shell fields (shell_fields(scaled=True)): value = hi + lo per cell."""
        nrow, ncell = fields.shape
        h = fields.view(torch.float16).view(nrow, ncell // 2, 2, 2)        # [pair][hi word | lo word][cell parity]
        return (h[:, :, 0, :].float() + h[:, :, 1, :].float()).reshape(nrow, ncell)

    # ------------------------------------------------------------------ K6
    @staticmethod
    def _tri_key(tri):
        return None if tri is None else (len(tri), hash(np.ascontiguousarray(tri, dtype=np.int32).tobytes()))

    def triangle_tiles(self, Nmax, Ncut, step, tri=None):
        """4x4x4 tile descriptors of the FFMA kernel for the loop-nest triangles (or the given subset), cached."""
        key = (Nmax, Ncut, step, self._tri_key(tri))
        if key not in self._tiles:
            if tri is None:
                tri = triangle_list(Nmax, Ncut, step)
            s0 = Ncut // step
            nt = ctypes.c_int(0)
            tri_c = np.ascontiguousarray(tri, dtype=np.int32)
            check(self.L.psb_bk_build_tiles(_np_ptr(tri_c), len(tri_c), s0, None, ctypes.byref(nt)), 'psb_bk_build_tiles')
            tiles = np.empty((nt.value, 68), np.int32)
            check(self.L.psb_bk_build_tiles(_np_ptr(tri_c), len(tri_c), s0, _np_ptr(tiles), ctypes.byref(nt)), 'psb_bk_build_tiles')
            self._tiles[key] = (tri_c, torch.from_numpy(tiles).to(self.dev), nt.value)
        return self._tiles[key]

    def triangle_sums(self, fields, Nmax, Ncut, step, engine='auto', field_rows=None, packed=None, tri=None):
        """K6: sum_x I_i I_j I_l for every triangle of the loop nest (float64 tensor, loop order).
        engine: 'tc' = tcgen05 split-fp16 kernel (needs the scaled + packed fields of shell_fields(scaled=True)),
                'fma' = FFMA/DFMA register-tile kernel (plain float32/float64 fields, or packed ones which it decodes),
                'auto' = tc when the fields are packed and the shapes allow it.
        packed: whether `fields` holds packed hi/lo halves (default: the tag shell_fields put on the tensor).
        tri: optional subset of triangles, int array [n,3] of shell indices (i,j,l) all <= Nmax (the largest shell held in
             `fields`); the sums come back in the order of `tri`."""
        S = Nmax - Ncut // step + 1
        rows = list(range(S)) if field_rows is None else list(field_rows)                  # row of `fields` holding shell slot f
        if packed is None:
            packed = bool(getattr(fields, 'psb_packed', False))
        tc_ok = packed and fields.dtype == torch.float32 and fields.shape[1] % 64 == 0 and S <= 128
        if engine == 'tc' and not tc_ok:
            raise ValueError('tensor-core triangle kernel needs packed float32 fields, N^3 % 64 == 0 and <= 128 shells')
        if engine == 'tc' or (engine == 'auto' and tc_ok):
            return self._triangle_sums_tc(fields, Nmax, Ncut, step, rows, tri)
        tri, tiles, ntiles = self.triangle_tiles(Nmax, Ncut, step, tri)
        nf = (S + 3) // 4 * 4
        ptrs = [fields[rows[min(f, S - 1)]].data_ptr() for f in range(nf)]
        dptr = torch.tensor(ptrs, dtype=torch.int64).to(self.dev)
        sums = torch.zeros(len(tri), dtype=torch.float64, device=self.dev)
        wsb = self.L.psb_bk_triangle_workspace_bytes(ntiles)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        if fields.dtype == torch.float64:
            rc = self.L.psb_bk_triangle_sums_f64(_ptr(dptr), nf, fields.shape[1], _ptr(tiles), ntiles, _ptr(sums), _ptr(ws), wsb, _stream())
        else:
            rc = self.L.psb_bk_triangle_sums_f32(_ptr(dptr), nf, fields.shape[1], _ptr(tiles), ntiles, _ptr(sums), _ptr(ws), wsb,
                                                 1 if packed else 0, _stream())
        check(rc, 'psb_bk_triangle_sums')
        return sums

    def tc_passes(self, Nmax, Ncut, step, layout=None, tri=None):
        """Device copy of the tensor-core plan (`build_tc_plan`) for the loop-nest triangles or a subset, cached."""
        if layout is None:
            layout = int(os.environ.get('PSB_TC_LAYOUT', '1'))       # 2x2 blocks: 17 % faster (profiles/r1_summary.md)
        MT = int(os.environ['PSB_TC_MT']) if os.environ.get('PSB_TC_MT') else None      # experiments only
        key = ('tc', Nmax, Ncut, step, layout, MT, self._tri_key(tri))
        if key not in self._tiles:
            if tri is None:
                tri = triangle_list(Nmax, Ncut, step)
            tri = np.ascontiguousarray(tri, dtype=np.int32)
            NT, MT, layout_used, passes = build_tc_plan(tri, Ncut // step, Nmax, layout, MT)
            dev_passes = [(torch.from_numpy(lij).to(self.dev), MT, torch.from_numpy(rc).to(self.dev)) for lij, rc in passes]
            self._tiles[key] = (tri, NT, dev_passes, layout_used)
        return self._tiles[key]

    def _triangle_sums_tc(self, fields, Nmax, Ncut, step, rows, tri=None):
        tri, NT, passes, layout = self.tc_passes(Nmax, Ncut, step, tri=tri)
        S = Nmax - Ncut // step + 1
        dptr = torch.tensor([fields[rows[f]].data_ptr() for f in range(S)], dtype=torch.int64).to(self.dev)
        sums = torch.zeros(len(tri), dtype=torch.float64, device=self.dev)
        MT = passes[0][1]
        wsb = self.L.psb_bk_triangle_tc_workspace_bytes(MT, NT)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        for lij, MT, rc in passes:
            check(self.L.psb_bk_triangle_sums_tc(_ptr(dptr), S, fields.shape[1], _ptr(lij), layout, MT, NT, _ptr(rc), len(tri),
                                                 _ptr(sums), _ptr(ws), wsb, _stream()), 'psb_bk_triangle_sums_tc')
        return sums

    def _pinned(self, n):
        """float64 pinned host buffer of n elements from a small free list (cudaHostAlloc per call would cost more than K4)."""
        free = self._pin_pool.setdefault(n, [])
        return free.pop() if free else torch.empty(n, dtype=torch.float64, pin_memory=True)

    def bk_levels(self, step, Ncut, Nmax):
        """(triangle list, [(pipeline of the transform grid, triangle indices (numpy), the same on the device, largest shell)])
        for this configuration: `coarse_levels` with the default / PSB_BK_LEVELS grid sizes, cached."""
        key = ('levels', step, Ncut, Nmax, os.environ.get('PSB_BK_LEVELS', 'auto'))
        if key not in self._tiles:
            tri = triangle_list(Nmax, Ncut, step)
            lev = []
            for Nc, idx, smax in coarse_levels(self.N, step, tri, default_level_sizes(self.N, step, tri, Ncut // step)):
                pc = self if Nc == self.N else PeriodicPipeline.get(Nc)
                lev.append((pc, idx, torch.from_numpy(idx).to(self.dev), smax))
            self._tiles[key] = (tri, lev)
        return self._tiles[key]

    def bispectrum_launch(self, half, step, Ncut, Nmax, engine='auto', sumw=None, timers=None):
        """Enqueue K5 + K6 for one catalogue and the device->host copy of everything the host epilogue needs (triangle sums,
        shell powers, scales, max|I|, sum of weights) into pinned memory.  Returns a handle for `bispectrum_finish`; nothing
        here waits for the GPU, so the caller can enqueue the next catalogue before collecting this one.
        The triangles are split over transform grids (`bk_levels`): each level transforms the shells it needs on its own grid
        straight from `half` and sums its triangles there; sums and shell powers are brought to the units of this grid
        (x (N/Nc)^3).  timers: optional list receiving (stage, start event, end event) for bench.py."""
        s0 = Ncut // step
        S = Nmax - s0 + 1
        SA = S + (S % 2)
        tri, levels = self.bk_levels(step, Ncut, Nmax)
        scales = self.shell_scales(half, step, s0, Nmax)
        sums = torch.zeros(len(tri), dtype=torch.float64, device=self.dev)
        sumsq = torch.zeros(SA, dtype=torch.float64, device=self.dev)
        maxabs = torch.zeros(SA, dtype=torch.float32, device=self.dev)
        have, all_tc = 0, True

        def mark():
            if timers is None:
                return None
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        for pc, idx, idx_dev, smax in levels:
            Sl = smax - s0 + 1
            vol = (float(self.N) / pc.N) ** 3
            t0 = mark()
            with _Range('K5 shell fields %d^3' % pc.N):
                fields, sq, _, mx = pc.shell_fields(half, step, s0, smax, scaled=True, scales=scales, src=None if pc is self else self)
            t1 = mark()
            use_tc = engine in ('auto', 'tc') and fields.shape[1] % 64 == 0 and Sl <= 128
            all_tc = all_tc and use_tc
            with _Range('K6 triangle sums %d^3' % pc.N):
                sl = pc.triangle_sums(fields, smax, Ncut, step, engine='tc' if use_tc else 'fma', tri=tri[idx])
            t2 = mark()
            if timers is not None:
                timers.append(('shell_fields', t0, t1))
                timers.append(('triangles', t1, t2))
            del fields
            sums.index_copy_(0, idx_dev, sl if vol == 1.0 else sl * vol)
            if Sl > have:                                   # shell powers from the coarsest level that holds the shell (Parseval)
                sumsq[have:Sl] = sq[have:Sl] if vol == 1.0 else sq[have:Sl] * vol
                have = Sl
            maxabs[:Sl] = torch.maximum(maxabs[:Sl], mx.view(torch.float32)[:Sl])
        sw = sumw.double().reshape(1) if sumw is not None else torch.zeros(1, dtype=torch.float64, device=self.dev)
        dev = torch.cat([sums, sumsq, scales.double(), maxabs.double(), sw])
        host = self._pinned(dev.numel())
        host.copy_(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return {'host': host, 'ev': ev, 'half': half, 'nt': sums.numel(), 'SA': SA, 'use_tc': all_tc and engine != 'fma',
                'args': (Nmax, Ncut, step)}

    def bispectrum_finish(self, h):
        """Wait for a `bispectrum_launch` and return host arrays (sum_x I_i I_j I_l per triangle, sum_x I_j^2 per shell,
        sum of weights) in the reference's units.  The tensor-core path works on power-of-two scaled fields; if a pair
        product could leave the fp16 range (max|I_i| max|I_j| >= 4e4 after scaling -- only for pathologically concentrated
        catalogues) the shell and triangle stage is redone with the FFMA kernel."""
        h['ev'].synchronize()
        host = h['host'].numpy().copy()
        self._pin_pool[h['host'].numel()].append(h['host'])
        Nmax, Ncut, step = h['args']
        s0 = Ncut // step
        nt, SA = h['nt'], h['SA']
        sums_h, sumsq_h, sc, mx = host[:nt], host[nt:nt + SA], host[nt + SA:nt + 2 * SA], host[nt + 2 * SA:nt + 3 * SA]
        if h['use_tc'] and mx.max() ** 2 >= 4.0e4:
            redo = self.bispectrum_launch(h['half'], step, Ncut, Nmax, engine='fma')
            h['half'] = None
            return self.bispectrum_finish(redo)[:2] + (float(host[-1]),)
        h['half'] = None
        tri = triangle_list(Nmax, Ncut, step)
        sums_h = sums_h / (sc[tri[:, 0] - s0] * sc[tri[:, 1] - s0] * sc[tri[:, 2] - s0])
        sumsq_h = sumsq_h / sc ** 2
        return sums_h, sumsq_h, float(host[-1])

    def bispectrum_sums(self, half, step, Ncut, Nmax, engine='auto'):
        """K5 + K6 for one catalogue with one device->host read: (triangle sums, shell powers) as host arrays."""
        sums_h, sumsq_h, _ = self.bispectrum_finish(self.bispectrum_launch(half, step, Ncut, Nmax, engine))
        return sums_h, sumsq_h

    # ------------------------------------------------------------------ counts
    def counts(self, Nmax, Ncut, step, fft='pyfftw', silent=True):
        """Raw triangle counts array (Nmax,Nmax,Nmax) float64 = N^3 * (#closed triangles), as the reference's
        cache files hold (pyspectrum.py:962-1030).  Memory cache -> file in dat_dir() -> GPU float64 compute."""
        key = (Nmax, Ncut, step)
        if key in self._counts:
            return self._counts[key]
        N = self.N
        fcnt = ''.join(['counts', '.Ngrid', str(N), '.Nmax', str(Nmax), '.Ncut', str(Ncut), '.step', str(step), '.', fft])
        f_counts = os.path.join(dat_dir(), fcnt)
        counts = None
        if os.path.isfile(f_counts):
            try:
                counts = _read_fortran_record(f_counts, Nmax)
            except ValueError as exc:                    # truncated / foreign file: recompute and replace it
                if not silent:
                    print('--- ignoring %s (%s) ---' % (f_counts, exc))
        if counts is None:
            if not silent:
                print('--- calculating %s ---' % f_counts)
            counts = self.compute_counts(Nmax, Ncut, step)
            try:
                os.makedirs(dat_dir(), exist_ok=True)
                _write_fortran_record(f_counts, counts)
            except OSError:
                pass
        self._counts[key] = counts
        return counts

    def compute_counts(self, Nmax, Ncut, step):
        """Exact triangle counts N^3 * #{closed mode triples} in float64 on the GPU (delta == 1 through K5 + K6, py:975-1023).
        The number of closed triples of a triangle is the same on every grid that holds it without wrap, so each level of
        `bk_levels` counts its triangles on its own (coarser) grid."""
        N = self.N
        s0 = Ncut // step
        tri, levels = self.bk_levels(step, Ncut, Nmax)
        counts = np.zeros((Nmax, Nmax, Nmax), dtype=np.float64)
        for pc, idx, _, smax in levels:
            fields, _ = pc.shell_fields(None, step, s0, smax, dtype=torch.float64)
            sums = pc.triangle_sums(fields, smax, Ncut, step, tri=tri[idx]).cpu().numpy()
            del fields
            n3 = float(pc.N) ** 3
            nint = np.rint(sums / n3)
            if np.abs(sums / n3 - nint).max() > 1e-3:
                raise RuntimeError('triangle counts did not come out as integers (max dev %g)' % np.abs(sums / n3 - nint).max())
            t = tri[idx]
            counts[t[:, 0] - 1, t[:, 1] - 1, t[:, 2] - 1] = nint * float(N) ** 3
        return counts


def _read_fortran_record(fname, Nmax):
    """One Fortran sequential record of Nmax^3 float64 (scipy.io.FortranFile.read_reals, py:970-972)."""
    from .util import fortran_records
    recs = fortran_records(fname)
    if len(recs) != 1 or len(recs[0]) != 8 * Nmax ** 3:
        raise ValueError('%s: expected one record of %d float64' % (fname, Nmax ** 3))
    return np.frombuffer(recs[0], '<f8').reshape(Nmax, Nmax, Nmax).copy()


def _write_fortran_record(fname, counts):
    b = np.ascontiguousarray(counts, dtype='<f8').tobytes()
    mark = np.array([len(b)], dtype='<i4').tobytes()
    tmp = '%s.tmp%d' % (fname, os.getpid())          # atomic: several ranks may fill the same cache concurrently
    with open(tmp, 'wb') as f:
        f.write(mark + b + mark)
    os.replace(tmp, fname)


# ----------------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------------
def _npart(xyz):
    return int(xyz.shape[1])


def _sum_w(w, N, sumw_dev):
    """np.sum(w) of pyspectrum.py:323/679 (host value when the caller's weights live on the host)."""
    if w is None:
        return float(N)
    if isinstance(w, torch.Tensor):
        return float(sumw_dev.item())
    return float(np.sum(w))


def FFT_periodic(xyz, w=None, Lbox=2600., Ngrid=360, fft='pyfftw', silent=True):
    """pyspectrum.py:909-959: returns delta(k) on the half grid, complex64, shape (Ngrid//2+1, Ngrid, Ngrid)
    indexed [kx,ky,kz] (Fortran-contiguous view, exactly the reference's memory layout)."""
    pipe = PeriodicPipeline.get(Ngrid)
    half, _ = pipe.fft_periodic(xyz, w, Lbox)
    if not silent:
        print('position grid FFTed')
        print('fcomb complete')
    a = half.cpu().numpy().view(np.complex64)[..., 0]        # (kz,ky,kx) C-order
    return a.transpose(2, 1, 0)


def reflect_delta(delt, Ngrid=360, silent=True):
    """pyspectrum.py:1134-1157: half field -> full Hermitian complex64 field (host utility; the GPU
    kernels work from the half field and never materialise this)."""
    if not silent:
        print('reflecting the half field')
    h = Ngrid // 2
    idx = (-np.arange(Ngrid)) % Ngrid
    delta = np.zeros((Ngrid, Ngrid, Ngrid), dtype=np.complex64)
    delta[:h + 1] = delt
    mirror = np.conj(delt[1:h])[:, idx][:, :, idx]            # value at (i,-j,-k) of rows 1..h-1
    delta[Ngrid - 1:h:-1] = mirror
    for p in [(h, 0, 0), (0, h, 0), (0, 0, h), (0, h, h), (h, 0, h), (h, h, 0), (h, h, h)]:
        delta[p] = np.real(delt[p])
    return delta


def _counts_Bk123(Ngrid=360, Nmax=40, Ncut=3, step=3, fft='pyfftw', silent=True):
    """pyspectrum.py:962-1030."""
    return PeriodicPipeline.get(Ngrid).counts(Nmax, Ncut, step, fft=fft, silent=silent)


def Pk_periodic(xyz, w=None, Lbox=2600, Ngrid=360, fft='pyfftw', silent=True):
    """Power-spectrum monopole of a periodic box; see pyspectrum.py:644-728 for the contract."""
    N = _npart(xyz)
    pipe = PeriodicPipeline.get(Ngrid)
    if not silent:
        print('------------------')
        print('%i positions in %i box' % (N, Lbox))
        print('--- calculating the FFT ---')
    half, sumw = pipe.fft_periodic(xyz, w, Lbox)
    out = pipe.pk_monopole(half, Lbox).cpu().numpy()
    nbar = _sum_w(w, N, sumw) / Lbox ** 3
    kf = 2 * np.pi / float(Lbox)
    Nbins = Ngrid // 2
    nk, ksum, psum = out[:Nbins], out[Nbins:2 * Nbins], out[2 * Nbins:]
    k = np.zeros(Nbins)
    p0k = np.zeros(Nbins)
    counts = np.zeros(Nbins)
    ok = nk > 0
    k[ok] = ksum[ok] / nk[ok]
    p0k[ok] = psum[ok] / nk[ok] / kf ** 3
    counts[ok] = nk[ok]
    p0k *= (2. * np.pi) ** 3
    if not silent:
        print('--- correcting for shotnoise ---')
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'N': N, 'nbar': nbar, 'kf': 2 * np.pi / Lbox}
    return {'meta': meta, 'k': k, 'p0k': p0k - 1. / nbar, 'counts': counts, 'p0k_sn': 1. / nbar}


def _pk_rsd_normalise(raw, kf32, Nbins, Nmubin):
    """The normalisation of estimator.f:246-262 and pyspectrum.py:634-640 applied to the raw K4 sums (host float64 array)."""
    nk = raw[:Nbins].copy()
    ks, p0k, p2k, p4k = [raw[(a + 1) * Nbins:(a + 2) * Nbins].copy() for a in range(4)]
    tb = Nbins * Nmubin
    o = 5 * Nbins
    n_kmu, k_kmu, mu_kmu, p_kmu = [raw[o + a * tb:o + (a + 1) * tb].reshape(Nmubin, Nbins).T.copy(order='F') for a in range(4)]
    kf3 = np.float64(np.float32(kf32 * kf32) * kf32)                 # dble(kf**3), f:249
    ok = nk > 0
    ks[ok] = ks[ok] / nk[ok]
    for p in (p0k, p2k, p4k):
        p[ok] = p[ok] / nk[ok] / kf3
    okm = n_kmu > 0
    k_kmu[okm] = k_kmu[okm] / n_kmu[okm]
    mu_kmu[okm] = mu_kmu[okm] / n_kmu[okm]
    p_kmu[okm] = p_kmu[okm] / n_kmu[okm] / kf3
    pk_norm = (2. * np.pi) ** 3
    p0k *= pk_norm
    p2k *= pk_norm
    p4k *= pk_norm
    p_kmu *= pk_norm
    return ks, p0k, p2k, p4k, nk, k_kmu, mu_kmu, p_kmu, n_kmu


def _pk_rsd_from_half(pipe, half, Lbox, rsd, Nmubin):
    """K4 multipoles of a device half field + the normalisation of estimator.f:246-262 and pyspectrum.py:634-640.
    Returns (ks, p0k, p2k, p4k, nk, k_kmu, mu_kmu, p_kmu, n_kmu) like the reference's _Pk_periodic_rsd."""
    raw, kf32 = pipe.pk_multipoles(half, int(Lbox), rsd, Nmubin)      # Lbox is an INTEGER dummy: estimator.f:158
    return _pk_rsd_normalise(raw.cpu().numpy(), kf32, pipe.N // 2, Nmubin)


def _Pk_periodic_rsd(delta, Lbox=None, rsd=2, Nmubin=5, code='fortran'):
    """pyspectrum.py:541-641.
    code='fortran' (py:628-641): multipoles and P(k,mu) of a host HALF field delta(k) of shape (Ngrid//2+1, Ngrid, Ngrid) indexed
        [kx,ky,kz] (what FFT_periodic returns) through the kernel that replaces estimator.pk_pbox_rsd (estimator.f:155-264).
    code='python' (py:545-626): the float64 (k,mu) estimator on a FULL field (Ngrid,Ngrid,Ngrid) indexed [kx,ky,kz] (what
        reflect_delta returns): |k_a| = min(i, N-i) so mu >= 0, mu bins ceil(mu*Nmubin), Legendre sums over the mu-binned modes;
        Lbox=None gives k in units of the fundamental mode.  Same outputs as the reference's branch (which also prints every
        (i,j) bin pair; that is not reproduced)."""
    if code == 'python':
        if rsd not in (0, 1, 2):
            raise ValueError('rsd must be 0, 1 or 2')
        delta = np.asarray(delta)
        Ngrid = delta.shape[0]
        if delta.shape != (Ngrid, Ngrid, Ngrid):
            raise ValueError("code='python' takes the full field (Ngrid,Ngrid,Ngrid), e.g. reflect_delta's output")
        pipe = PeriodicPipeline.get(Ngrid)
        kf = 1. if Lbox is None else 2 * np.pi / float(Lbox)                  # py:552-555
        arr = np.ascontiguousarray(delta.astype(np.complex64, copy=False))
        full = torch.from_numpy(arr.view(np.float32).reshape(Ngrid, Ngrid, Ngrid, 2)).to(pipe.dev)
        raw = pipe.pk_kmu_python(full, kf, rsd, Nmubin).cpu().numpy()
        Nbins = Ngrid // 2
        nks, ksum, p0, p2, p4 = [raw[a * Nbins:(a + 1) * Nbins].copy() for a in range(5)]
        tb = Nbins * Nmubin
        N_kmu, kk, mm, pp = [raw[5 * Nbins + a * tb:5 * Nbins + (a + 1) * tb].reshape(Nbins, Nmubin).copy() for a in range(4)]
        ok = nks > 0
        nk1 = np.where(ok, nks, 1.)
        ks = np.where(ok, ksum / nk1, 0.)
        pk_norm = (2. * np.pi) ** 3
        p0k = np.where(ok, p0 / nk1 / kf ** 3, 0.) * pk_norm
        p2k = np.where(ok, p2 / nk1 / kf ** 3 * 5., 0.) * pk_norm              # py:606-607
        p4k = np.where(ok, p4 / nk1 / kf ** 3 * 9., 0.) * pk_norm
        okm = N_kmu > 0
        n1 = np.where(okm, N_kmu, 1.)
        return (ks, p0k, p2k, p4k, nks, np.where(okm, kk / n1, 0.), np.where(okm, mm / n1, 0.),
                np.where(okm, pp / n1 / kf ** 3, 0.) * pk_norm, N_kmu)
    if code != 'fortran':
        raise ValueError("code must be 'fortran' or 'python'")
    if Lbox is None:
        raise ValueError('Lbox is required (pk_pbox_rsd takes it as an integer, estimator.f:158)')
    if rsd not in (0, 1, 2):
        raise ValueError('rsd must be 0, 1 or 2')
    delta = np.asarray(delta)
    Ngrid = delta.shape[1]
    if delta.shape != (Ngrid // 2 + 1, Ngrid, Ngrid):
        raise ValueError('delta must be the half field (Ngrid//2+1, Ngrid, Ngrid)')
    pipe = PeriodicPipeline.get(Ngrid)
    arr = np.ascontiguousarray(delta.astype(np.complex64, copy=False).transpose(2, 1, 0))           # [kz][ky][kx]
    half = torch.from_numpy(arr.view(np.float32).reshape(Ngrid, Ngrid, Ngrid // 2 + 1, 2)).to(pipe.dev)
    return _pk_rsd_from_half(pipe, half, Lbox, rsd, Nmubin)


def _counts_Bk123_f77(Ngrid=360, Nmax=40, Ncut=3, step=3, silent=True):
    """pyspectrum.py:1033-1057: triangle counts through estimator.bk_counts (estimator.f:2-102), cached in the reference's
    `.fort77` file.  Returned array is Fortran-ordered and indexed like the Fortran's coun(i<=j<=l) (NOT like _counts_Bk123).
    The reference allocates a fixed (40,40,40) array (py:1049); here the array is (Nmax,Nmax,Nmax).  The cache file holds the
    array in Fortran order and is read back the same way (the reference writes it with ndarray.tofile, i.e. in C order,
    and reads it with order='F', so its cache hits return the transposed array)."""
    from . import estimator as fEstimate
    fcnt = ''.join(['counts', '.Ngrid', str(Ngrid), '.Nmax', str(Nmax), '.Ncut', str(Ncut), '.step', str(step), '.fort77'])
    f_counts = os.path.join(dat_dir(), fcnt)
    if os.path.isfile(f_counts):
        return np.reshape(_read_fortran_record(f_counts, Nmax).ravel(), (Nmax, Nmax, Nmax), order='F')
    if not silent:
        print('-- %s does not exist --' % f_counts)
        print('-- computing %s --' % f_counts)
    counts = np.zeros((Nmax, Nmax, Nmax), dtype=np.float64, order='F')
    fEstimate.bk_counts(counts, Ngrid, float(step), Ncut, Nmax)
    try:
        os.makedirs(dat_dir(), exist_ok=True)
        _write_fortran_record(f_counts, counts.ravel(order='F'))
    except OSError:
        pass
    return counts


def Pk_periodic_rsd(xyz, w=None, Lbox=2600, Ngrid=360, rsd=2, Nmubin=10, fft='pyfftw', code='fortran', silent=True):
    """Power-spectrum multipoles + P(k,mu); see pyspectrum.py:460-538 for the contract."""
    N = _npart(xyz)
    nbar = float(N) / Lbox ** 3                      # py:510 (N, not sum w: SURVEY Q11)
    kf = 2 * np.pi / Lbox
    if rsd not in (0, 1, 2):
        raise ValueError('rsd must be 0, 1 or 2')
    pipe = PeriodicPipeline.get(Ngrid)
    if not silent:
        print('------------------')
        print('%i positions in %i box' % (N, Lbox))
        print('nbar = %f' % nbar)
    half, _ = pipe.fft_periodic(xyz, w, Lbox)
    if code == 'python':
        # the reference hands the HALF field to its full-grid python branch (py:519-522 -> IndexError at py:591); the evident
        # intent is the reflected field, which is what is passed here
        a = half.cpu().numpy().view(np.complex64)[..., 0].transpose(2, 1, 0)
        ks, p0k, p2k, p4k, nk, k_kmu, mu_kmu, p_kmu, n_kmu = _Pk_periodic_rsd(reflect_delta(a, Ngrid), Lbox=Lbox, rsd=rsd,
                                                                                Nmubin=Nmubin, code='python')
    elif code == 'fortran':
        ks, p0k, p2k, p4k, nk, k_kmu, mu_kmu, p_kmu, n_kmu = _pk_rsd_from_half(pipe, half, Lbox, rsd, Nmubin)
    else:
        raise ValueError("code must be 'fortran' or 'python'")
    if not silent:
        print('--- correcting for shotnoise ---')
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'N': N, 'nbar': nbar, 'kf': kf}
    return {'meta': meta, 'k': ks, 'p0k': p0k - 1. / nbar, 'p2k': p2k, 'p4k': p4k,
            'p_sn': np.repeat(1. / nbar, len(ks)), 'counts': nk, 'k_kmu': k_kmu, 'mu_kmu': mu_kmu,
            'p_kmu': p_kmu - 1. / nbar, 'counts_kmu': n_kmu}


def _bk_epilogue(Ngrid, tri, sums, sumsq, Nk, counts, step, Ncut, Nmax):
    """pyspectrum.py:404 and 415-456 on the host (a few thousand float64 operations)."""
    s0 = Ncut // step
    p0k = np.zeros(Nmax)
    for j in range(s0, Nmax + 1):
        p0k[j - 1] = sumsq[j - s0] / Ngrid ** 3 / Nk[j]
    i, j, l = tri[:, 0], tri[:, 1], tri[:, 2]
    fac = _shell_fac(tri)
    c = counts[i - 1, j - 1, l - 1]
    pos = c > 0
    cs = np.where(pos, c, 1.)
    pi_, pj_, pl_ = p0k[i - 1], p0k[j - 1], p0k[l - 1]
    b123 = np.where(pos, sums / cs, 0.)
    with np.errstate(divide='ignore', invalid='ignore'):
        q123 = np.where(pos, sums / cs / (pi_ * pj_ + pj_ * pl_ + pl_ * pi_), 0.)
    out = {}
    out['i_k1'] = i[pos].astype(np.int64) * step       # index lists skip empty triangles, value lists do not (Q9)
    out['i_k2'] = j[pos].astype(np.int64) * step
    out['i_k3'] = l[pos].astype(np.int64) * step
    out['p0k1'] = np.where(pos, pi_, 0.)
    out['p0k2'] = np.where(pos, pj_, 0.)
    out['p0k3'] = np.where(pos, pl_, 0.)
    out['b123'] = b123
    out['q123'] = q123
    out['counts'] = np.where(pos, c / (fac * float(Ngrid ** 3)), 0.)
    return out


def _prefetched(catalogues, Ngrid):
    """Yield (xyz_dev, w_dev) for every catalogue of the iterable, uploading catalogue n+1 on a copy stream while the
    caller works on catalogue n (host->device copies overlap the kernels when the host arrays are pinned torch tensors;
    pageable numpy arrays work too, the copy then just does not overlap).  Items are `xyz` or `(xyz, w)`."""
    pipe = PeriodicPipeline.get(Ngrid)
    if getattr(pipe, '_copy_stream', None) is None:
        pipe._copy_stream = torch.cuda.Stream(device=pipe.dev)      # kept: its allocator pool then recycles the upload buffers
    copy_stream = pipe._copy_stream

    def stage(item):
        xyz, w = item if isinstance(item, (tuple, list)) else (item, None)
        with torch.cuda.stream(copy_stream):
            pos, aos, wt = pipe.to_device(xyz, w)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return (pos.t() if aos else pos), wt, ev

    it = iter(catalogues)
    nxt = None
    for item in it:
        nxt = stage(item)
        break
    while nxt is not None:
        cur, nxt = nxt, None
        for item in it:                                  # enqueue the next upload BEFORE this catalogue's blocking result read
            nxt = stage(item)
            break
        main = torch.cuda.current_stream(pipe.dev)
        main.wait_event(cur[2])
        for t in cur[:2]:
            if t is not None and t.is_cuda:
                t.record_stream(main)                    # allocated on the copy stream, consumed on the main stream
        yield cur[0], cur[1]


def Bk_periodic_many(catalogues, Lbox=2600, Ngrid=360, step=3, Ncut=3, Nmax=40, fft='pyfftw', nthreads=1, silent=True):
    """Generator: Bk_periodic (pyspectrum.py:285-356) for every catalogue of an iterable -- the thousands-of-mocks use of the
    reference -- as a software pipeline: the upload of catalogue n+1 runs on a copy stream and the kernels of catalogue n are
    queued while the host still waits for / post-processes catalogue n-1.  Yields the same dictionaries as Bk_periodic, in order
    (each one after the following catalogue has been queued).  Two catalogues' shell fields are alive at a time."""
    pending = None
    for xyz_d, w_d in _prefetched(catalogues, Ngrid):
        h = _bk_launch(xyz_d, w_d, Lbox, Ngrid, step, Ncut, Nmax, fft, silent)     # kernels of catalogue n are queued ...
        if pending is not None:
            yield _bk_finish(pending)                                              # ... before the host waits for catalogue n-1
        pending = h
    if pending is not None:
        yield _bk_finish(pending)


def Pk_periodic_many(catalogues, Lbox=2600, Ngrid=360, fft='pyfftw', silent=True):
    """Generator: Pk_periodic (pyspectrum.py:644-728) over many catalogues with overlapped uploads."""
    for xyz_d, w_d in _prefetched(catalogues, Ngrid):
        yield Pk_periodic(xyz_d, w=w_d, Lbox=Lbox, Ngrid=Ngrid, fft=fft, silent=silent)


def Pk_periodic_rsd_many(catalogues, Lbox=2600, Ngrid=360, rsd=2, Nmubin=10, fft='pyfftw', code='fortran', silent=True):
    """Generator: Pk_periodic_rsd (pyspectrum.py:460-538) over many catalogues with overlapped uploads."""
    for xyz_d, w_d in _prefetched(catalogues, Ngrid):
        yield Pk_periodic_rsd(xyz_d, w=w_d, Lbox=Lbox, Ngrid=Ngrid, rsd=rsd, Nmubin=Nmubin, fft=fft, code=code, silent=silent)


def _bk_launch(xyz, w, Lbox, Ngrid, step, Ncut, Nmax, fft, silent):
    """First half of Bk_periodic: everything that only enqueues GPU work (assignment, FFT, shell fields, triangle sums, result
    copy).  The exact triangle counts and shell mode counts are cached per configuration."""
    N = _npart(xyz)
    s0 = Ncut // step
    if s0 < 1:
        raise ValueError('Ncut//step must be >= 1 (the reference wraps p0k[-1] there, SURVEY Q9)')
    pipe = PeriodicPipeline.get(Ngrid)
    if not silent:
        print('------------------')
        print('%i positions in %i box' % (N, Lbox))
        print('--- calculating the FFT ---')
    Nk = pipe.shell_mode_counts(step, Nmax)
    counts = pipe.counts(Nmax, Ncut, step, fft=fft, silent=silent)
    sink = [] if _Range.timers else None
    with _Range('assign+fft+fcomb', sink):
        half, sumw = pipe.fft_periodic(xyz, w, Lbox)
    if not silent:
        print('--- calculating the bispectrum ---')
    with _Range('shells+triangles', None):
        h = pipe.bispectrum_launch(half, step, Ncut, Nmax, sumw=sumw, timers=sink)
    h.update(timers=sink, N=N, w=None if (w is None or isinstance(w, torch.Tensor)) else w, w_is_dev=isinstance(w, torch.Tensor), Nk=Nk, counts=counts,
             cfg=(Lbox, Ngrid, step, Ncut, Nmax), silent=silent)
    return h


def _bk_finish(h):
    """Second half of Bk_periodic: wait for the results and run the float64 host epilogue (pyspectrum.py:321-356)."""
    Lbox, Ngrid, step, Ncut, Nmax = h['cfg']
    silent, N = h['silent'], h['N']
    pipe = PeriodicPipeline.get(Ngrid)
    kf = 2 * np.pi / Lbox
    sums_h, sumsq_h, sumw_dev = pipe.bispectrum_finish(h)
    tri = triangle_list(Nmax, Ncut, step)
    # np.sum(w) of pyspectrum.py:323 (host value when the caller's weights live on the host)
    sum_w = float(np.sum(h['w'])) if h['w'] is not None else (sumw_dev if h['w_is_dev'] else float(N))
    nbar = sum_w / Lbox ** 3
    if not silent:
        print('sum w_i = %f' % (nbar * Lbox ** 3))
        print('nbar = %f' % nbar)
    bispec = _bk_epilogue(Ngrid, tri, sums_h, sumsq_h, h['Nk'], h['counts'], step, Ncut, Nmax)
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'step': step, 'Ncut': Ncut, 'Nmax': Nmax, 'N': N, 'nbar': nbar, 'kf': kf}
    if h.get('timers'):
        ms = {}
        for name, a, b in h['timers']:
            ms[name] = ms.get(name, 0.) + a.elapsed_time(b)
        meta['stage_ms'] = ms                            # PSB_TIMERS=1 only
    bispec['meta'] = meta
    if not silent:
        print('--- correcting for shotnoise ---')
    bispec['p0k1'] = bispec['p0k1'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k2'] = bispec['p0k2'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k3'] = bispec['p0k3'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k_sn'] = 1. / nbar
    b_shotnoise = (bispec['p0k1'] + bispec['p0k2'] + bispec['p0k3']) / nbar + 1. / nbar ** 2
    bispec['b123'] = bispec['b123'] * (2 * np.pi) ** 6 / kf ** 6 - b_shotnoise
    bispec['b123_sn'] = b_shotnoise
    with np.errstate(divide='ignore', invalid='ignore'):
        bispec['q123'] = bispec['b123'] / (bispec['p0k1'] * bispec['p0k2'] + bispec['p0k1'] * bispec['p0k3'] + bispec['p0k2'] * bispec['p0k3'])
    return bispec


def Bk_periodic(xyz, w=None, Lbox=2600, Ngrid=360, step=3, Ncut=3, Nmax=40, fft='pyfftw', nthreads=1, silent=True):
    """Bispectrum of a periodic box; see pyspectrum.py:285-356 for the contract."""
    return _bk_finish(_bk_launch(xyz, w, Lbox, Ngrid, step, Ncut, Nmax, fft, silent))


# ----------------------------------------------------------------------------------------------
# survey geometry: monopole bispectrum with the FKP / Scoccimarro (2015) normalisation  
# ----------------------------------------------------------------------------------------------
def _shell_fac(tri):
    """Symmetry factor of pyspectrum.py:237-241 / 419-423 for the triples (i,j,l)."""
    i, j, l = tri[:, 0], tri[:, 1], tri[:, 2]
    fac = np.ones(len(tri))
    fac[(j == l) & (i == j)] = 6.
    fac[(i == j) & (j != l)] = 2.
    fac[(i == l) & (l != j)] = 2.
    fac[(j == l) & (l != i)] = 2.
    return fac


def _fft_survey_mono_dev(radecz, nb, w, P0_fkp, Lbox, Ngrid, cosmo, silent):
    """Device half of FFT_survey_mono: returns (half field on the device, Ntot, I12, I13, I22, I23, I33)."""
    from . import util as UT
    radecz = np.asarray(radecz)
    assert radecz.shape[0] == 3, "radecz has to be have shape [3,N]"
    N = int(radecz.shape[1])
    if cosmo is None:
        cosmo = UT.FlatLambdaCDM(H0=67.6, Om0=0.31)                  # py:769-770
    pipe = PeriodicPipeline.get(Ngrid)
    # util.radecz_to_cartesian, the float32 cast (py:789-792), the FKP weights (py:797-799; the caller's w is left alone)
    # and the sums of py:794-806 in one pass over the catalogue on the device
    xyz, wf, out = pipe.survey_prepare(radecz, nb, w, P0_fkp, cosmo)
    Ntot, I12, I13, I22, I23, I33 = [float(v) for v in out[:6]]
    xyz_min, xyz_max = out[6:9], out[9:12]
    if not silent:
        print(['%.1f < %s < %.1f\n' % (mi, _axis, ma) for _axis, mi, ma in zip(['x', 'y', 'z'], xyz_min, xyz_max)])
    assert np.sum(xyz_max >= 0.5 * Lbox) + np.sum(np.abs(xyz_min) >= 0.5 * Lbox) == 0, 'box not big enough!'
    if not silent:
        print('%i positions' % N)
        print('Ntot=%.2f' % Ntot)
        for name, val in zip(['I12', 'I13', 'I22', 'I23', 'I33'], [I12, I13, I22, I23, I33]):
            print('%s=%.2e' % (name, val))
    assert np.all([(I12 >= 0), (I13 >= 0), (I22 >= 0), (I23 >= 0), (I33 >= 0)])
    half = pipe.fft_survey(xyz, wf, Lbox)
    if not silent:
        print('delta_0(k) complete')
    return half, Ntot, I12, I13, I22, I23, I33


def FFT_survey_mono(radecz, nb, w=None, P0_fkp=1e6, Lbox=2600., Ngrid=360, cosmo=None, fft='pyfftw', silent=True):
    """pyspectrum.py:731-826: (RA, Dec, z) catalogue -> FKP-weighted delta_0(k) on the half grid, complex64
    (Ngrid//2+1, Ngrid, Ngrid) indexed [kx,ky,kz], plus Ntot = sum w and the normalisation sums I12, I13, I22, I23, I33.
    Unlike the reference, neither `w` (py:799 multiplies it by the FKP weight in place) nor the RA/Dec rows of `radecz`
    (util.py:41-42 converts them to radians in place) are modified."""
    half, Ntot, I12, I13, I22, I23, I33 = _fft_survey_mono_dev(radecz, nb, w, P0_fkp, Lbox, Ngrid, cosmo, silent)
    a = half.cpu().numpy().view(np.complex64)[..., 0]                # (kz,ky,kx) C-order
    return a.transpose(2, 1, 0), Ntot, I12, I13, I22, I23, I33


def _b0_survey_dev(half, Ngrid, alpha, I12, I13, I22, I23, I33, Nmax, Ncut, step, fft, silent):
    """Shell fields + triangle sums of a device half field and the survey normalisation of pyspectrum.py:216-281."""
    s0 = Ncut // step
    if s0 < 1:
        raise ValueError('Ncut//step must be >= 1 (the reference wraps p0k[-1] there, SURVEY Q9)')
    pipe = PeriodicPipeline.get(Ngrid)
    Nk = pipe.shell_mode_counts(step, Nmax)
    counts = pipe.counts(Nmax, Ncut, step, fft=fft, silent=silent)
    sums, sumsq = pipe.bispectrum_sums(half, step, Ncut, Nmax)
    tri = triangle_list(Nmax, Ncut, step)
    p0k = np.zeros(Nmax)
    for j in range(s0, Nmax + 1):
        p0k[j - 1] = sumsq[j - s0] / Ngrid ** 3 / Nk[j]
    p0k /= I22                                                       # py:218-220
    p0k -= (1. + alpha) * I12 / I22
    i, j, l = tri[:, 0], tri[:, 1], tri[:, 2]
    c = counts[i - 1, j - 1, l - 1]
    pos = c > 0
    cs = np.where(pos, c, 1.)
    pi_, pj_, pl_ = p0k[i - 1], p0k[j - 1], p0k[l - 1]
    b = sums / cs
    b = b - ((pi_ + pj_ + pl_) * I23 + (1. - alpha ** 2) * I13)      # py:250-253
    b = b / I33
    with np.errstate(divide='ignore', invalid='ignore'):
        q = b / (pi_ * pj_ + pj_ * pl_ + pl_ * pi_)
    out = {}
    out['i_k1'] = i[pos].astype(np.int64) * step                     # index lists skip empty triangles, value lists do not
    out['i_k2'] = j[pos].astype(np.int64) * step
    out['i_k3'] = l[pos].astype(np.int64) * step
    out['p0k1'] = np.where(pos, pi_, 0.)
    out['p0k2'] = np.where(pos, pj_, 0.)
    out['p0k3'] = np.where(pos, pl_, 0.)
    out['b123'] = np.where(pos, b, 0.)
    out['q123'] = np.where(pos, q, 0.)
    out['counts'] = np.where(pos, c / (_shell_fac(tri) * float(Ngrid ** 3)), 0.)
    return out


def _B0_survey(delta, alpha, I12, I13, I22, I23, I33, Nmax=40, Ncut=3, step=3, fft='pyfftw', nthreads=1, silent=True):
    """pyspectrum.py:135-282: bispectrum monopole of a full (Ngrid,Ngrid,Ngrid) delta(k) = data - alpha * randoms with
    the survey shot-noise terms and normalisation.  `delta` is a host array indexed [kx,ky,kz] as reflect_delta returns."""
    Ngrid = np.asarray(delta).shape[0]
    half = PeriodicPipeline.get(Ngrid).half_from_full(delta)
    return _b0_survey_dev(half, Ngrid, alpha, I12, I13, I22, I23, I33, Nmax, Ncut, step, fft, silent)


def _Bk_periodic(delta, Nmax=40, Ncut=3, step=3, fft='pyfftw', nthreads=1, silent=True):
    """pyspectrum.py:359-457: raw bispectrum (no units, no shot noise) of a full (Ngrid,Ngrid,Ngrid) delta(k) host array."""
    Ngrid = np.asarray(delta).shape[0]
    s0 = Ncut // step
    if s0 < 1:
        raise ValueError('Ncut//step must be >= 1 (the reference wraps p0k[-1] there, SURVEY Q9)')
    pipe = PeriodicPipeline.get(Ngrid)
    half = pipe.half_from_full(delta)
    sums, sumsq = pipe.bispectrum_sums(half, step, Ncut, Nmax)
    return _bk_epilogue(Ngrid, triangle_list(Nmax, Ncut, step), sums, sumsq, pipe.shell_mode_counts(step, Nmax),
                        pipe.counts(Nmax, Ncut, step, fft=fft, silent=silent), step, Ncut, Nmax)


def B0_survey(radecz, nbar, w=None, radecz_r=None, nbar_r=None, w_r=None, P0_fkp=1e6, Lbox=2600, Ngrid=360, step=3, Ncut=3,
              Nmax=40, cosmo=None, fft='pyfftw', nthreads=1, silent=True):
    """Bispectrum monopole for survey geometry (Scoccimarro 2015 estimator); see pyspectrum.py:13-132 for the contract.
    Data and random catalogues are assigned and transformed on the GPU, delta = delta_d - alpha * delta_r is formed on the
    device half fields (in float64, rounded once to complex64 -- what numpy >= 2 gives the reference at py:123 and 204), and
    the shell / triangle stage is the one of Bk_periodic.  The `Ngrid == 360` assert (py:90) is relaxed as for Bk_periodic."""
    if radecz_r is None:
        raise ValueError('specify a random catalog')                 # the reference fails on radecz_r.shape (py:105)
    if nbar_r is None:
        raise ValueError('specify nbar for random catalog')
    kf = 2 * np.pi / Lbox
    N = int(np.asarray(radecz).shape[1])
    if not silent:
        print('--- %i positions in %i box ---' % (N, Lbox))
        print('--- calculating the FFT for data ---')
    half_d, Ngtot, I12d, I13d, I22d, I23d, I33d = _fft_survey_mono_dev(radecz, nbar, w, P0_fkp, Lbox, Ngrid, cosmo, silent)
    if not silent:
        print('--- calculating the FFT for random ---')
    half_r, Nrtot, I12r, I13r, I22r, I23r, I33r = _fft_survey_mono_dev(radecz_r, nbar_r, w_r, P0_fkp, Lbox, Ngrid, cosmo, silent)
    alpha = Ngtot / Nrtot
    if not silent:
        print('alpha=%e' % alpha)
    half = (half_d.double() - float(alpha) * half_r.double()).float()
    del half_d, half_r
    if not silent:
        print('--- calculating the bispectrum ---')
    bispec = _b0_survey_dev(half, Ngrid, alpha, alpha * I12r, alpha * I13r, alpha * I22r, alpha * I23r, alpha * I33r,
                            Nmax, Ncut, step, fft, True)
    bispec['meta'] = {'Lbox': Lbox, 'Ngrid': Ngrid, 'step': step, 'Ncut': Ncut, 'Nmax': Nmax, 'N': N, 'nbar': nbar, 'kf': kf}
    return bispec
