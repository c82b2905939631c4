"""pyspectrum_b200.multigpu -- ONE catalogue sharded over the GPUs of a node (one process per GPU, torchrun + NCCL).

The slab design of SURVEY 8e; every exchange is a data-plane collective on NVLink, nothing is replicated at grid size:

  1. ROUTE      every rank holds an arbitrary shard of the particles.  Rank q owns the mesh planes [q nz, (q+1) nz), nz = N/G; a
                particle in cell c touches planes c-1 .. c+3, so it is sent (float32 {x,y,z,w} after the float64 clip of py:938-941)
                to the owner of plane c-1 and, if different, of plane c+3 (ghost copy)      [16 B per particle (+ ghosts): peer stores
                from inside the route kernel under NCCL, else an all-to-all]
  2. ASSIGN     K1 on the received particles onto the rank's OWN planes only (psb_assign_slab): no mesh collective at all
  3. FFT        x and y passes in the slab, separation of the two interlaced grids' spectra (the conjugate partner (-kx,-ky) is
                in the same plane), z-slabs -> ky-slabs                                    [all-to-all: 2 x 4 (N/2+1) N^2 / G B per rank]
                z pass + point-wise fcomb: the rank now owns the rows ky in [q ny, (q+1) ny) of delta(k)
  4. P(k)       K4 on the ky-slab (psb_pk_*_slab) -> partial bins                          [all-reduce: Nbin (5 + 4 Nmu) float64]
  5. CARRIER    the shells of B(k) only reach |k_a| <= R = floor(step (Nmax + 1/2)), and all their triangles close without a wrap
                on any grid with more than R_i + R_j + R_l points per side (pyspectrum.coarse_levels): every rank copies the low-k
                modes it owns into the half field of a small carrier grid Ng                [all-reduce: 4 (Ng/2+1) Ng^2 B]
                (Ng = N when the reference's own grid lets triangles wrap, e.g. Ngrid=360 with Nmax=40: then an all-gather of the ky-slabs.)
  6. SHELLS     per transform grid (level) of the carrier: the packed shell PAIRS are dealt round-robin, each rank transforms its
                pairs (K5), every field is cut into G slabs of cells, slab q -> rank q      [ONE all-to-all per level: 4 S Nc^3 (G-1)/G B]
  7. TRIANGLES  slab-local K6 on every rank, partial sums                                  [all-reduce: Ntri + 3 S float64, once]
  8. COUNTS     the float64 exact counts use steps 6-7 with delta == 1 (cached per configuration)

Every kernel is the single-GPU one (K1 with a plane window, K4 with a row window); the collectives are torch.distributed calls so
the same code runs under NCCL (GPU) and gloo (the CPU tests of the host logic: tests/test_multigpu_host.py).  They are bulk
exchanges between kernels, except step 6 under NCCL: there the z pass of K5 stores every output plane straight into the slab
buffer of the rank that owns it (peer pointers from symmetric memory, `SlabBuffers`), so the exchange rides inside the transform
and the separate all-to-all disappears (it remains the path for float64 counts, gloo, and grids that do not divide into planes).  `stats` dictionaries collect the bytes each collective puts on the wire
and, with `timed=True`, its device time (bench.py reports them)."""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import pyspectrum as P


# ------------------------------------------------------------------------------------------ plumbing
def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _rank():
    return dist.get_rank() if _world() > 1 else 0


def _allreduce(t, op=None):
    if _world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op is None else op)
    return t


class Stats(object):
    """Bytes on the wire and (optionally) device time per named collective / stage of one sharded call."""

    def __init__(self, timed=False):
        self.timed, self.bytes, self.ev = timed, {}, []

    def add_bytes(self, name, n):
        self.bytes[name] = self.bytes.get(name, 0) + int(n)

    def mark(self):
        if not self.timed or not torch.cuda.is_available():
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def span(self, name, e0):
        if e0 is not None:
            self.ev.append((name, e0, self.mark()))

    def times_ms(self):
        out = {}
        for name, a, b in self.ev:
            out[name] = out.get(name, 0.) + a.elapsed_time(b)
        return out


def _stats(stats):
    return stats if stats is not None else Stats()


# ------------------------------------------------------------------------------------------ shell pairs and slabs of cells
def pair_assignment(npairs, world):
    """Round-robin deal of packed shell pairs; every rank gets the same number (padding pairs = empty shells)."""
    per = (npairs + world - 1) // world
    return [[r + world * k for k in range(per)] for r in range(world)], per


def slab_field_rows(S, world, per):
    """After the all-to-all the slab buffer is laid out [owner rank][k][e] (k = local pair index, e = 0/1 within the pair);
    returns for every shell slot f the row of that buffer which holds it."""
    rows = []
    for f in range(S):
        p, e = divmod(f, 2)
        owner, k = p % world, p // world
        rows.append(owner * 2 * per + 2 * k + e)
    return rows


def check_cell_slabs(ncell, world, packed=True):
    """The cells of a field are cut into `world` equal slabs; packed fields hold aligned cell PAIRS and the tensor-core kernel
    works on 64-cell chunks.  Raised on every rank before any collective."""
    if ncell % world:
        raise ValueError('the number of cells (%d) must be a multiple of the number of ranks (%d)' % (ncell, world))
    if packed and (ncell // world) % 2:
        raise ValueError('a slab of %d cells would split a packed cell pair' % (ncell // world))


def exchange_to_slabs(fields_local, world, stats=None):
    """fields_local: [2*per, ncell] on every rank (its shell pairs, full grid).  ONE all-to-all; returns [world*2*per, ncell/world]:
    row (owner*2*per + 2*k + e) = this rank's slab of shell e of the k-th pair of rank `owner` (see slab_field_rows)."""
    nrow, ncell = fields_local.shape
    check_cell_slabs(ncell, world, packed=False)
    slab = ncell // world
    if world == 1:
        return fields_local
    send = fields_local.view(nrow, world, slab).permute(1, 0, 2).contiguous()          # [dest][row][cell]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                                                   # recv[q] = rank q's rows of MY slab
    _stats(stats).add_bytes('shell_fields_all_to_all', send.numel() * send.element_size() * (world - 1) // world)
    return recv.view(world * nrow, slab)


class SlabBuffers(object):
    """The slab buffers of one level -- [world*2*per, slab] 32-bit words on every rank -- with every rank holding a device pointer
    to every other rank's copy, so that K5's z-pass epilogue can store each output plane straight into the memory of the rank that
    owns it (NVLink peer stores from inside the transform; SURVEY 8e "z-pass epilogue -> peer-memory slab write") instead of
    writing a local field and exchanging it afterwards.  The buffers come from torch's symmetric memory (CUDA VMM allocations
    mapped into every process of the group at rendezvous); `barrier()` is its stream-ordered device barrier.
    `emulated(...)` builds the same object from plain tensors of ONE device (tests: the ranks take turns)."""
    _cache = {}

    def __init__(self, world, rank, nrow, slab, local, ptrs, handle=None, all_local=None):
        self.world, self.rank, self.nrow, self.slab = world, rank, nrow, slab
        self.local, self.ptrs, self.handle, self.all_local = local, [int(p) for p in ptrs], handle, all_local
        self._tables = {}

    @classmethod
    def get(cls, dev, world, rank, nrow, slab):
        """Cached per shape: the rendezvous is a host-synchronising collective.  Raises if symmetric memory is unavailable."""
        key = (str(dev), world, nrow, slab)
        if key not in cls._cache:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty((nrow, slab), dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            cls._cache[key] = cls(world, rank, nrow, slab, buf, hdl.buffer_ptrs, hdl)
        return cls._cache[key]

    @classmethod
    def emulated(cls, dev, world, nrow, slab):
        bufs = [torch.zeros((nrow, slab), dtype=torch.float32, device=dev) for _ in range(world)]
        return [cls(world, r, nrow, slab, bufs[r], [b.data_ptr() for b in bufs], None, bufs) for r in range(world)]

    def barrier(self):
        if self.handle is not None:
            self.handle.barrier()

    def row_table(self, dev):
        """int64 [nrow, world] (device): address of row r of every rank's buffer (the slab-FFT exchange: rows = P, Q)."""
        if 'rows' not in self._tables:
            t = np.array([[p + 4 * r * self.slab for p in self.ptrs] for r in range(self.nrow)], np.int64)
            self._tables['rows'] = torch.from_numpy(t).to(dev)
        return self._tables['rows']

    def route(self, per, plist, S, dev):
        """Route table [len(plist), 2, world] (int64, device) for this rank's pairs `plist` (local pair k = position in the list):
        entry (k, e, q) = address at which shell e of the pair WOULD start in rank q's buffer if that buffer held the whole field,
        i.e. row (rank*2*per + 2k + e) of the slab buffer (slab_field_rows) minus the q slabs in front; 0 = not stored."""
        key = (per, tuple(plist), S)
        if key not in self._tables:
            t = np.zeros((len(plist), 2, self.world), np.int64)
            for k, pidx in enumerate(plist):
                for e in (0, 1):
                    if 2 * pidx + e >= S:
                        continue
                    row = self.rank * 2 * per + 2 * k + e
                    for q in range(self.world):
                        t[k, e, q] = self.ptrs[q] + 4 * (row * self.slab - q * self.slab)
            self._tables[key] = torch.from_numpy(t).to(dev)
        return self._tables[key]


ROUTED_GRIDS = (24, 32, 36, 48, 64, 128, 256, 320, 360, 400, 512, 1024)      # grids with a compiled multi-stage plan (psb_fft_lines.cuh)
_ROUTED_BROKEN = []


def routed_slabs(dev, world, rank, nrow, slab):
    """SlabBuffers for the peer-store exchange, or None (PSB_SHARDED_ROUTED=0, or symmetric memory cannot be set up on this
    system: reported once, the bulk all-to-all is used instead).  Every rank takes the same decision: the outcome of the
    rendezvous is agreed with an all-reduce."""
    if world == 1 or os.environ.get('PSB_SHARDED_ROUTED', '1') == '0' or _ROUTED_BROKEN:
        return None
    if (str(dev), world, nrow, slab) in SlabBuffers._cache:
        return SlabBuffers._cache[(str(dev), world, nrow, slab)]
    ok, bufs = 1, None
    try:
        bufs = SlabBuffers.get(dev, world, rank, nrow, slab)
    except Exception as exc:                             # noqa: BLE001 -- any failure of the VMM / handle exchange
        ok, err = 0, exc
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if not _ROUTED_BROKEN and rank == 0:
            print('pyspectrum_b200.multigpu: symmetric memory unavailable (%s); shell fields go through all_to_all'
                  % (repr(err) if not ok else 'on another rank'))
        _ROUTED_BROKEN.append(True)
        return None
    return bufs


# ------------------------------------------------------------------------------------------ 1-2: route + slab assignment
def slab_geometry(N, world):
    """(planes per rank, padded half-row length): N must divide by the number of ranks; rows hold kx = 0..N/2 padded to even."""
    if N % world:
        raise ValueError('Ngrid must be a multiple of the number of ranks for the slab path')
    return N // world, (N // 2 + 2) // 2 * 2


def route_counts(pipe, pos, aos, wt, Lbox, world, offset=0., clip=True):
    """Per-destination particle counts (ghost copies included) of this rank's shard + the float64 sum of its weights (device)."""
    N = pipe.N
    nz = N // world
    Np = pos.shape[0] if aos else pos.shape[1]
    counts = torch.zeros(world, dtype=torch.int64, device=pipe.dev)
    sumw = torch.zeros(1, dtype=torch.float64, device=pipe.dev)
    P.check(pipe.L.psb_slab_route_count(P._ptr(pos), int(pos.dtype == torch.float64), aos, P._ptr(wt),
                                        int(wt is not None and wt.dtype == torch.float64), Np, N, float(Lbox) if clip else 0.0,
                                        np.float32(float(N) / Lbox), np.float32(offset), nz, world, P._ptr(counts), P._ptr(sumw),
                                        P._stream()), 'psb_slab_route_count')
    return counts, sumw


def route_scatter(pipe, pos, aos, wt, Lbox, world, counts, offset=0., clip=True, total=None):
    """Destination-major send buffer [sum(counts), 4] float32 {x,y,z,w} (total: sum(counts) if the caller already has it on the host)."""
    N = pipe.N
    Np = pos.shape[0] if aos else pos.shape[1]
    base = (torch.cumsum(counts, 0) - counts).contiguous()
    cursor = torch.zeros(world, dtype=torch.int64, device=pipe.dev)
    if total is None:
        total = int(counts.sum().item())
    send = torch.empty((max(total, 1), 4), dtype=torch.float32, device=pipe.dev)
    P.check(pipe.L.psb_slab_route_scatter(P._ptr(pos), int(pos.dtype == torch.float64), aos, P._ptr(wt),
                                          int(wt is not None and wt.dtype == torch.float64), Np, N, float(Lbox) if clip else 0.0,
                                          np.float32(float(N) / Lbox), np.float32(offset), N // world, world, P._ptr(base), P._ptr(cursor),
                                          P._ptr(send), P._stream()), 'psb_slab_route_scatter')
    return send[:total]


def route_particles(pipe, xyz_local, w_local, Lbox, offset=0., clip=True, stats=None):
    """Step 1.  Returns (received particles [n,4] float32 on the device, global sum of weights (device, float64), global N)."""
    st = _stats(stats)
    world = _world()
    slab_geometry(pipe.N, world)
    if pipe.N // world < 8:
        raise ValueError('need at least 8 mesh planes per rank')
    pos, aos, wt = pipe.to_device(xyz_local, w_local)
    Np = pos.shape[0] if aos else pos.shape[1]
    e0 = st.mark()
    counts, sumw = route_counts(pipe, pos, aos, wt, Lbox, world, offset, clip)
    if world == 1:
        send = route_scatter(pipe, pos, aos, wt, Lbox, world, counts, offset, clip)
        st.span('route_kernels', e0)
        return send, sumw, Np
    # the small collectives first, then ONE device -> host copy with everything the host needs (split sizes both ways, N)
    meta = torch.cat([sumw, torch.tensor([float(Np)], dtype=torch.float64, device=pipe.dev)])
    _allreduce(meta)
    if _peer_exchange_enabled():
        got = _route_particles_peer(pipe, pos, aos, wt, Lbox, world, counts, meta, offset, clip, st, e0)
        if got is not None:
            return got
    rc = torch.empty_like(counts)
    dist.all_to_all_single(rc, counts)
    host = torch.cat([counts.double(), rc.double(), meta[1:]]).cpu().tolist()
    sc_h, rc_h, ntot = [int(v) for v in host[:world]], [int(v) for v in host[world:2 * world]], int(round(host[2 * world]))
    send = route_scatter(pipe, pos, aos, wt, Lbox, world, counts, offset, clip, total=sum(sc_h))
    st.span('route_kernels', e0)
    e0 = st.mark()
    recv = torch.empty((max(sum(rc_h), 1), 4), dtype=torch.float32, device=pipe.dev)[:sum(rc_h)]
    dist.all_to_all_single(recv, send, output_split_sizes=rc_h, input_split_sizes=sc_h)
    st.span('particles_all_to_all', e0)
    st.add_bytes('particles_all_to_all', 16 * (sum(sc_h) - sc_h[_rank()]))
    return recv, meta[:1].clone(), ntot


_RECV_CAP = {}


def peer_segments(counts_matrix, rank):
    """Host bookkeeping of the peer-store particle exchange.  counts_matrix[s][d] = copies rank s sends to rank d (all-gathered, the
    same on every rank).  Returns (recv_tot[d] = particles rank d receives, seg[d] = offset in particles of THIS rank's segment in
    rank d's receive buffer: the ranks in front of it write first)."""
    m = np.asarray(counts_matrix, dtype=np.int64)
    return m.sum(axis=0), m[:rank].sum(axis=0)


def _peer_exchange_enabled():
    return os.environ.get('PSB_SHARDED_ROUTED', '1') != '0' and not _ROUTED_BROKEN


def _route_particles_peer(pipe, pos, aos, wt, Lbox, world, counts, meta, offset, clip, st, e0):
    """The particle exchange as NVLink peer stores from inside the route kernel (psb_slab_route_scatter_peer) instead of a send
    buffer + all-to-all: the counts of all ranks are all-gathered, so every rank knows where its segment starts in every receive
    buffer; the buffers are one symmetric allocation per process group, grown when a catalogue needs more.  Returns None when
    symmetric memory is unavailable (the caller then takes the bulk path)."""
    rank = _rank()
    N = pipe.N
    Np = pos.shape[0] if aos else pos.shape[1]
    mat = torch.empty(world * world, dtype=counts.dtype, device=pipe.dev)
    dist.all_gather_into_tensor(mat, counts.contiguous())                     # mat[s * world + d] = copies rank s sends to rank d
    host = torch.cat([mat.double(), meta[1:]]).cpu().numpy()
    M_ = np.rint(host[:world * world]).astype(np.int64).reshape(world, world)
    ntot = int(round(host[world * world]))
    recv_tot, seg = peer_segments(M_, rank)
    need = int(recv_tot.max())
    key = (str(pipe.dev), world)
    cap = _RECV_CAP.get(key, 0)
    if cap < need:                                       # same decision on every rank (same matrix): grow the symmetric buffer
        if cap:
            SlabBuffers._cache.pop((str(pipe.dev), world, 1, 4 * cap), None)
        cap = int(1.25 * need) + 4096
        _RECV_CAP[key] = cap
    bufs = routed_slabs(pipe.dev, world, rank, 1, 4 * cap)
    if bufs is None:
        _RECV_CAP.pop(key, None)
        return None
    dest = torch.from_numpy(np.array([bufs.ptrs[d] + 16 * int(seg[d]) for d in range(world)], np.int64)).to(pipe.dev)
    cursor = torch.zeros(world, dtype=torch.int64, device=pipe.dev)
    st.span('route_kernels', e0)
    e1 = st.mark()
    bufs.barrier()                                       # every rank has consumed its previous catalogue
    P.check(pipe.L.psb_slab_route_scatter_peer(P._ptr(pos), int(pos.dtype == torch.float64), aos, P._ptr(wt),
                                               int(wt is not None and wt.dtype == torch.float64), Np, N, float(Lbox) if clip else 0.0,
                                               np.float32(float(N) / Lbox), np.float32(offset), N // world, world, P._ptr(dest),
                                               P._ptr(cursor), P._stream()), 'psb_slab_route_scatter_peer')
    bufs.barrier()                                       # every rank's copies have landed
    st.span('particles_peer_stores', e1)
    st.add_bytes('particles_peer_stores', 16 * int(M_[rank].sum() - M_[rank, rank]))
    recv = bufs.local.view(-1, 4)[:int(recv_tot[rank])]
    return recv, meta[:1].clone(), ntot


def assign_slab(pipe, xyzw, zbase, nzs, Lbox, offset=0.):
    """Step 2: K1 on routed particles onto the planes [zbase, zbase+nzs) only.  Returns mesh_slab [nzs, N, N, 2] float32."""
    N = pipe.N
    n = int(xyzw.shape[0])
    mesh = torch.empty((nzs, N, N, 2), dtype=torch.float32, device=pipe.dev)
    wsb = pipe.L.psb_assign_workspace_bytes(n, N)
    ws = torch.empty(wsb, dtype=torch.uint8, device=pipe.dev)
    scratch = torch.empty(1, dtype=torch.float64, device=pipe.dev)
    P.check(pipe.L.psb_assign_slab(P._ptr(xyzw), n, N, np.float32(float(N) / Lbox), np.float32(offset), zbase, nzs,
                                   P._ptr(mesh), 1, P._ptr(ws), wsb, P._ptr(scratch), P._stream()), 'psb_assign_slab')
    return mesh


# ------------------------------------------------------------------------------------------ 3: slab FFT
def z_to_y_chunks(t, world):
    """t: [nz, N, hp, 2] (a rank's z-planes, all ky) -> send buffer [world, nz, ny, hp, 2]: chunk q = the ky range of rank q."""
    nz, N, hp, two = t.shape
    ny = N // world
    return t.view(nz, world, ny, hp, two).permute(1, 0, 2, 3, 4).contiguous()


def z_to_y_slabs(t, world, stats=None):
    """All-to-all of z-slabs into ky-slabs: [nz, N, hp, 2] on every rank -> [N, ny, hp, 2] (all z of this rank's ky range)."""
    nz, N, hp, two = t.shape
    send = z_to_y_chunks(t, world)
    if world == 1:
        return send.view(N, N, hp, two)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                   # recv[g] = rank g's planes (z = g*nz .. ) of my ky range
    _stats(stats).add_bytes('fft_all_to_all', send.numel() * send.element_size() * (world - 1) // world)
    return recv.view(world * nz, N // world, hp, two)


def slab_phase1(pipe, mesh_slab):
    """x and y passes of a z-slab [nz, N, N, 2] (destroyed) + separation -> (P, Q) each [nz, N, hp, 2]."""
    N = pipe.N
    nz = mesh_slab.shape[0]
    hp = (N // 2 + 2) // 2 * 2
    st = P._stream()
    P.check(pipe.L.psb_fft_slab_xy(P._ptr(mesh_slab), N, nz, 1, P._ptr(pipe.tw32), st), 'psb_fft_slab_xy')
    p = torch.empty((nz, N, hp, 2), dtype=torch.float32, device=pipe.dev)
    q = torch.empty_like(p)
    P.check(pipe.L.psb_slab_split_ab(P._ptr(mesh_slab), P._ptr(p), P._ptr(q), N, nz, hp, st), 'psb_slab_split_ab')
    return p, q


def slab_phase2(pipe, py, qy, ky0, sumw, periodic=1):
    """z passes of a ky-slab of P and Q [N, ny, hp, 2] (in place) + fcomb -> rows ky0.. of the half field, [N, ny, N/2+1, 2]."""
    N = pipe.N
    ny, hp = py.shape[1], py.shape[2]
    st = P._stream()
    for t in (py, qy):
        P.check(pipe.L.psb_fft_slab_z(P._ptr(t), N, ny, hp, 1, P._ptr(pipe.tw32), st), 'psb_fft_slab_z')
    half = torch.empty((N, ny, N // 2 + 1, 2), dtype=torch.float32, device=pipe.dev)
    P.check(pipe.L.psb_slab_fcomb(P._ptr(py), P._ptr(qy), P._ptr(half), N, ky0, ny, hp, P._ptr(pipe.rec), P._ptr(pipe.wk), P._ptr(sumw),
                                  periodic, st), 'psb_slab_fcomb')
    return half


def slab_mesh_to_delta_emulated(pipe, mesh, sumw, world, periodic=1):
    """The slab FFT for `world` emulated ranks on ONE device (the exchanges are slices): mesh [N,N,N,2] -> half field
    [N,N,N/2+1,2].  Used to validate the building blocks without a multi-GPU box."""
    N = pipe.N
    nz, hp = slab_geometry(N, world)
    pq = [slab_phase1(pipe, mesh[g * nz:(g + 1) * nz].clone()) for g in range(world)]
    half = torch.empty((N, N, N // 2 + 1, 2), dtype=torch.float32, device=pipe.dev)
    for r in range(world):
        py = torch.cat([z_to_y_chunks(pq[g][0], world)[r] for g in range(world)], dim=0)
        qy = torch.cat([z_to_y_chunks(pq[g][1], world)[r] for g in range(world)], dim=0)
        half[:, r * nz:(r + 1) * nz] = slab_phase2(pipe, py, qy, r * nz, sumw, periodic)
    return half


def slab_phase1_routed(pipe, mesh_slab, bufs, zbase):
    """x and y passes of a z-slab (destroyed) + separation, with the z-slab -> ky-slab exchange fused in: every (z, ky) row of P and Q
    is stored straight into the arrays [N, ny, hp] of the rank that owns ky (`bufs`: SlabBuffers with rows P, Q)."""
    N = pipe.N
    nz = mesh_slab.shape[0]
    hp = (N // 2 + 2) // 2 * 2
    st = P._stream()
    P.check(pipe.L.psb_fft_slab_xy(P._ptr(mesh_slab), N, nz, 1, P._ptr(pipe.tw32), st), 'psb_fft_slab_xy')
    P.check(pipe.L.psb_slab_split_ab_routed(P._ptr(mesh_slab), N, nz, hp, zbase, bufs.world, P._ptr(bufs.row_table(pipe.dev)), st),
            'psb_slab_split_ab_routed')


def slab_delta(pipe, mesh_slab, sumw, periodic=1, stats=None):
    """Step 3: a rank's mesh planes [nz,N,N,2] (destroyed) -> its rows ky in [rank*ny, (rank+1)*ny) of delta(k): [N, ny, N/2+1, 2]."""
    st = _stats(stats)
    world = _world()
    nz, hp = slab_geometry(pipe.N, world)
    bufs = routed_slabs(pipe.dev, world, _rank(), 2, pipe.N * nz * hp * 2) if world > 1 else None
    if bufs is not None:                                 # exchange fused into the separation kernel (peer stores)
        e0 = st.mark()
        bufs.barrier()                                   # every rank is done with the previous catalogue's P, Q
        slab_phase1_routed(pipe, mesh_slab, bufs, _rank() * nz)
        bufs.barrier()                                   # every rank's rows have landed
        st.span('fft_xy_peer_stores', e0)
        st.add_bytes('fft_xy_peer_stores', 2 * 8 * nz * pipe.N * hp * (world - 1) // world)
        py = bufs.local[0].view(pipe.N, nz, hp, 2)
        qy = bufs.local[1].view(pipe.N, nz, hp, 2)
        e0 = st.mark()
        half = slab_phase2(pipe, py, qy, _rank() * nz, sumw, periodic)
        st.span('fft_z_fcomb', e0)
        return half
    e0 = st.mark()
    p, q = slab_phase1(pipe, mesh_slab)
    st.span('fft_xy', e0)
    e0 = st.mark()
    py, qy = z_to_y_slabs(p, world, st), z_to_y_slabs(q, world, st)
    del p, q
    st.span('fft_all_to_all', e0)
    e0 = st.mark()
    half = slab_phase2(pipe, py, qy, _rank() * nz, sumw, periodic)
    st.span('fft_z_fcomb', e0)
    return half


def gather_ky_slabs(mine, world):
    """All-gather of the ky-slabs [N, ny, hx, 2] of the half field into the replicated [N, N, hx, 2] array."""
    if world == 1:
        return mine
    N, ny, hx, two = mine.shape
    gathered = torch.empty((world * N, ny, hx, two), dtype=mine.dtype, device=mine.device)     # rank-major concatenation
    dist.all_gather_into_tensor(gathered, mine.contiguous())
    return gathered.view(world, N, ny, hx, two).permute(1, 0, 2, 3, 4).reshape(N, world * ny, hx, two)


def sharded_delta_slab(pipe, xyz_local, w_local, Lbox, offset=0., clip=True, periodic=1, stats=None):
    """Steps 1-3.  Returns (this rank's ky-slab of delta(k) [N, ny, N/2+1, 2], ky0, global sum of weights (device), global N)."""
    st = _stats(stats)
    world, rank = _world(), _rank()
    nz, _ = slab_geometry(pipe.N, world)
    xyzw, sumw, ntot = route_particles(pipe, xyz_local, w_local, Lbox, offset, clip, st)
    e0 = st.mark()
    mesh = assign_slab(pipe, xyzw, rank * nz, nz, Lbox, offset)
    st.span('assign_slab', e0)
    del xyzw
    half = slab_delta(pipe, mesh, sumw, periodic, st)
    return half, rank * nz, sumw, ntot


# ------------------------------------------------------------------------------------------ 4: spectra on ky-slabs
def slab_pk_monopole(pipe, half_slab, ky0, Lbox, stats=None):
    Nbins = pipe.N // 2
    out = torch.empty(3 * Nbins, dtype=torch.float64, device=pipe.dev)
    kf = 2 * np.pi / float(Lbox)
    P.check(pipe.L.psb_pk_monopole_slab(P._ptr(half_slab), pipe.N, ky0, half_slab.shape[1], P._ptr(pipe.pk_bin_table(Lbox)), Nbins, kf,
                                        P._ptr(out), P._stream()), 'psb_pk_monopole_slab')
    _allreduce(out)
    _stats(stats).add_bytes('pk_all_reduce', out.numel() * 8)
    return out


def slab_pk_multipoles(pipe, half_slab, ky0, Lbox_int, rsd, Nmubin, stats=None):
    Nbins = pipe.N // 2
    out = torch.empty((5 + 4 * Nmubin) * Nbins, dtype=torch.float64, device=pipe.dev)
    trig = np.empty(4, np.float32)
    P.check(pipe.L.psb_rsd_trig(int(rsd), P._np_ptr(trig)), 'psb_rsd_trig')
    kf32 = np.float32(np.float32(2.) * np.float32(3.141592654)) / np.float32(Lbox_int)
    P.check(pipe.L.psb_pk_multipoles_slab(P._ptr(half_slab), pipe.N, ky0, half_slab.shape[1], P._ptr(pipe.rsd_bin_table(Nbins)), Nbins,
                                          int(Nmubin), kf32, P._np_ptr(trig), P._ptr(out), P._stream()), 'psb_pk_multipoles_slab')
    _allreduce(out)
    _stats(stats).add_bytes('pk_all_reduce', out.numel() * 8)
    return out, kf32


# ------------------------------------------------------------------------------------------ 5: carrier grid
def carrier_grid(N, step, Nmax, Ncut):
    """Smallest compiled grid on which EVERY triangle of the configuration closes without a wrap (and every shell fits), else N:
    the grid the replicated low-k copy of delta(k) lives on."""
    tri = P.triangle_list(Nmax, Ncut, step)
    R = P.shell_reach(tri, step)
    need = int(max(R.sum(axis=1).max(), 2 * R.max()))
    for c in sorted(set(P.COARSE_GRIDS) | {360, 512}):
        if need < c < N:
            return int(c)
    return int(N)


def low_k_carrier(pipe, half_slab, ky0, Ng, stats=None):
    """Step 5: replicated half field on the carrier grid Ng from the ranks' ky-slabs (zero outside |k_a| < Ng/2 when Ng < N)."""
    N = pipe.N
    st = _stats(stats)
    if Ng == N and _world() > 1:                         # the carrier IS delta(k): an all-gather of the ky-slabs moves 1/G of the bytes
        e0 = st.mark()
        car = gather_ky_slabs(half_slab, _world()).contiguous()
        st.span('carrier_all_gather', e0)
        st.add_bytes('carrier_all_gather', half_slab.numel() * 4 * (_world() - 1))
        return car
    car = torch.empty((Ng, Ng, Ng // 2 + 1, 2), dtype=torch.float32, device=pipe.dev)
    P.check(pipe.L.psb_half_extract(P._ptr(half_slab), N, ky0, half_slab.shape[1], P._ptr(car), Ng, P._stream()), 'psb_half_extract')
    e0 = st.mark()
    _allreduce(car)
    st.span('carrier_all_reduce', e0)
    if _world() > 1:
        st.add_bytes('carrier_all_reduce', car.numel() * 4)
    return car


# ------------------------------------------------------------------------------------------ 6-7: sharded shell / triangle stage
def sharded_triangle_sums(pipe, half, step, Ncut, Nmax, dtype=torch.float32, stats=None):
    """Steps 6-7 on a replicated half field `half` of `pipe`'s grid (the carrier).  Returns host arrays (sum_x I_i I_j I_l per
    triangle, sum_x I_j^2 per shell) in the units of `pipe`'s grid, identical on every rank.  half=None with float64: the exact
    counts (delta == 1)."""
    st = _stats(stats)
    world, rank = _world(), _rank()
    s0 = Ncut // step
    S = Nmax - s0 + 1
    SA = S + (S % 2)
    f64 = dtype == torch.float64
    tri, levels = pipe.bk_levels(step, Ncut, Nmax)
    for pc, _, _, _ in levels:                           # every rank checks every level before the first collective
        check_cell_slabs(pc.N ** 3, world, packed=not f64)
    scales = None
    if not f64:
        scales = pipe.shell_scales(half, step, s0, Nmax)
        if world > 1:
            dist.broadcast(scales, 0)                    # the sampled shell power uses unordered atomics: make ranks agree exactly
    sums = torch.zeros(len(tri), dtype=torch.float64, device=pipe.dev)
    glob = torch.zeros((3, SA), dtype=torch.float64, device=pipe.dev)       # sum_x I^2, scale, max|I| per shell (owner rank writes)
    have = 0
    keep = []
    for pc, idx, idx_dev, smax in levels:
        Sl = smax - s0 + 1
        npairs = (Sl + 1) // 2
        deal, per = pair_assignment(npairs, world)
        mine = deal[rank]
        vol = (float(pipe.N) / pc.N) ** 3
        src = None if pc is pipe else pipe
        slab = pc.N ** 3 // world
        bufs = None
        if not f64 and pc.N % world == 0 and pc.N in ROUTED_GRIDS:      # whole planes per rank: K5 stores into the owners' slab buffers
            bufs = routed_slabs(pipe.dev, world, rank, world * 2 * per, slab)
        e0 = st.mark()
        if bufs is not None:
            bufs.barrier()                               # every rank is done with the previous contents of the buffers
            _, sq, sc_rows, mx = pc.shell_fields(half, step, s0, smax, scaled=True, pairs=mine, scales=scales, src=src,
                                                 routed=(bufs.route(per, mine, Sl, pipe.dev), pc.N // world, world))
            bufs.barrier()                               # every rank's planes have landed
            slabs = bufs.local
            st.span('shell_fields_peer_stores', e0)
            st.add_bytes('shell_fields_peer_stores', 4 * sum(1 for q in mine for e in (0, 1) if 2 * q + e < Sl) * slab * (world - 1))
        else:
            if f64:
                fields, sq = pc.shell_fields(half, step, s0, smax, dtype=torch.float64, pairs=mine, src=src)
                sc_rows = mx = None
            else:
                fields, sq, sc_rows, mx = pc.shell_fields(half, step, s0, smax, scaled=True, pairs=mine, scales=scales, src=src)
            st.span('shell_fields', e0)
            e0 = st.mark()
            slabs = exchange_to_slabs(fields, world, st)
            del fields
            st.span('shell_fields_all_to_all', e0)
        rows = slab_field_rows(Sl, world, per)
        use_tc = (not f64) and slabs.shape[1] % 64 == 0 and Sl <= 128
        e0 = st.mark()
        sl = pc.triangle_sums(slabs, smax, Ncut, step, engine='tc' if use_tc else 'fma', field_rows=rows, packed=not f64, tri=tri[idx])
        st.span('triangles', e0)
        keep.append((pc, slabs, rows, idx_dev, smax, vol, use_tc))
        sums.index_copy_(0, idx_dev, sl if vol == 1.0 else sl * vol)
        # per-shell quantities live on the owner rank of the pair: three indexed updates per level instead of one tiny kernel per shell
        loc = [(2 * k + e, 2 * pidx + e) for k, pidx in enumerate(mine) for e in (0, 1) if 2 * pidx + e < Sl]
        if loc:
            idx_t = lambda v: torch.tensor(v, dtype=torch.long, device=pipe.dev)
            src_i, dst_i = idx_t([a for a, _ in loc]), idx_t([b for _, b in loc])
            fresh = [(a, b) for a, b in loc if b >= have]    # shell power from the coarsest level that holds the shell (Parseval)
            if fresh:
                glob[0].index_copy_(0, idx_t([b for _, b in fresh]), sq[idx_t([a for a, _ in fresh])] * vol)
            if not f64:
                glob[1].index_copy_(0, dst_i, sc_rows[src_i].double())
                glob[2].index_copy_(0, dst_i, torch.maximum(glob[2][dst_i], mx.view(torch.float32)[src_i].double()))
        have = max(have, Sl)
    e0 = st.mark()
    flat = torch.cat([sums, glob[0], glob[2]])
    _allreduce(flat)
    st.span('sums_all_reduce', e0)
    if world > 1:
        st.add_bytes('sums_all_reduce', flat.numel() * 8)
    nt = len(tri)
    host = (flat if f64 else torch.cat([flat, scales.double()[:SA]])).cpu().numpy()      # one device -> host copy
    sums_h, sq_h, mx_h = host[:nt], host[nt:nt + SA], host[nt + SA:nt + 2 * SA]
    if f64:
        return sums_h, sq_h[:S]
    sc = host[nt + 2 * SA:]
    if mx_h.max() ** 2 >= 4.0e4:                         # fp16 range guard (pathological catalogues): redo with the FFMA kernel
        sums.zero_()
        for pc, slabs, rows, idx_dev, smax, vol, _ in keep:
            sl = pc.triangle_sums(slabs, smax, Ncut, step, engine='fma', field_rows=rows, packed=True, tri=tri[idx_dev.cpu().numpy()])
            sums.index_copy_(0, idx_dev, sl * vol)
        _allreduce(sums)
        sums_h = sums.cpu().numpy()
    del keep
    sums_h = sums_h / (sc[tri[:, 0] - s0] * sc[tri[:, 1] - s0] * sc[tri[:, 2] - s0])
    return sums_h, (sq_h / sc ** 2)[:S]


def sharded_counts(pipe, step, Ncut, Nmax):
    """Exact triangle counts with the float64 fields sharded over the ranks, in the units of `pipe`'s grid (N^3 * closed triples)."""
    key = (Nmax, Ncut, step)
    if key in pipe._counts:
        return pipe._counts[key]
    tri, levels = pipe.bk_levels(step, Ncut, Nmax)
    sums, _ = sharded_triangle_sums(pipe, None, step, Ncut, Nmax, dtype=torch.float64)
    n3 = float(pipe.N) ** 3
    nint = np.rint(sums / n3)
    if np.abs(sums / n3 - nint).max() > 1e-3:
        raise RuntimeError('triangle counts did not come out as integers')
    counts = np.zeros((Nmax, Nmax, Nmax), dtype=np.float64)
    counts[tri[:, 0] - 1, tri[:, 1] - 1, tri[:, 2] - 1] = nint * n3
    pipe._counts[key] = counts
    return counts


# ------------------------------------------------------------------------------------------ public API
def Bk_periodic_sharded(xyz_local, w_local=None, Lbox=2600, Ngrid=360, step=3, Ncut=3, Nmax=40, silent=True, stats=None,
                        return_pk=False):
    """Bk_periodic (pyspectrum.py:285-356) for ONE catalogue whose particles are spread over the ranks; every rank passes its own
    shard (any split) and receives the full result dictionary.  return_pk=True also returns Pk_periodic's dictionary from the same
    delta(k) (BASELINE configs[4] asks for both)."""
    pipe = P.PeriodicPipeline.get(Ngrid)
    s0 = Ncut // step
    if s0 < 1:
        raise ValueError('Ncut//step must be >= 1')
    st = _stats(stats)
    half, ky0, sumw, N = sharded_delta_slab(pipe, xyz_local, w_local, Lbox, stats=st)
    pk_dev = _pk_launch(pipe, half, ky0, Lbox, st) if return_pk else None       # collected after the B(k) stage is queued
    Ng = carrier_grid(Ngrid, step, Nmax, Ncut)
    car = low_k_carrier(pipe, half, ky0, Ng, st)
    del half
    pg = P.PeriodicPipeline.get(Ng)
    Nk = pg.shell_mode_counts(step, Nmax)                # shells lie inside the carrier: the same mode counts as on Ngrid
    counts_g = sharded_counts(pg, step, Ncut, Nmax)      # N_g^3 * closed triples; the epilogue below works in the carrier's units
    sums_h, sumsq_h = sharded_triangle_sums(pg, car, step, Ncut, Nmax, stats=st)
    tri = P.triangle_list(Nmax, Ncut, step)
    pk = _pk_finish(pipe, pk_dev, Lbox, N, w_local, sumw) if return_pk else None
    nbar = (float(N) if w_local is None else float(sumw.item())) / Lbox ** 3
    kf = 2 * np.pi / Lbox
    bispec = P._bk_epilogue(Ng, tri, sums_h, sumsq_h, Nk, counts_g, step, Ncut, Nmax)
    bispec['meta'] = {'Lbox': Lbox, 'Ngrid': Ngrid, 'step': step, 'Ncut': Ncut, 'Nmax': Nmax, 'N': N, 'nbar': nbar, 'kf': kf}
    for k in ('p0k1', 'p0k2', 'p0k3'):
        bispec[k] = bispec[k] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k_sn'] = 1. / nbar
    b_sn = (bispec['p0k1'] + bispec['p0k2'] + bispec['p0k3']) / nbar + 1. / nbar ** 2
    bispec['b123'] = bispec['b123'] * (2 * np.pi) ** 6 / kf ** 6 - b_sn
    bispec['b123_sn'] = b_sn
    with np.errstate(divide='ignore', invalid='ignore'):
        bispec['q123'] = bispec['b123'] / (bispec['p0k1'] * bispec['p0k2'] + bispec['p0k1'] * bispec['p0k3'] + bispec['p0k2'] * bispec['p0k3'])
    return (bispec, pk) if return_pk else bispec


def _pk_from_slab(pipe, half, ky0, Lbox, N, w_local, sumw, stats):
    return _pk_finish(pipe, _pk_launch(pipe, half, ky0, Lbox, stats), Lbox, N, w_local, sumw)


def _pk_launch(pipe, half, ky0, Lbox, stats):
    """K4 on the ky-slab + the all-reduce of the bins; the result stays on the device (no host wait in the middle of the pipeline)."""
    st = _stats(stats)
    e0 = st.mark()
    out = slab_pk_monopole(pipe, half, ky0, Lbox, st)
    st.span('pk_binning', e0)
    return out


def _pk_finish(pipe, out_dev, Lbox, N, w_local, sumw):
    out = out_dev.cpu().numpy()
    nbar = (float(N) if w_local is None else float(sumw.item())) / Lbox ** 3
    kf = 2 * np.pi / float(Lbox)
    Nbins = pipe.N // 2
    nk, ksum, psum = out[:Nbins], out[Nbins:2 * Nbins], out[2 * Nbins:]
    k = np.zeros(Nbins); p0k = np.zeros(Nbins); cnt = np.zeros(Nbins)
    ok = nk > 0
    k[ok] = ksum[ok] / nk[ok]
    p0k[ok] = psum[ok] / nk[ok] / kf ** 3
    cnt[ok] = nk[ok]
    p0k *= (2. * np.pi) ** 3
    meta = {'Lbox': Lbox, 'Ngrid': pipe.N, 'N': N, 'nbar': nbar, 'kf': 2 * np.pi / Lbox}
    return {'meta': meta, 'k': k, 'p0k': p0k - 1. / nbar, 'counts': cnt, 'p0k_sn': 1. / nbar}


def Pk_periodic_sharded(xyz_local, w_local=None, Lbox=2600, Ngrid=360, silent=True, stats=None):
    """Pk_periodic (pyspectrum.py:644-728) for one catalogue sharded over the ranks: slab assignment, slab FFT, binning on the
    ky-slabs and an all-reduce of the bins."""
    pipe = P.PeriodicPipeline.get(Ngrid)
    st = _stats(stats)
    half, ky0, sumw, N = sharded_delta_slab(pipe, xyz_local, w_local, Lbox, stats=st)
    return _pk_from_slab(pipe, half, ky0, Lbox, N, w_local, sumw, st)


def Pk_periodic_rsd_sharded(xyz_local, w_local=None, Lbox=2600, Ngrid=360, rsd=2, Nmubin=10, silent=True, stats=None):
    """Pk_periodic_rsd (pyspectrum.py:460-538) for one catalogue sharded over the ranks."""
    pipe = P.PeriodicPipeline.get(Ngrid)
    st = _stats(stats)
    half, ky0, sumw, N = sharded_delta_slab(pipe, xyz_local, w_local, Lbox, stats=st)
    raw, kf32 = slab_pk_multipoles(pipe, half, ky0, int(Lbox), rsd, Nmubin, st)
    ks, p0k, p2k, p4k, nk, k_kmu, mu_kmu, p_kmu, n_kmu = P._pk_rsd_normalise(raw.cpu().numpy(), kf32, pipe.N // 2, Nmubin)
    nbar = float(N) / Lbox ** 3
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'N': N, 'nbar': nbar, 'kf': 2 * np.pi / Lbox}
    return {'meta': meta, 'k': ks, 'p0k': p0k - 1. / nbar, 'p2k': p2k, 'p4k': p4k, 'p_sn': np.repeat(1. / nbar, len(ks)), 'counts': nk,
            'k_kmu': k_kmu, 'mu_kmu': mu_kmu, 'p_kmu': p_kmu - 1. / nbar, 'counts_kmu': n_kmu}
