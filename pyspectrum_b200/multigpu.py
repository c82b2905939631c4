"""pyspectrum_b200.multigpu -- ONE catalogue sharded over the GPUs of a node (one process per GPU, torchrun + NCCL).

Design ("alternative for grids that fit one GPU's memory", and what makes Ngrid=1024 fit at all):

  1. particles are sharded arbitrarily over the ranks; every rank assigns its shard onto a FULL mesh (K1)
     -> NCCL all-reduce (sum) of the mesh and of sum(w)                                   [exchange 1: 8 N^3 bytes]
  2. every rank runs the FFT + fcomb (K2+K3) on the reduced mesh: identical delta(k) half field everywhere
  3. the packed shell PAIRS are dealt round-robin to the ranks; each rank transforms only its pairs (K5)
  4. NCCL all-to-all: every shell field is cut into G slabs of N^3/G cells, slab q goes to rank q
     -> every rank holds all shells on its slab                                          [exchange 2: 4 S N^3 (G-1)/G bytes]
  5. slab-local triangle sums (K6) and shell powers -> all-reduce of Ntri + S float64   [exchange 3: ~50-400 kB]
  6. the float64 exact triangle counts use the same steps 3-5 with delta == 1

Steps 1-2 have a slab-decomposed alternative (`sharded_delta(..., fft='slab')` or PSB_SHARDED_FFT=slab): the mesh is
reduce-scattered into z-slabs instead of all-reduced, every rank transforms its slab along x and y, separates the two
interlaced grids' spectra inside each plane (P = A^xy, Q = B^xy: the conjugate partner (-kx,-ky) is local), one all-to-all turns
z-slabs into ky-slabs, the z pass and the point-wise fcomb follow, and an all-gather rebuilds the replicated half field that
steps 3-5 use.  No conjugate-partner exchange; the FFT work is divided by G instead of replicated.  Validated on one GPU with
emulated ranks (tests/test_gpu_slab.py: 2e-6 of max|delta| against the single-GPU K2+K3 for G = 1..8), the exchange bookkeeping
on gloo (tests/test_multigpu_host.py), and under NCCL on 2 GPUs at C2 (tools/run_sharded.py: same outputs as the single-GPU
path, 21.4 vs 20.9 ms -- at 360^3 on two GPUs the extra exchanges cost what the halved FFT saves).  It is meant for 1024^3 on 8
GPUs, where the 8.6 GB mesh all-reduce and the replicated 25 ms FFT dominate P(k); the default stays 'replicated' until that
configuration has been timed (profiles/r1_multigpu.jsonl).

Every kernel is the single-GPU one; only the orchestration differs.  The collectives are torch.distributed calls so the
same code runs under NCCL (GPU) and gloo (the CPU tests of the host logic: tests/test_multigpu_host.py)."""
import numpy as np
import torch
import torch.distributed as dist

from . import pyspectrum as P


def pair_assignment(npairs, world):
    """Round-robin deal of packed shell pairs; every rank gets the same number (padding pairs = empty shells)."""
    per = (npairs + world - 1) // world
    return [[r + world * k for k in range(per)] for r in range(world)], per


def slab_field_rows(S, world, per):
    """After the all-to-all the slab buffer is laid out [k][owner rank][e] (k = local pair index, e = 0/1 within the pair);
    returns for every shell slot f the row of that buffer which holds it."""
    rows = []
    for f in range(S):
        p, e = divmod(f, 2)
        owner, k = p % world, p // world
        rows.append((k * world + owner) * 2 + e)
    return rows


def exchange_to_slabs(fields_local, world):
    """fields_local: [2*per, ncell] on every rank (its shell pairs, full grid).  Returns [2*per*world, ncell/world]:
    row ((k*world + owner)*2 + e) = slab of this rank of shell e of the k-th pair of rank `owner`."""
    nrow, ncell = fields_local.shape
    slab = ncell // world
    per = nrow // 2
    out = torch.empty((per, world, 2, slab), dtype=fields_local.dtype, device=fields_local.device)
    if world == 1:
        out.copy_(fields_local.view(per, 2, 1, slab).permute(0, 2, 1, 3))
        return out.view(nrow, slab)
    for k in range(per):
        for e in range(2):
            src = fields_local[2 * k + e].view(world, slab)                 # contiguous: slab q -> rank q
            dst = torch.empty((world, slab), dtype=fields_local.dtype, device=fields_local.device)
            dist.all_to_all_single(dst, src)
            out[k, :, e, :] = dst                                            # dst[q'] = my slab of rank q''s row (k, e)
    return out.view(nrow * world, slab)


# ------------------------------------------------------------------------------------------ slab-decomposed mesh -> delta(k)
def slab_geometry(N, world):
    """(planes per rank, padded half-row length): N must divide by the number of ranks; rows hold kx = 0..N/2 padded to even."""
    if N % world:
        raise ValueError('Ngrid must be a multiple of the number of ranks for the slab FFT')
    return N // world, (N // 2 + 2) // 2 * 2


def z_to_y_chunks(t, world):
    """t: [nz, N, hp, 2] (a rank's z-planes, all ky) -> send buffer [world, nz, ny, hp, 2]: chunk q = the ky range of rank q."""
    nz, N, hp, two = t.shape
    ny = N // world
    return t.view(nz, world, ny, hp, two).permute(1, 0, 2, 3, 4).contiguous()


def z_to_y_slabs(t, world):
    """All-to-all of z-slabs into ky-slabs: [nz, N, hp, 2] on every rank -> [N, ny, hp, 2] (all z of this rank's ky range)."""
    nz, N, hp, two = t.shape
    send = z_to_y_chunks(t, world)
    if world == 1:
        return send.view(N, N, hp, two)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                   # recv[g] = rank g's planes (z = g*nz .. ) of my ky range
    return recv.view(world * nz, N // world, hp, two)


def slab_phase1(pipe, mesh_slab):
    """x and y passes of a reduced z-slab [nz, N, N, 2] (destroyed) + separation -> (P, Q) each [nz, N, hp, 2]."""
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
    """The slab pipeline for `world` emulated ranks on ONE device (the exchanges are slices): mesh [N,N,N,2] -> half field
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


def slab_mesh_to_delta(pipe, mesh, sumw, periodic=1):
    """Distributed slab pipeline: every rank passes its UNREDUCED full mesh [N,N,N,2] (destroyed); returns the replicated half
    field.  reduce-scatter (z-slabs) -> x,y passes + separation -> all-to-all -> z pass + fcomb -> all-gather."""
    world = _world()
    N = pipe.N
    nz, hp = slab_geometry(N, world)
    rank = dist.get_rank() if world > 1 else 0
    if world > 1:
        slab = torch.empty((nz, N, N, 2), dtype=torch.float32, device=pipe.dev)
        dist.reduce_scatter_tensor(slab, mesh)
    else:
        slab = mesh
    p, q = slab_phase1(pipe, slab)
    del slab
    py, qy = z_to_y_slabs(p, world), z_to_y_slabs(q, world)
    del p, q
    mine = slab_phase2(pipe, py, qy, rank * nz, sumw, periodic)
    return gather_ky_slabs(mine, world)


def gather_ky_slabs(mine, world):
    """All-gather of the ky-slabs [N, ny, hx, 2] of the half field into the replicated [N, N, hx, 2] array."""
    if world == 1:
        return mine
    N, ny, hx, two = mine.shape
    gathered = torch.empty((world * N, ny, hx, two), dtype=mine.dtype, device=mine.device)     # rank-major concatenation
    dist.all_gather_into_tensor(gathered, mine.contiguous())
    return gathered.view(world, N, ny, hx, two).permute(1, 0, 2, 3, 4).reshape(N, world * ny, hx, two)


def _allreduce(t, op=dist.ReduceOp.SUM):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op)
    return t


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def sharded_delta(pipe, xyz_local, w_local, Lbox, fft=None):
    """Steps 1-2: returns (half field, global sum of weights as a device tensor).  fft = 'replicated' (all-reduce of the mesh,
    every rank transforms all of it) or 'slab' (see the module docstring); default from PSB_SHARDED_FFT, else 'replicated'."""
    import os
    fft = fft or os.environ.get('PSB_SHARDED_FFT', 'replicated')
    if fft not in ('replicated', 'slab'):
        raise ValueError("fft must be 'replicated' or 'slab'")
    pos, aos, wt = pipe.to_device(xyz_local, w_local)
    mesh, sumw = pipe.assign(pos, aos, wt, Lbox)
    _allreduce(sumw)
    if fft == 'slab' and pipe.N % _world() == 0:
        return slab_mesh_to_delta(pipe, mesh, sumw), sumw
    _allreduce(mesh)
    half = pipe.mesh_to_delta(mesh, sumw)
    return half, sumw


def sharded_triangle_sums(pipe, half, step, Ncut, Nmax, dtype=torch.float32):
    """Steps 3-5.  Returns host arrays (sums per triangle, sum_x I_j^2 per shell) in the reference's units, identical
    on every rank."""
    world = _world()
    rank = dist.get_rank() if world > 1 else 0
    s0 = Ncut // step
    S = Nmax - s0 + 1
    npairs = (S + 1) // 2
    assign, per = pair_assignment(npairs, world)
    mine = assign[rank]
    f64 = dtype == torch.float64
    if f64:
        fields, sumsq = pipe.shell_fields(half, step, s0, Nmax, dtype=torch.float64, pairs=mine)
        sc_rows = None
    else:
        scales = pipe.shell_scales(half, step, s0, Nmax)
        if world > 1:
            dist.broadcast(scales, 0)                  # the sampled shell power uses unordered atomics: make ranks agree exactly
        fields, sumsq, sc_rows, maxabs = pipe.shell_fields(half, step, s0, Nmax, scaled=True, pairs=mine, scales=scales)
    slabs = exchange_to_slabs(fields, world)
    del fields
    rows = slab_field_rows(S, world, per)
    engine = 'fma' if f64 else 'auto'
    sums = pipe.triangle_sums(slabs, Nmax, Ncut, step, engine=engine, field_rows=rows, packed=not f64)
    _allreduce(sums)
    # per-shell quantities live on the owner rank: scatter them into global shell order, then sum over ranks
    glob_sq = torch.zeros(2 * npairs, dtype=torch.float64, device=pipe.dev)
    glob_sc = torch.zeros(2 * npairs, dtype=torch.float64, device=pipe.dev)
    glob_mx = torch.zeros(2 * npairs, dtype=torch.float64, device=pipe.dev)
    for k, pidx in enumerate(mine):
        if pidx < npairs:
            glob_sq[2 * pidx:2 * pidx + 2] = sumsq[2 * k:2 * k + 2]
            if not f64:
                glob_sc[2 * pidx:2 * pidx + 2] = sc_rows[2 * k:2 * k + 2].double()
                glob_mx[2 * pidx:2 * pidx + 2] = maxabs[2 * k:2 * k + 2].view(torch.float32).double()
    _allreduce(glob_sq)
    if f64:
        return sums.cpu().numpy(), glob_sq.cpu().numpy()[:S]
    _allreduce(glob_sc)
    _allreduce(glob_mx)
    host = torch.cat([sums, glob_sq, glob_sc, glob_mx]).cpu().numpy()
    nt = sums.numel()
    n2 = 2 * npairs
    sums_h, sq, sc, mx = host[:nt], host[nt:nt + n2], host[nt + n2:nt + 2 * n2], host[nt + 2 * n2:]
    if mx.max() ** 2 >= 4.0e4:                                           # fp16 range guard: redo with the FFMA kernel
        sums = pipe.triangle_sums(slabs, Nmax, Ncut, step, engine='fma', field_rows=rows, packed=True)
        _allreduce(sums)
        sums_h = sums.cpu().numpy()
    tri = P.triangle_list(Nmax, Ncut, step)
    sums_h = sums_h / (sc[tri[:, 0] - s0] * sc[tri[:, 1] - s0] * sc[tri[:, 2] - s0])
    return sums_h, (sq / sc ** 2)[:S]


def sharded_counts(pipe, step, Ncut, Nmax):
    """Exact triangle counts with the float64 fields sharded over the ranks (at Ngrid=1024 they do not fit one GPU)."""
    key = (Nmax, Ncut, step)
    if key in pipe._counts:
        return pipe._counts[key]
    sums, _ = sharded_triangle_sums(pipe, None, step, Ncut, Nmax, dtype=torch.float64)
    tri = P.triangle_list(Nmax, Ncut, step)
    n3 = float(pipe.N) ** 3
    nint = np.rint(sums / n3)
    if np.abs(sums / n3 - nint).max() > 1e-3:
        raise RuntimeError('triangle counts did not come out as integers')
    counts = np.zeros((Nmax, Nmax, Nmax), dtype=np.float64)
    counts[tri[:, 0] - 1, tri[:, 1] - 1, tri[:, 2] - 1] = nint * n3
    pipe._counts[key] = counts
    return counts


def Bk_periodic_sharded(xyz_local, w_local=None, Lbox=2600, Ngrid=360, step=3, Ncut=3, Nmax=40, silent=True):
    """Bk_periodic (pyspectrum.py:285-356) for ONE catalogue whose particles are spread over the ranks; every rank
    passes its own shard and receives the full result dictionary."""
    pipe = P.PeriodicPipeline.get(Ngrid)
    s0 = Ncut // step
    if s0 < 1:
        raise ValueError('Ncut//step must be >= 1')
    Nloc = torch.tensor([float(xyz_local.shape[1])], dtype=torch.float64, device=pipe.dev)
    _allreduce(Nloc)
    N = int(Nloc.item())
    half, sumw = sharded_delta(pipe, xyz_local, w_local, Lbox)
    Nk = pipe.shell_mode_counts(step, Nmax)
    counts = sharded_counts(pipe, step, Ncut, Nmax)
    sums_h, sumsq_h = sharded_triangle_sums(pipe, half, step, Ncut, Nmax)
    tri = P.triangle_list(Nmax, Ncut, step)
    nbar = (float(N) if w_local is None else float(sumw.item())) / Lbox ** 3
    kf = 2 * np.pi / Lbox
    bispec = P._bk_epilogue(Ngrid, tri, sums_h, sumsq_h, Nk, counts, step, Ncut, Nmax)
    bispec['meta'] = {'Lbox': Lbox, 'Ngrid': Ngrid, 'step': step, 'Ncut': Ncut, 'Nmax': Nmax, 'N': N, 'nbar': nbar, 'kf': kf}
    for k in ('p0k1', 'p0k2', 'p0k3'):
        bispec[k] = bispec[k] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar
    bispec['p0k_sn'] = 1. / nbar
    b_sn = (bispec['p0k1'] + bispec['p0k2'] + bispec['p0k3']) / nbar + 1. / nbar ** 2
    bispec['b123'] = bispec['b123'] * (2 * np.pi) ** 6 / kf ** 6 - b_sn
    bispec['b123_sn'] = b_sn
    with np.errstate(divide='ignore', invalid='ignore'):
        bispec['q123'] = bispec['b123'] / (bispec['p0k1'] * bispec['p0k2'] + bispec['p0k1'] * bispec['p0k3'] + bispec['p0k2'] * bispec['p0k3'])
    return bispec


def Pk_periodic_sharded(xyz_local, w_local=None, Lbox=2600, Ngrid=360, silent=True):
    """Pk_periodic (pyspectrum.py:644-728) for one catalogue sharded over the ranks (binning is done redundantly: it is
    one pass over the half field)."""
    pipe = P.PeriodicPipeline.get(Ngrid)
    Nloc = torch.tensor([float(xyz_local.shape[1])], dtype=torch.float64, device=pipe.dev)
    _allreduce(Nloc)
    N = int(Nloc.item())
    half, sumw = sharded_delta(pipe, xyz_local, w_local, Lbox)
    out = pipe.pk_monopole(half, Lbox).cpu().numpy()
    nbar = (float(N) if w_local is None else float(sumw.item())) / Lbox ** 3
    kf = 2 * np.pi / float(Lbox)
    Nbins = Ngrid // 2
    nk, ksum, psum = out[:Nbins], out[Nbins:2 * Nbins], out[2 * Nbins:]
    k = np.zeros(Nbins); p0k = np.zeros(Nbins); cnt = np.zeros(Nbins)
    ok = nk > 0
    k[ok] = ksum[ok] / nk[ok]
    p0k[ok] = psum[ok] / nk[ok] / kf ** 3
    cnt[ok] = nk[ok]
    p0k *= (2. * np.pi) ** 3
    meta = {'Lbox': Lbox, 'Ngrid': Ngrid, 'N': N, 'nbar': nbar, 'kf': 2 * np.pi / Lbox}
    return {'meta': meta, 'k': k, 'p0k': p0k - 1. / nbar, 'counts': cnt, 'p0k_sn': 1. / nbar}
