"""GPU tests of the slab-decomposed mesh -> delta(k) building blocks of the sharded path (multigpu.slab_*) on ONE
device: the ranks are emulated and the exchanges are slices, so every kernel and the index bookkeeping of the distributed
version run exactly as they would under NCCL (whose collectives are covered on gloo in tests/test_multigpu_host.py).
Reference = the single-GPU K2+K3 (psb_fft_mesh_to_delta), itself pinned to the oracle in tests/test_gpu_parity.py.
Tolerance: 2e-6 of max|delta| (the slab path rebuilds F(k) from the separated spectra: one more float32 rounding),
including the self-conjugate planes where the Fortran's last write wins."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pyspectrum_b200 import pyspectrum as pySpec, multigpu
    return pySpec, multigpu


def _mesh(pySpec, N, Np, L, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(0, L, (3, Np))
    xyz[:, :Np // 2] = (xyz[:, :Np // 2] * 0.25 + 0.3 * L) % L
    w = rng.uniform(0.5, 2., Np)
    pipe = pySpec.PeriodicPipeline.get(N)
    pos, aos, wt = pipe.to_device(xyz, w)
    mesh, sumw = pipe.assign(pos, aos, wt, L)
    return pipe, mesh, sumw


@pytest.mark.parametrize('N', [24, 36, 40, 64])
def test_slab_passes_compose_to_the_3d_transform(mods, N):
    import torch
    pySpec, M = mods
    pipe = pySpec.PeriodicPipeline.get(N)
    x = torch.randn((N, N, N, 2), device='cuda', dtype=torch.float32)
    ref = x.clone()
    pySpec.check(pipe.L.psb_fft_c2c_3d(pySpec._ptr(ref), N, 1, pySpec._ptr(pipe.tw32), pySpec._stream()), 'psb_fft_c2c_3d')
    for nz in (N, N // 2, N // 4):                                   # slabs of different thickness
        a = x.clone()
        for z0 in range(0, N, nz):
            s = a[z0:z0 + nz]
            pySpec.check(pipe.L.psb_fft_slab_xy(pySpec._ptr(s), N, nz, 1, pySpec._ptr(pipe.tw32), pySpec._stream()), 'psb_fft_slab_xy')
        # z pass on ky-slabs [N][ny][N] cut out of the cube
        for y0 in range(0, N, nz):
            t = a[:, y0:y0 + nz].contiguous()
            pySpec.check(pipe.L.psb_fft_slab_z(pySpec._ptr(t), N, nz, N, 1, pySpec._ptr(pipe.tw32), pySpec._stream()), 'psb_fft_slab_z')
            a[:, y0:y0 + nz] = t
        assert (a - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize('N,world', [(24, 1), (24, 2), (24, 4), (36, 3), (40, 2), (64, 8), (360, 2)])
@pytest.mark.parametrize('periodic', [1, 0])
def test_slab_pipeline_matches_single_gpu_delta(mods, N, world, periodic):
    pySpec, M = mods
    if N == 360 and not periodic:
        pytest.skip('one large case is enough')
    pipe, mesh, sumw = _mesh(pySpec, N, 20000 if N < 360 else 2000000, 100., N + world)
    ref = pipe.mesh_to_delta(mesh.clone(), sumw, periodic=periodic)
    got = M.slab_mesh_to_delta_emulated(pipe, mesh, sumw, world, periodic=periodic)
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    assert err.max().item() <= 2e-6 * scale
    h = N // 2
    for sl in (err[:, :, 0], err[:, :, h], err[:, h], err[h]):      # the planes where images coincide (last write wins)
        assert sl.max().item() <= 2e-6 * scale


def test_slab_single_rank_distributed_entry_point(mods):
    pySpec, M = mods
    pipe, mesh, sumw = _mesh(pySpec, 32, 20000, 100., 5)
    ref = pipe.mesh_to_delta(mesh.clone(), sumw)
    got = M.slab_delta(pipe, mesh.clone(), sumw)                     # no process group: world = 1, the slab is the whole mesh
    assert (got - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    with pytest.raises(ValueError):
        M.slab_geometry(32, 5)


# ------------------------------------------------------------------------------------------ slab-owned assignment
def _catalogue(N, Np, L, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(0, L, (3, Np))
    xyz[:, :Np // 2] = (xyz[:, :Np // 2] * 0.25 + 0.3 * L) % L
    xyz[:, :6] = np.array([[0., L, -3., L * (1 - 1e-6), 0.5 * L, L + 2.],          # faces and outside the box: clip, wrap of the stencil
                           [0., 0.5 * L, L, 0., L * (1 - 1e-6), -1.],
                           [0., L * (1 - 1e-6), 0.25 * L, L + 5., -2., 0.5 * L]])
    return np.ascontiguousarray(xyz), rng.uniform(0.5, 2., Np)


@pytest.mark.parametrize('N,world,Np', [(32, 1, 20000), (32, 2, 20000), (32, 4, 30000), (64, 8, 100000), (360, 8, 1000000), (40, 5, 20000)])
@pytest.mark.parametrize('weighted', [True, False])
def test_routed_slab_assignment_equals_the_full_mesh(mods, N, world, Np, weighted):
    """Every emulated rank routes the catalogue (it is the only sender here), takes its own segment of the destination-major
    send buffer -- its planes' particles plus the ghost copies -- and assigns it onto its planes only: the slabs stacked are the
    single-GPU mesh (same float32 contributions, another summation order), and nothing outside a rank's planes is written."""
    import torch
    pySpec, M = mods
    if N == 360 and not weighted:
        pytest.skip('one large case is enough')
    L = 100.
    xyz, w = _catalogue(N, Np, L, N + world)
    w = w if weighted else None
    pipe = pySpec.PeriodicPipeline.get(N)
    pos, aos, wt = pipe.to_device(xyz, w)
    mesh, sumw = pipe.assign(pos, aos, wt, L)
    nz = N // world
    counts, sw = M.route_counts(pipe, pos, aos, wt, L, world)
    send = M.route_scatter(pipe, pos, aos, wt, L, world, counts)
    c = counts.cpu().numpy()
    assert Np <= c.sum() <= (2 * Np if world > 1 else Np) and abs(sw.item() - sumw.item()) <= 1e-9 * abs(sumw.item())
    base = np.concatenate([[0], np.cumsum(c)])
    scale = mesh.abs().max().item()
    for r in range(world):
        seg = send[base[r]:base[r + 1]].contiguous()
        slab = M.assign_slab(pipe, seg, r * nz, nz, L)
        assert slab.shape == (nz, N, N, 2)
        # the same adds in another float32 order (the sort places equal-key particles by atomic arrival): a few ulp of the largest cell
        assert (slab - mesh[r * nz:(r + 1) * nz]).abs().max().item() <= 4e-6 * scale
    # routed positions are the float32 values assign_quad receives (clip in float64, then the cast: py:938-941)
    ref32 = np.clip(xyz, 0., L * (1 - 1e-6)).astype(np.float32)
    got = send[:, :3].cpu().numpy()
    assert set(map(bytes, np.unique(got, axis=0))) <= set(map(bytes, np.unique(ref32.T, axis=0)))


@pytest.mark.parametrize('N,world', [(32, 2), (48, 4), (64, 8)])
def test_slab_spectra_partial_bins_sum_to_the_full_binning(mods, N, world):
    import torch
    pySpec, M = mods
    pipe, mesh, sumw = _mesh(pySpec, N, 30000, 100., 7 * N)
    half = pipe.mesh_to_delta(mesh, sumw)
    ny = N // world
    full = pipe.pk_monopole(half, 100.)
    fullm, _ = pipe.pk_multipoles(half, 100, 2, 5)
    acc, accm = torch.zeros_like(full), torch.zeros_like(fullm)
    for r in range(world):
        sl = half[:, r * ny:(r + 1) * ny].contiguous()
        acc += M.slab_pk_monopole(pipe, sl, r * ny, 100.)
        accm += M.slab_pk_multipoles(pipe, sl, r * ny, 100, 2, 5)[0]
    Nb = N // 2
    assert torch.equal(acc[:Nb], full[:Nb]) and torch.equal(accm[:Nb], fullm[:Nb])                 # mode counts: exact
    assert torch.equal(accm[5 * Nb:5 * Nb + 5 * Nb], fullm[5 * Nb:5 * Nb + 5 * Nb])                # (k,mu) counts: exact
    assert torch.allclose(acc, full, rtol=1e-12, atol=0) and torch.allclose(accm, fullm, rtol=1e-11, atol=1e-9 * fullm.abs().max().item())


@pytest.mark.parametrize('N,Ng,world', [(64, 32, 2), (64, 64, 4), (96, 48, 3), (128, 64, 8)])
def test_carrier_extraction(mods, N, Ng, world):
    import torch
    pySpec, M = mods
    pipe = pySpec.PeriodicPipeline.get(N)
    h, hg = N // 2, Ng // 2
    half = torch.randn((N, N, h + 1, 2), device='cuda', dtype=torch.float32)
    ny = N // world
    car = torch.zeros((Ng, Ng, hg + 1, 2), device='cuda', dtype=torch.float32)
    for r in range(world):
        car += M.low_k_carrier(pipe, half[:, r * ny:(r + 1) * ny].contiguous(), r * ny, Ng)
    a = half.cpu().numpy()
    ref = np.zeros((Ng, Ng, hg + 1, 2), np.float32)
    if Ng == N:
        ref[:] = a
    else:
        k = np.arange(-hg + 1, hg)
        ref[np.ix_(k % Ng, k % Ng, np.arange(hg))] = a[np.ix_(k % N, k % N, np.arange(hg))]
    assert np.array_equal(car.cpu().numpy(), ref)


def test_sharded_entry_points_on_one_rank(mods, monkeypatch, tmp_path):
    """World size 1 runs every step of the sharded path (route, slab assignment, slab FFT, slab binning, carrier, levels, dealt
    pairs, cell slabs) without a process group: same results as the single-GPU API."""
    pySpec, M = mods
    monkeypatch.setattr(pySpec, '_DAT_DIR', str(tmp_path))
    for N, L, Np, step, Ncut, Nmax in [(64, 500., 100000, 2, 3, 12), (512, 2600., 1000000, 3, 3, 20)]:
        xyz, w = _catalogue(N, Np, L, N)
        assert M.carrier_grid(N, step, Nmax, Ncut) == (N if N == 64 else 256)
        ref = pySpec.Bk_periodic(xyz, w=w, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
        st = M.Stats(timed=True)
        got, pk = M.Bk_periodic_sharded(xyz, w, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax, stats=st, return_pk=True)
        assert np.array_equal(got['i_k1'], ref['i_k1']) and np.array_equal(got['i_k3'], ref['i_k3'])
        assert np.array_equal(got['counts'], ref['counts'])
        assert got['meta']['N'] == Np and abs(got['meta']['nbar'] / ref['meta']['nbar'] - 1) < 1e-12
        np.testing.assert_allclose(got['p0k1'] + got['p0k_sn'], ref['p0k1'] + ref['p0k_sn'], rtol=1e-5)
        scale = np.abs(ref['b123'] + ref['b123_sn'])
        assert np.all(np.abs(got['b123'] - ref['b123']) <= 1e-5 * scale + 1e-7 * scale.max())
        rpk = pySpec.Pk_periodic(xyz, w=w, Lbox=L, Ngrid=N)
        assert np.array_equal(pk['counts'], rpk['counts'])
        np.testing.assert_allclose(pk['p0k'] + pk['p0k_sn'], rpk['p0k'] + rpk['p0k_sn'], rtol=1e-5)
        assert 'assign_slab' in st.times_ms() and 'triangles' in st.times_ms()
    xyz, w = _catalogue(64, 100000, 500., 3)
    a = M.Pk_periodic_rsd_sharded(xyz, None, Lbox=500., Ngrid=64, rsd=2, Nmubin=5)
    b = pySpec.Pk_periodic_rsd(xyz, Lbox=500., Ngrid=64, rsd=2, Nmubin=5)
    assert np.array_equal(a['counts'], b['counts']) and np.array_equal(a['counts_kmu'], b['counts_kmu'])
    np.testing.assert_allclose(a['p0k'] + a['p_sn'], b['p0k'] + b['p_sn'], rtol=1e-5)


@pytest.mark.parametrize('N,Ns,world,step,smax', [(64, 64, 2, 2, 13), (64, 64, 4, 2, 12), (48, 48, 3, 1, 9), (256, 360, 8, 3, 9), (360, 360, 4, 3, 6)])
def test_routed_shell_planes_land_where_the_all_to_all_puts_them(mods, N, Ns, world, step, smax):
    """K5 with routed output (psb_bk_shell_pair_f32_routed): every emulated rank transforms its dealt pairs and stores each plane
    through the route table into the slab buffer of the plane's owner (here: `world` buffers of one device standing in for the
    peers' symmetric memory).  The buffers must equal, bit for bit, what the local fields + the all-to-all deliver (same kernel,
    only the store address differs), shell powers and max|I| included; rows of padding pairs / missing shells stay untouched."""
    import torch
    pySpec, M = mods
    src = pySpec.PeriodicPipeline.get(Ns)
    pc = pySpec.PeriodicPipeline.get(N)
    h = Ns // 2
    g = torch.Generator(device='cuda').manual_seed(N + world)
    half = torch.randn((Ns, Ns, h + 1, 2), device='cuda', dtype=torch.float32, generator=g)
    s0 = 1
    Sl = smax - s0 + 1
    scales = src.shell_scales(half, step, s0, smax)
    deal, per = M.pair_assignment((Sl + 1) // 2, world)
    slab = N ** 3 // world
    M.check_cell_slabs(N ** 3, world)
    ranks = M.SlabBuffers.emulated(pc.dev, world, world * 2 * per, slab)
    for b in ranks:
        b.local.fill_(-7.0)
    other = None if Ns == N else src
    for r in range(world):
        table = ranks[r].route(per, deal[r], Sl, pc.dev)
        none, sq, sc, mx = pc.shell_fields(half, step, s0, smax, scaled=True, pairs=deal[r], scales=scales, src=other,
                                           routed=(table, N // world, world))
        assert none is None
        fields, sq_ref, sc_ref, mx_ref = pc.shell_fields(half, step, s0, smax, scaled=True, pairs=deal[r], scales=scales, src=other)
        assert torch.equal(sq, sq_ref) and torch.equal(mx, mx_ref) and torch.equal(sc, sc_ref)
        for k, pidx in enumerate(deal[r]):
            for e in (0, 1):
                row = r * 2 * per + 2 * k + e
                for q in range(world):
                    got = ranks[q].local[row].view(torch.int32)
                    if 2 * pidx + e < Sl:
                        assert torch.equal(got, fields[2 * k + e, q * slab:(q + 1) * slab].view(torch.int32)), (r, k, e, q)
                    else:
                        assert bool((ranks[q].local[row] == -7.0).all())
    rows = M.slab_field_rows(Sl, world, per)
    assert sorted(set(rows)) == sorted(rows) and max(rows) < world * 2 * per


@pytest.mark.parametrize('N,world', [(24, 2), (36, 3), (64, 4), (64, 8)])
def test_routed_slab_fft_exchange_equals_the_all_to_all(mods, N, world):
    """psb_slab_split_ab_routed: every emulated rank stores the (z, ky) rows of its separated spectra P, Q straight into the ky-slab
    arrays of the owning rank; the arrays must equal, bit for bit, what split + all-to-all chunks deliver."""
    import torch
    pySpec, M = mods
    pipe, mesh, sumw = _mesh(pySpec, N, 20000, 100., N + world)
    nz, hp = M.slab_geometry(N, world)
    ranks = M.SlabBuffers.emulated(pipe.dev, world, 2, N * nz * hp * 2)
    for b in ranks:
        b.local.fill_(float('nan'))
    pq = []
    for r in range(world):
        M.slab_phase1_routed(pipe, mesh[r * nz:(r + 1) * nz].clone(), ranks[r], r * nz)
        pq.append(M.slab_phase1(pipe, mesh[r * nz:(r + 1) * nz].clone()))
    for q in range(world):
        for k in (0, 1):
            want = torch.cat([M.z_to_y_chunks(pq[g][k], world)[q] for g in range(world)], dim=0)        # [N, ny, hp, 2]
            got = ranks[q].local[k].view(N, nz, hp, 2)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32)), (q, k)


@pytest.mark.parametrize('N,world,Np', [(32, 2, 20000), (64, 4, 100000), (64, 8, 300000), (360, 8, 1000000)])
def test_peer_store_route_scatter_delivers_the_all_to_all_segments(mods, N, world, Np):
    """psb_slab_route_scatter_peer with emulated ranks: every rank routes ITS shard of the catalogue and stores the copies straight
    into the receive buffers of their destinations (buffers of this device standing in for peer memory), at the segment offsets that
    follow from the all-gathered counts.  Every (source, destination) segment must hold exactly the particles the send-buffer
    scatter + all-to-all delivers (the order inside a segment is not defined in either path: compared as sorted byte rows)."""
    import torch
    pySpec, M = mods
    L = 100.
    xyz, w = _catalogue(N, Np, L, 3 * N + world)
    pipe = pySpec.PeriodicPipeline.get(N)
    shards = [(np.ascontiguousarray(xyz[:, r::world]), np.ascontiguousarray(w[r::world])) for r in range(world)]
    dev = [pipe.to_device(x, ww) for x, ww in shards]
    counts = [M.route_counts(pipe, pos, aos, wt, L, world)[0] for pos, aos, wt in dev]
    Mh = np.stack([c.cpu().numpy() for c in counts])                          # [source][destination]
    recv_tot = Mh.sum(axis=0)
    bufs = [torch.full((int(recv_tot[d]) + 8, 4), float('nan'), dtype=torch.float32, device=pipe.dev) for d in range(world)]
    for r, (pos, aos, wt) in enumerate(dev):
        seg = Mh[:r].sum(axis=0)
        dest = torch.tensor([bufs[d].data_ptr() + 16 * int(seg[d]) for d in range(world)], dtype=torch.int64, device=pipe.dev)
        cursor = torch.zeros(world, dtype=torch.int64, device=pipe.dev)
        npart = pos.shape[0] if aos else pos.shape[1]
        pySpec.check(pipe.L.psb_slab_route_scatter_peer(pySpec._ptr(pos), int(pos.dtype == torch.float64), aos, pySpec._ptr(wt),
                                                        int(wt.dtype == torch.float64), npart, N, float(L), np.float32(N / L), np.float32(0.),
                                                        N // world, world, pySpec._ptr(dest), pySpec._ptr(cursor), pySpec._stream()),
                     'psb_slab_route_scatter_peer')
        assert np.array_equal(cursor.cpu().numpy(), Mh[r])
    rows = lambda t: np.sort(np.ascontiguousarray(t.cpu().numpy()).view([('', np.float32)] * 4).ravel())
    for r, (pos, aos, wt) in enumerate(dev):
        send = M.route_scatter(pipe, pos, aos, wt, L, world, counts[r])
        base = np.concatenate([[0], np.cumsum(Mh[r])])
        seg = Mh[:r].sum(axis=0)
        for d in range(world):
            got = bufs[d][int(seg[d]):int(seg[d]) + int(Mh[r, d])]
            assert np.array_equal(rows(got), rows(send[base[d]:base[d + 1]])), (r, d)
    for d in range(world):                                                    # nothing written past the end
        assert bool(torch.isnan(bufs[d][int(recv_tot[d]):]).all())


def test_c5_grid_through_every_step_of_the_sharded_path(mods, monkeypatch, tmp_path):
    """BASELINE configs[4]'s grid and shells (Ngrid = 1024, step 3, Ncut 3, Nmax 40: 6350 triangles, 400^3 carrier) at a particle
    count one GPU holds, through every step of the sharded path on one rank (route, slab assignment incl. the two-pass sort candidate,
    slab FFT, slab binning, carrier, dealt pairs, cell slabs) against the single-GPU API.  (The multi-rank run of the same
    comparison is tools/run_sharded.py c5check: profiles/r2_summary.md.)"""
    pySpec, M = mods
    monkeypatch.setattr(pySpec, '_DAT_DIR', str(tmp_path))
    N, L, Np, step, Ncut, Nmax = 1024, 4000., 3000000, 3, 3, 40
    xyz, w = _catalogue(N, Np, L, 5)
    assert M.carrier_grid(N, step, Nmax, Ncut) == 400
    ref = pySpec.Bk_periodic(xyz, w=w, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
    got, pk = M.Bk_periodic_sharded(xyz, w, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax, return_pk=True)
    assert len(ref['b123']) == 6350 and np.array_equal(got['i_k1'], ref['i_k1']) and np.array_equal(got['i_k3'], ref['i_k3'])
    assert np.array_equal(got['counts'], ref['counts'])
    np.testing.assert_allclose(got['p0k1'] + got['p0k_sn'], ref['p0k1'] + ref['p0k_sn'], rtol=1e-5)
    scale = np.abs(ref['b123'] + ref['b123_sn'])
    assert np.all(np.abs(got['b123'] - ref['b123']) <= 1e-5 * scale + 1e-7 * scale.max())
    rpk = pySpec.Pk_periodic(xyz, w=w, Lbox=L, Ngrid=N)
    assert np.array_equal(pk['counts'], rpk['counts'])
    np.testing.assert_allclose(pk['p0k'] + pk['p0k_sn'], rpk['p0k'] + rpk['p0k_sn'], rtol=1e-5)
