"""CPU: host logic of the single-catalogue multi-GPU path (pyspectrum_b200/multigpu.py) on world_size-2 gloo:
pair dealing, the all-to-all to slabs and the row bookkeeping that K6 relies on."""
import os

import torch


def test_pair_assignment_and_rows():
    from pyspectrum_b200.multigpu import pair_assignment, slab_field_rows
    for S, world in [(40, 1), (40, 2), (40, 8), (7, 2), (80, 4), (5, 4)]:
        npairs = (S + 1) // 2
        assign, per = pair_assignment(npairs, world)
        assert all(len(a) == per for a in assign)
        got = sorted(p for a in assign for p in a if p < npairs)
        assert got == list(range(npairs))
        rows = slab_field_rows(S, world, per)
        assert len(set(rows)) == S and max(rows) < 2 * per * world
        for f, r in enumerate(rows):                       # row -> (owner, k, e) -> shell slot
            owner, rem = divmod(r, 2 * per)
            k, e = divmod(rem, 2)
            assert 2 * (owner + world * k) + e == f


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo')
    from pyspectrum_b200.multigpu import pair_assignment, slab_field_rows, exchange_to_slabs
    S, ncell = 7, 24
    npairs = (S + 1) // 2
    assign, per = pair_assignment(npairs, world)
    # field of shell slot f at cell x has the value 100*f + x  (padding pairs: -1)
    rows = []
    for p in assign[rank]:
        for e in (0, 1):
            f = 2 * p + e
            rows.append(100. * f + torch.arange(ncell, dtype=torch.float32) if p < npairs else torch.full((ncell,), -1.))
    local = torch.stack(rows)
    slabs = exchange_to_slabs(local, world)
    slab = ncell // world
    fr = slab_field_rows(S, world, per)
    ok = True
    for f in range(S):
        expect = 100. * f + torch.arange(rank * slab, (rank + 1) * slab, dtype=torch.float32)
        ok &= bool(torch.equal(slabs[fr[f]], expect))
    out.put((rank, ok, tuple(slabs.shape)))
    dist.destroy_process_group()


def test_exchange_to_slabs_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, ok, shape in res:
        assert ok and shape == (8, 12)


def test_exchange_single_rank_is_a_reshape():
    from pyspectrum_b200.multigpu import exchange_to_slabs, slab_field_rows
    local = torch.arange(4 * 10, dtype=torch.float32).view(4, 10)
    slabs = exchange_to_slabs(local, 1)
    rows = slab_field_rows(4, 1, 2)
    for f in range(4):
        assert torch.equal(slabs[rows[f]], local[f])


def _slab_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo')
    from pyspectrum_b200.multigpu import slab_geometry, z_to_y_slabs, gather_ky_slabs
    N = 8
    nz, hp = slab_geometry(N, world)
    z, y, x, c = torch.meshgrid(torch.arange(N), torch.arange(N), torch.arange(hp), torch.arange(2), indexing='ij')
    T = (z * 10000 + y * 100 + x * 2 + c).float()                 # global array [z][ky][kx][re/im]
    mine = T[rank * nz:(rank + 1) * nz].contiguous()              # this rank's z-slab
    ys = z_to_y_slabs(mine, world)                                # -> all z of this rank's ky range
    ok = bool(torch.equal(ys, T[:, rank * nz:(rank + 1) * nz]))
    full = gather_ky_slabs(ys, world)                             # -> replicated
    ok &= bool(torch.equal(full, T))
    out.put((rank, ok, tuple(ys.shape)))
    dist.destroy_process_group()


def test_slab_exchange_two_ranks_gloo():
    """z-slabs -> ky-slabs all-to-all and the ky-slab all-gather of the slab-decomposed FFT (multigpu.slab_mesh_to_delta)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31700 + os.getpid() % 2000
    ps = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, ok, shape in res:
        assert ok and shape == (8, 4, 6, 2)


def test_slab_geometry_and_single_rank_exchange():
    from pyspectrum_b200.multigpu import slab_geometry, z_to_y_slabs, z_to_y_chunks, gather_ky_slabs
    assert slab_geometry(360, 2) == (180, 182) and slab_geometry(1024, 8) == (128, 514) and slab_geometry(24, 4) == (6, 14)
    try:
        slab_geometry(360, 7)
        assert False
    except ValueError:
        pass
    t = torch.arange(4 * 8 * 6 * 2, dtype=torch.float32).view(4, 8, 6, 2)
    ch = z_to_y_chunks(t, 2)                                      # chunk q = ky range q of these planes
    assert torch.equal(ch[0], t[:, :4]) and torch.equal(ch[1], t[:, 4:])
    full = torch.arange(8 * 8 * 6 * 2, dtype=torch.float32).view(8, 8, 6, 2)
    assert torch.equal(z_to_y_slabs(full, 1), full) and torch.equal(gather_ky_slabs(full, 1), full)


def test_cell_slab_checks_raise_before_any_collective():
    import pytest
    from pyspectrum_b200.multigpu import check_cell_slabs
    check_cell_slabs(360 ** 3, 8)
    with pytest.raises(ValueError):
        check_cell_slabs(360 ** 3, 7)                       # 360^3 is not a multiple of 7
    with pytest.raises(ValueError):
        check_cell_slabs(10 ** 3, 8)                        # 125 cells per slab would split a packed cell pair
    check_cell_slabs(10 ** 3, 8, packed=False)


def test_carrier_grid_choice():
    from pyspectrum_b200.multigpu import carrier_grid
    assert carrier_grid(1024, 3, 40, 3) == 400              # 3*(3*40+1) = 363 < 400: BASELINE configs[4] never needs the 1024^3 shell fields
    assert carrier_grid(360, 3, 40, 3) == 360               # the reference's own grid lets the largest triangles wrap: stay on it
    assert carrier_grid(512, 2, 80, 3) == 512               # 483 < 512 but no compiled grid in between
    assert carrier_grid(512, 3, 20, 3) == 256


def _worker_segments(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import numpy as np
    import torch.distributed as dist
    dist.init_process_group('gloo')
    from pyspectrum_b200.multigpu import peer_segments
    counts = torch.tensor([[5, 2, 0], [1, 7, 3], [4, 0, 6]][rank][:world], dtype=torch.int64)        # copies this rank sends to every rank
    gathered = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(gathered, counts)
    m = np.stack([g.numpy() for g in gathered])
    recv_tot, seg = peer_segments(m, rank)
    # every rank "stores" its segment ids into every destination buffer through an all-to-all of (offset, length) pairs
    out.put((rank, recv_tot.tolist(), seg.tolist(), counts.tolist()))
    dist.destroy_process_group()


def test_peer_segments_tile_the_receive_buffers_gloo():
    """The segments of all sources tile every destination's receive buffer exactly (no gap, no overlap), in rank order."""
    import torch.multiprocessing as mp
    world = 3
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_segments, args=(r, world, 29631, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for d in range(world):
        spans = sorted((res[s][2][d], res[s][2][d] + res[s][3][d]) for s in range(world))
        assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == res[0][1][d]
        assert [sp[0] for sp in spans] == [res[s][2][d] for s in range(world)]           # rank order
    assert all(r[1] == res[0][1] for r in res)


def test_route_tables_address_the_rows_the_all_to_all_would_fill():
    """SlabBuffers.route / row_table (host bookkeeping of the peer-store exchanges) with made-up base addresses: entry (k, e, q) of a
    rank's route table + the byte offset of a cell in slab q must be the address of that cell in row slab_field_rows()[shell] of rank
    q's buffer; shells that do not exist get 0."""
    import numpy as np
    from pyspectrum_b200.multigpu import SlabBuffers, pair_assignment, slab_field_rows
    world, S, slab = 4, 13, 1000                                        # 13 shells -> 7 pairs -> 2 per rank, one padding pair
    deal, per = pair_assignment((S + 1) // 2, world)
    ptrs = [10 ** 9 * (q + 1) for q in range(world)]
    rows = slab_field_rows(S, world, per)
    seen = set()
    for rank in range(world):
        b = SlabBuffers(world, rank, world * 2 * per, slab, None, ptrs)
        t = b.route(per, deal[rank], S, 'cpu').numpy()
        assert t.shape == (per, 2, world)
        for k, pidx in enumerate(deal[rank]):
            for e in (0, 1):
                f = 2 * pidx + e
                for q in range(world):
                    if f >= S:
                        assert t[k, e, q] == 0
                        continue
                    cell = q * slab + 17                                  # a cell of slab q, addressed as if the buffer held the whole field
                    assert t[k, e, q] + 4 * cell == ptrs[q] + 4 * (rows[f] * slab + 17)
                    seen.add((f, q))
        rt = b.row_table('cpu').numpy()
        assert rt.shape == (world * 2 * per, world) and rt[3, 2] == ptrs[2] + 4 * 3 * slab
    assert len(seen) == S * world
