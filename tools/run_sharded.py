#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/run_sharded.py [small|c2|c4|c5|c5check]: one catalogue sharded over G GPUs
(pyspectrum_b200.multigpu: slab-owned assignment, slab FFT, slab binning, carrier grid, sharded shell/triangle stage)
checked against the single-GPU API on rank 0 (when the problem fits one GPU)."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyspectrum_b200 import dist as D, multigpu as M, pyspectrum as pySpec

rank, world, local = D.rank_info()
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
D.init('nccl', dev)
which = sys.argv[1] if len(sys.argv) > 1 else 'small'
cfg = {'small': dict(N=64, L=500., Np=200000, step=2, Ncut=3, Nmax=12, check=True),
       'c2': dict(N=360, L=2600., Np=10 ** 7, step=3, Ncut=3, Nmax=40, check=True),
       'c4': dict(N=512, L=2600., Np=10 ** 7, step=2, Ncut=3, Nmax=80, check=True),
       'c5': dict(N=1024, L=4000., Np=10 ** 9, step=3, Ncut=3, Nmax=40, check=False),
       'c5check': dict(N=1024, L=4000., Np=10 ** 8, step=3, Ncut=3, Nmax=40, check=True)}[which]    # C5's grid at a size one GPU holds
N, L = cfg['N'], cfg['L']
if which == 'c5':
    shard, full = bench.c5_shard(dev, rank, world), None
else:
    full = bench.lognormal_catalogue_torch(2, dev, cfg['Np'], L, min(N, 360))      # same seed on every rank
    shard = full[:, rank::world].contiguous()
res = bench.run_sharded_config(which, shard, cfg, dev, world, reps=3 if which != 'c5' else 2)
out, pk = M.Bk_periodic_sharded(shard, None, Lbox=L, Ngrid=N, step=cfg['step'], Ncut=cfg['Ncut'], Nmax=cfg['Nmax'], return_pk=True)
pr = M.Pk_periodic_rsd_sharded(shard, None, Lbox=L, Ngrid=N, rsd=2, Nmubin=10)
if cfg['check'] and rank == 0:
    ref = pySpec.Bk_periodic(full, Lbox=L, Ngrid=N, step=cfg['step'], Ncut=cfg['Ncut'], Nmax=cfg['Nmax'])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ref = pySpec.Bk_periodic(full, Lbox=L, Ngrid=N, step=cfg['step'], Ncut=cfg['Ncut'], Nmax=cfg['Nmax'])
    torch.cuda.synchronize(); res['single_gpu_s'] = time.perf_counter() - t0
    raw = lambda d: d['b123'] + d['b123_sn']
    res['counts_equal'] = bool(np.array_equal(out['counts'], ref['counts']))
    res['ik_equal'] = bool(np.array_equal(out['i_k1'], ref['i_k1']) and np.array_equal(out['i_k3'], ref['i_k3']))
    res['p0k1_max_rel'] = float(np.abs((out['p0k1'] + out['p0k_sn']) / (ref['p0k1'] + ref['p0k_sn']) - 1).max())
    scale = np.abs(raw(ref))
    res['b123_within_1e-5'] = bool(np.all(np.abs(raw(out) - raw(ref)) <= 1e-5 * scale + 1e-7 * scale.max()))
    res['b123_raw_median_rel'] = float(np.median(np.abs(raw(out) / raw(ref) - 1)))
    res['b123_raw_max_rel'] = float(np.abs(raw(out) / raw(ref) - 1).max())
    refpk = pySpec.Pk_periodic(full, Lbox=L, Ngrid=N)
    res['pk_counts_equal'] = bool(np.array_equal(pk['counts'], refpk['counts']))
    res['pk_raw_max_rel'] = float(np.abs((pk['p0k'] + pk['p0k_sn']) / (refpk['p0k'] + refpk['p0k_sn']) - 1).max())
    refpr = pySpec.Pk_periodic_rsd(full, Lbox=L, Ngrid=N, rsd=2, Nmubin=10)
    res['rsd_counts_equal'] = bool(np.array_equal(pr['counts'], refpr['counts']) and np.array_equal(pr['counts_kmu'], refpr['counts_kmu']))
    res['rsd_p2k_max_abs_over_p0'] = float((np.abs(pr['p2k'] - refpr['p2k']) / np.abs(refpr['p0k'] + refpr['p_sn'])).max())
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
