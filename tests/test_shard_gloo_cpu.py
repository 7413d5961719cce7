"""CPU test of the N>1 host logic: two gloo ranks each evaluate their share of the
k list (structure factor) with the oracle, all-reduce the partial energies the way
bench.py does over NCCL, and must reproduce the single-rank result; the contiguous
split covers every unit exactly once; timings reduce with MAX."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from gomc_b200 import shard


def test_split_covers_everything_once():
    for n in (0, 1, 7, 148, 729, 102978):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                a, b = shard.split_range(n, r, world)
                assert 0 <= a <= b <= n
                seen.extend(range(a, b))
            assert seen == list(range(n))
            sizes = [shard.split_range(n, r, world) for r in range(world)]
            assert max(b - a for a, b in sizes) - min(b - a for a, b in sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from gomc_b200 import synth
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    po.set_threads(1)
    s = synth.make_spce(125, r_cut=7.0)
    o = po.Oracle.from_system(s)
    bm = np.arange(s.n_mols, dtype=np.int32)
    kx, ky, kz, hs, pf, _ = o.recip_init_orth()
    k0, k1 = shard.split_range(len(kx), rank, world)
    sR, sI = o.box_recip_sums(bm, s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz, k0, k1)
    recip_part = o.box_reciprocal(sR[k0:k1], sI[k0:k1], pf[k0:k1])
    lj = re = 0.0
    if rank == 0:   # the pair sweep of this tiny box lives on one rank
        lj, re = o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge,
                             np.arange(s.n_atoms, dtype=np.int32))
    tot = shard.allreduce_energies([lj, re, recip_part], world)
    t = shard.max_over_ranks(10.0 + rank, world)
    if rank == 0:
        q.put((tot, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_matches_single_rank():
    from gomc_b200 import synth
    from oracle import pyoracle as po
    s = synth.make_spce(125, r_cut=7.0)
    o = po.Oracle.from_system(s)
    po.set_threads(1)
    bm = np.arange(s.n_mols, dtype=np.int32)
    kx, ky, kz, hs, pf, _ = o.recip_init_orth()
    sR, sI = o.box_recip_sums(bm, s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
    full = o.box_reciprocal(sR, sI, pf)
    lj, re = o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge,
                         np.arange(s.n_atoms, dtype=np.int32))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tot[0] == lj and tot[1] == re
    assert abs(tot[2] - full) <= 1e-12 * abs(full)
    assert t == 11.0
