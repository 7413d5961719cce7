"""Multi-rank parity probe (torchrun --nproc-per-node N): every rank builds the same box twice,
once as a plain single-GPU engine and once as rank r of an N-rank sharded engine with the
engine's own NCCL communicator (gomcb200_set_comm), and compares everything the sharded engine
returns -- full-box energies, BoxForce atom / molecule forces, the reciprocal force, torques,
the MultiParticle acceptance weight of one trial -- with the single-GPU values.  Rank 0 prints
one JSON object of maximum relative differences (tests/test_shard_nccl_gpu.py asserts on it)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from gomc_b200 import engine as eng, synth

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "spce4096"
s = synth.make_spce(int(name[4:]), r_cut=9.0) if name.startswith("spce") else synth.make_argon(4000)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def evaluate(e):
    out = {}
    out["full"] = e.call_full_box_energy(0, s.x, s.y, s.z)
    out["inter"] = e.box_inter(0)
    e.set_com(*s.com())
    if s.ff.ewald:
        out["recip"] = e.box_reciprocal_sums(0)
    out["force_en"] = e.box_force(0)
    out["atomF"] = np.array(e.get_forces(eng.ATOM_FORCE))
    out["molF"] = np.array(e.get_forces(eng.MOL_FORCE))
    if s.ff.ewald:
        e.box_force_reciprocal(0)
        out["atomFrec"] = np.array(e.get_forces(eng.ATOM_FORCE_REC))
    e.calculate_torque(0)
    out["torque"] = np.array(e.get_forces(eng.MOL_TORQUE))
    # one MultiParticle trial (displacement): reference forces are the ones above
    e.copy_recip(0)
    e.mp_transform(0, 0, 0.02, 0.5 / 300.0, 1234, 0, 123)
    e.mp_select(1)
    if s.ff.ewald:
        e.box_reciprocal_sums(0)
    out["trial_en"] = e.box_force(0)
    if s.ff.ewald:
        e.box_force_reciprocal(0)
    e.calculate_torque(0)
    out["w"] = e.mp_coeff(0, 0, 0.02, 0.5 / 300.0)
    e.mp_select(0)
    return out


a = eng.Engine.from_system(s, device=local)
ref = evaluate(a)
a.close()
b = eng.Engine.from_system(s, device=local)
b.set_comm(bench.share_unique_id(eng, rank, world, dist, torch), rank, world)
got = evaluate(b)
b.close()
diff = {}
for k, v in ref.items():
    if isinstance(v, np.ndarray):
        diff[k] = rel(got[k], v)
    else:
        vv, gg = np.atleast_1d(np.array(v, float)), np.atleast_1d(np.array(got[k], float))
        diff[k] = float(np.max(np.abs(vv - gg) / np.maximum(np.abs(vv), 1.0)))
worst = torch.tensor([max(diff.values())], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"system": name, "world": world, "max_rel_diff": diff,
                      "worst_over_ranks": float(worst.item()), "w": got["w"]}))
if world > 1:
    dist.destroy_process_group()
