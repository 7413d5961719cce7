"""Multi-rank timing probe (torchrun): E1 on one workload with the engine's communicator;
GOMCB200_NUFFT_TRACE=1 prints the per-phase times of the sharded non-uniform FFT."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from gomc_b200 import engine as eng
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = bench.make_system(sys.argv[1] if len(sys.argv) > 1 else "spce100k")
e = eng.Engine.from_system(s, device=local)
e.set_comm(bench.share_unique_id(eng, rank, world, dist, torch), rank, world)
e.enable_timing(True)
for i in range(6):
    e.mark_coords_changed()
    t0 = time.perf_counter(); en = e.call_full_box_energy(0); torch.cuda.synchronize()
    w = (time.perf_counter() - t0) * 1e3
    e.box_inter(0); p = e.last_timing()[0]
    if i >= 3: print(f"rank {rank} step {i} wall {w:.3f} pair-alone {p:.3f}", flush=True)
e.close()
if world > 1: dist.destroy_process_group()
