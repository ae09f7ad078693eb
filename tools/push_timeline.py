#!/usr/bin/env python
"""Per-tile phase times of the LAST projection pass of a linked slab step in push mode (under torch.distributed.run):
entry -> loaded (includes the wait for the neighbour's pass flag in the warps that load ghost rows) -> swept -> stored
(includes the peer stores, the system fence and the flag).  Prints edge tile rows and the rest separately.
   python -m torch.distributed.run --nproc-per-node 2 ... tools/push_timeline.py [--halo 18]"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import torch, torch.distributed as dist
from opensayal_b200.slab import SlabFluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fd = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier()
os.dup2(fd, 1)
W, rows, halo = 1920, 1080, int(arg("--halo", "18"))
H = rows * world
cfg = baseline_config(1, width=W, height=H)
cfg["sim.wind_tunnel.pipe_height"] = H // 4
sf = SlabFluid(cfg, rank, world, local, halo=halo)
u, v, sm = synthetic_fields(W, H, rows=(sf.row0, sf.rows))
sf.set_initial(u, v, sm)
sf.run(5); sf.sync(); dist.barrier()
sf.sim.set_option("debug_timeline", 1)
sf.sim.set_option("use_graph", 0)
sf.run(3); sf.sync(); dist.barrier()
t = sf.sim.debug_timeline().astype(np.float64)
T, ry = sf.sim.get_option("plan_temporal_block"), sf.sim.get_option("plan_rows_per_warp")
out = [f"rank {rank}: plan T={T} rows/warp={ry}, {len(t)} tiles in the last pass; pass span {(t[:, 3].max() - t[:, 0].min()) / 1e3:.1f} us"]
if len(t):
    t0 = t[:, 0].min()
    order = np.argsort(t[:, 0])
    load, sweep, store = (t[:, 1] - t[:, 0]) / 1e3, (t[:, 2] - t[:, 1]) / 1e3, (t[:, 3] - t[:, 2]) / 1e3
    k = np.argsort(-store)[:24]  # the pushers are the tiles with the longest store phase
    rest = np.setdiff1d(np.arange(len(t)), k)
    for name, idx in (("24 longest stores", k), ("rest", rest)):
        out.append(f"  {name:18s} entry +{np.median(t[idx, 0] - t0) / 1e3:6.2f}  load med {np.median(load[idx]):6.2f} max {load[idx].max():6.2f}  "
                   f"sweep med {np.median(sweep[idx]):6.2f} max {sweep[idx].max():6.2f}  store med {np.median(store[idx]):6.2f} max {store[idx].max():6.2f}  "
                   f"end +{np.median(t[idx, 3] - t0) / 1e3:6.2f} max +{(t[idx, 3] - t0).max() / 1e3:6.2f}")
    k2 = np.argsort(-load)[:24]
    out.append(f"  24 longest loads: load med {np.median(load[k2]):6.2f} max {load[k2].max():6.2f}, entry +{np.median(t[k2, 0] - t0) / 1e3:6.2f}")
for r in range(world):
    if r == rank:
        print("\n".join(out), flush=True)
    dist.barrier()
sf.close(); dist.barrier()
dist.destroy_process_group()
