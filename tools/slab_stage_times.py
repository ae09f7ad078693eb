#!/usr/bin/env python
"""Stage boundaries of a linked slab's step, per rank (under torch.distributed.run): eager steps with option
debug_events; prints the median over steps of the time from step start to each boundary (sayal_debug_stage_times).
   python -m torch.distributed.run --nproc-per-node 4 ... tools/slab_stage_times.py [--halo 118] [--push 0]"""
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
W, rows = 1920, 1080
H = rows * world
cfg = baseline_config(1, width=W, height=H)
cfg["sim.wind_tunnel.pipe_height"] = H // 4
halo = int(arg("--halo", "0")) or None
sf = SlabFluid(cfg, rank, world, local, halo=halo, push=None if halo is None else bool(int(arg("--push", "0"))))
u, v, sm = synthetic_fields(W, H, rows=(sf.row0, sf.rows))
sf.set_initial(u, v, sm)
sf.run(5); sf.sync(); dist.barrier()
sf.sim.set_option("use_graph", 0)
sf.sim.set_option("debug_events", 1)
times = []
for k in range(12):
    sf.sim.stream_hold()
    sf.update()
    dist.barrier()
    sf.sim.stream_release()
    sf.sync()
    times.append(sf.sim.debug_stage_times())
    dist.barrier()
t = np.median(np.array(times[2:]), axis=0) * 1e3
names = ["start", "projection", "velocity", "v edges", "s edges", "exchange(aux)", "s interior", "end"]
line = f"rank {rank} (halo {sf.halo}, push {sf.sim.get_option('push_mode')}, T {sf.sim.get_option('plan_temporal_block')} rows {sf.sim.get_option('plan_rows_per_warp')}): " + \
       ", ".join(f"{n} {x:.1f}" for n, x in zip(names, t) if x >= 0)
for r in range(world):
    if r == rank:
        print(line, flush=True)
    dist.barrier()
sf.close(); dist.barrier()
dist.destroy_process_group()
