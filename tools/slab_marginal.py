#!/usr/bin/env python
"""Marginal cost of the stages of a LINKED slab step (under torch.distributed.run, N >= 2): the replayed step is
timed with one stage left out at a time (option debug_skip; results are wrong, only the time is used).
   python -m torch.distributed.run --nproc-per-node 2 ... tools/slab_marginal.py [--halos 50,118] [--width 1920] [--rows-per-gpu 1080]"""
import os, sys
sys.path.insert(0, ".")
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
W, rows = int(arg("--width", "1920")), int(arg("--rows-per-gpu", "1080"))
H = rows * world
for halo in [int(x) for x in arg("--halos", "50,118").split(",")]:
    cfg = baseline_config(1, width=W, height=H)
    cfg["sim.wind_tunnel.pipe_height"] = H // 4
    sf = SlabFluid(cfg, rank, world, local, halo=halo)
    u, v, sm = synthetic_fields(W, H, rows=(sf.row0, sf.rows))
    st = torch.cuda.ExternalStream(sf.sim.stream)
    res = {}
    for label, mask in (("full", 0), ("-advect_v", 4), ("-advect_s", 8), ("-end exchange", 32), ("-projection", 16),
                        ("only projection", 4 | 8 | 32), ("full again", 0)):
        sf.sim.set_option("debug_skip", mask)
        sf.set_initial(u, v, sm)
        sf.run(5); sf.sync(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); sf.run(100); e1.record(st); sf.sync()
        t = torch.tensor([e0.elapsed_time(e1) / 100 * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[label] = float(t[0])
        dist.barrier()
    if rank == 0:
        print(f"N={world} {W}x{rows}/gpu halo {halo}: " + ", ".join(f"{k} {v:.1f}" for k, v in res.items()), flush=True)
    sf.close(); dist.barrier()
dist.destroy_process_group()
