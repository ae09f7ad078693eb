#!/usr/bin/env python
"""Multi-GPU probe (run under torch.distributed.run on the GPU box): step / exchange / projection timings per halo.
   python -m torch.distributed.run --nproc-per-node N tools/slab_probe.py [--halos 16,20,32,50] [--rows-per-gpu 1080] [--width 1920]"""
import os, sys, contextlib
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from opensayal_b200.slab import SlabFluid, F_U, F_V, F_SMOKE
from opensayal_b200.synthetic import baseline_config, synthetic_fields

def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fd = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
os.dup2(fd, 1)
W = int(arg("--width", "1920")); rows = int(arg("--rows-per-gpu", "1080")); n = int(arg("--n", "50"))
H = rows * world
for halo in [int(x) for x in arg("--halos", "32,50,118").split(",")]:
    cfg = baseline_config(1, width=W, height=H)
    cfg["sim.projection.n"] = n
    cfg["sim.wind_tunnel.pipe_height"] = H // 4
    sf = SlabFluid(cfg, rank, world, local, halo=halo)
    u, v, sm = synthetic_fields(W, H, rows=(sf.row0, sf.rows))
    sf.set_initial(u, v, sm)
    st = torch.cuda.ExternalStream(sf.sim.stream)
    sf.run(5); sf.sync(); dist.barrier()
    def timed(fn, reps):
        evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); fn(); b.record(st); evs.append((a, b))
        sf.sync()
        t = sorted(x.elapsed_time(y) for x, y in evs)
        return t[len(t) // 2] * 1e3
    step = timed(lambda: sf.run(1), 30)
    dist.barrier()
    ex2 = timed(lambda: sf.sim.slab_exchange(F_U | F_V), 30)
    dist.barrier()
    ex1 = timed(lambda: sf.sim.slab_exchange(F_SMOKE), 30)
    dist.barrier()
    per = halo // 2
    proj = timed(lambda: sf.sim.stage_projection(min(per, n), cfg.c.d_t), 10)
    t = torch.tensor([step, ex2, ex1, proj], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ovf = sf.halo_overflow(); err = sf.sim.get_option("link_error")
    if rank == 0:
        chunks = -(-n // per)
        print(f"N={world} {W}x{rows}/gpu halo {halo:3d} ({chunks} chunks of <= {per} it): step {t[0]:7.1f} us  exchange(u,v) {t[1]:6.1f} us  exchange(smoke) {t[2]:6.1f} us  "
              f"projection({min(per,n)} it, T={sf.sim.get_option('plan_temporal_block')} rows={sf.sim.get_option('plan_rows_per_warp')}) {t[3]:6.1f} us  overflow {ovf} link_error {err}", flush=True)
    sf.close(); dist.barrier()
dist.destroy_process_group()
