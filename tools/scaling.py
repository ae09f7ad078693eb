#!/usr/bin/env python
"""Strong / weak scaling of the step over y-slabs (GPU box, under torch.distributed.run; N = 1 works too):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
       tools/scaling.py --width 16384 --height 16384 [--iters 50] [--halos 32,118] [--steps 10] [--weak]
   --weak: --height is rows PER GPU (total = height x N); otherwise --height is the whole domain (strong scaling).
   Prints one JSON line per halo on rank 0: ms/step (CUDA events on the sim's stream, max over ranks), cell-steps/s."""
import json, os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from opensayal_b200.slab import SlabFluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
fd = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
os.dup2(fd, 1)
W, H, n = int(arg("--width", "16384")), int(arg("--height", "16384")), int(arg("--iters", "50"))
steps, weak = int(arg("--steps", "10")), "--weak" in sys.argv
if weak:
    H *= world
halos = [int(x) for x in arg("--halos", "32,118").split(",")] if world > 1 else [0]
for halo in halos:
    cfg = baseline_config(1, width=W, height=H)
    cfg["sim.projection.n"] = n
    cfg["sim.wind_tunnel.pipe_height"] = H // 4
    cfg["sim.obstacle.center_x"], cfg["sim.obstacle.center_y"] = W // 2, H // 2
    cfg["sim.obstacle.radius"] = float(max(4, H // 30))
    sf = SlabFluid(cfg, rank, world, local, halo=halo)
    u, v, sm = synthetic_fields(W, H, rows=(sf.row0, sf.rows))
    sf.set_initial(u, v, sm)
    del u, v, sm
    st = torch.cuda.ExternalStream(sf.sim.stream)
    sf.run(3); sf.sync(); dist.barrier()
    evs = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); sf.run(1); b.record(st); evs.append((a, b))
    sf.sync()
    ms_local = sum(x.elapsed_time(y) for x, y in evs) / steps
    t = torch.tensor([ms_local], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ovf = torch.tensor([sf.halo_overflow() if world > 1 else 0], device="cuda", dtype=torch.int64)
    dist.all_reduce(ovf)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "grid": [W, H], "sor_iterations": n, "scaling": "weak" if weak else "strong",
                          "halo_rows": halo, "steps": steps, "ms_per_step": round(float(t[0]), 4),
                          "cell_steps_per_s": W * H / (float(t[0]) * 1e-3), "halo_overflow": int(ovf[0]),
                          "plan": [sf.sim.get_option("plan_temporal_block"), sf.sim.get_option("plan_rows_per_warp")]}), flush=True)
    sf.close(); dist.barrier()
dist.destroy_process_group()
