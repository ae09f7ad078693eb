"""Worker of tests/test_slab_gpu.py::test_two_processes_ipc_link (launched by torch.distributed.run)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from opensayal_b200 import Fluid
from opensayal_b200.slab import SlabFluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = int(os.environ.get("SAYAL_IPC_TEST_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    dist.init_process_group("gloo")
    cfg = baseline_config(1, width=384, height=420)
    cfg["sim.projection.n"] = 20
    cfg["sim.wind_tunnel.speed"] = 60.0
    c = cfg.c
    sf = SlabFluid(cfg, rank, world, device, halo=16, transport="p2p", margin=6)
    u, v, sm = synthetic_fields(c.width, c.height, rows=(sf.row0, sf.rows))
    sf.set_initial(u, v, sm)
    sf.run(2)          # graph replays
    sf.update()        # and one eager step
    sf.sync()
    mine = {n: sf.sim.get_field(n) for n in ("u", "v", "smoke")}
    status = (sf.halo_overflow(), sf.sim.get_option("link_error"))
    parts = [None] * world
    dist.gather_object((mine, status), parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        f = Fluid(cfg, device=device)
        gu, gv, gs = synthetic_fields(c.width, c.height)
        for n, a in (("u", gu), ("v", gv), ("smoke", gs)):
            f.set_field(n, a)
        for _ in range(3):
            f.update(None)
        for n in ("u", "v", "smoke"):
            got = np.concatenate([p[0][n] for p in parts])
            if not np.array_equal(got, f.get_field(n)):
                ok = False
                print(f"MISMATCH {n}: {(got != f.get_field(n)).sum()} cells", flush=True)
        if any(p[1] != (0, 0) for p in parts):
            ok = False
            print("overflow / link_error:", [p[1] for p in parts], flush=True)
        f.close()
        print("IPC-SLABS-OK" if ok else "IPC-SLABS-FAILED", flush=True)
    sf.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
