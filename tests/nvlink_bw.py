"""Diagnostic (torchrun, >= 2 GPUs): kernel-driven NVLink peer-store time for the
sizes the multi-GPU BFS ships per level.  Not a test; prints a table."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    gdist = importlib.import_module("libgrape-lite_b200.dist")
    comm = gdist.make_comm(rank, world, 1 << 24)     # mirror slots of 128 MB
    for mb in (0.25, 1, 2, 4, 16, 64):
        n = int(mb * (1 << 20))
        for vec16 in (1, 0):
            for allp in ((0, 1) if world > 2 else (0,)):
                dist.barrier()
                us = comm.peer_write_us(n, bool(vec16), bool(allp), reps=20)
                if rank == 0:
                    tot = n * ((world - 1) if allp else 1)
                    print("[nvlink] %6.2f MB x %d peer(s) %s-byte stores: %8.1f us  %7.1f GB/s"
                          % (mb, (world - 1) if allp else 1, "16" if vec16 else " 4", us, tot / us * 1e-3), flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
