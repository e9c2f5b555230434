"""torch.distributed plumbing for fragment groups (one process per GPU).

torch is used ONLY for rendezvous and tiny host collectives (exchange of the
64-byte CUDA-IPC handles, the per-superstep vote); the halo data path is the
library's own peer-memory kernels (csrc/comm.cu)."""
import numpy as np

from . import capi


def make_comm(rank, world, ivnum, item_bytes=16, group=None):
    """Creates, exports, exchanges and opens a Comm for this rank."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)

    def allreduce(arr, op):
        ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}
        if backend == "nccl":
            t = torch.from_numpy(arr).cuda()
            dist.all_reduce(t, op=ops[op], group=group)
            arr[:] = t.cpu().numpy()
        else:
            t = torch.from_numpy(arr)
            dist.all_reduce(t, op=ops[op], group=group)

    # a slot must hold one item per inner vertex of the RECEIVER (each outer
    # copy sends at most one item per round); sized by the largest fragment
    iv = torch.tensor([ivnum], dtype=torch.int64)
    if backend == "nccl":
        iv = iv.cuda()
    dist.all_reduce(iv, op=dist.ReduceOp.MAX, group=group)
    landing = item_bytes * (int(iv.item()) + 1024)
    # dense mirror sync: up to 8 bytes (f64 / int64 state) per inner vertex of the owner
    comm = capi.Comm(rank, world, allreduce, landing_bytes=landing, mirror_bytes=8 * (int(iv.item()) + 1024))
    handles = [None] * world
    dist.all_gather_object(handles, comm.export(), group=group)
    comm.open(handles)
    return comm
