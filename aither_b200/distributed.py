"""One process per GPU: the host-side plumbing around the C ABI's multi-GPU entry points.

The reference is an MPI program with one rank per group of blocks (reference src/main.cpp:60-170,
src/parallel.cpp:44-178); here one rank drives one B200 and `torch.distributed` plays the part MPI
plays in the reference's main.cpp: it carries the 128-byte NCCL id from rank 0 to the other ranks
(MPI_Bcast there) and sums the residual norms every rank returns (MPI_Reduce at main.cpp:249-264).
The ghost-layer exchange itself never touches torch: it is ncclSend/ncclRecv inside the library.
"""
import ctypes as C
import os

import numpy as np

from . import load_library, AitherGpuError


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend=None):
    """torch.distributed rendezvous on 127.0.0.1 unless the launcher already said otherwise."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    rank, world, local = env_rank()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    dist.init_process_group(backend=backend, rank=rank, world_size=world)


def broadcast_bytes(payload, n, src=0):
    """Rank `src` passes `n` bytes to everyone (gloo: CPU tensor; nccl: a tensor on this GPU)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def all_gather_bytes(payload, n):
    """Every rank contributes `n` bytes; returns the world_size * n bytes in rank order."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    parts = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def make_comm(device):
    """NCCL communicator for the halo exchange, created through the library
    (aither_gpu_comm_unique_id on rank 0 -> broadcast -> aither_gpu_comm_create)."""
    import torch.distributed as dist
    L = load_library()
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = C.create_string_buffer(128)
    if rank == 0 and L.aither_gpu_comm_unique_id(ident) != 0:
        raise AitherGpuError(L.aither_gpu_last_error().decode())
    raw = broadcast_bytes(ident.raw, 128, src=0)
    comm = C.c_void_p()
    if L.aither_gpu_comm_create(raw, rank, world, device, C.byref(comm)) != 0:
        raise AitherGpuError(L.aither_gpu_last_error().decode())
    return comm


def destroy_comm(comm):
    L = load_library()
    if comm and L.aither_gpu_comm_destroy(comm) != 0:
        raise AitherGpuError(L.aither_gpu_last_error().decode())


def reduce_norms(local_l2, local_matrix_sumsq_over_size=None, local_size=None):
    """Sum of the per-rank residual sums (what MPI_Reduce does at reference src/main.cpp:249-255).
    The matrix residual is sum(mr^2)/size per rank (src/mgSolution.cpp:199-206); with sizes given it
    is recombined as the global sum over the global size."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    vals = list(np.asarray(local_l2, dtype=np.float64).ravel())
    if local_matrix_sumsq_over_size is not None:
        vals += [local_matrix_sumsq_over_size * local_size, float(local_size)]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy()
    if local_matrix_sumsq_over_size is None:
        return out.reshape(np.shape(local_l2))
    n = len(vals) - 2
    return out[:n].reshape(np.shape(local_l2)), out[n] / out[n + 1]


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
