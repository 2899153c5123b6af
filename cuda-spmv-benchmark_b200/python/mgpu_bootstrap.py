"""One-process-per-GPU bootstrap for the multi-GPU CG path (torch.distributed is plumbing only).

Each rank allocates its exchange block inside libspmv_b200.so (b200_mgpu_init_rank), the 64-byte
CUDA IPC handles are all-gathered here, and every rank maps its peers (b200_mgpu_connect).  After
that the library moves halos and scalar sums itself through NVLink peer memory.

The same functions run under the `gloo` backend on CPU (tests/test_dist_gloo.py) with a fake
handle provider, which covers rank ordering, the row-band partition and the barrier discipline
without GPUs.
"""
import ctypes as C

HANDLE_BYTES = 64


def partition(n_rows, world, rank):
    """Row band of `rank`: n_local = N // P, offset = rank * n_local, last rank takes the remainder
    (reference src/solvers/cg_solver_mgpu_partitioned.cu:262-268)."""
    q = n_rows // world
    off = rank * q
    return (n_rows - off if rank == world - 1 else q), off


def all_gather_handles(dist, my_handle, device):
    """All-gather fixed-size byte strings; returns world * HANDLE_BYTES bytes in rank order."""
    import torch
    assert len(my_handle) == HANDLE_BYTES
    mine = torch.tensor(list(my_handle), dtype=torch.uint8, device=device)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return torch.cat(out).cpu().numpy().tobytes()


def connect(L, dist, rank, world, local_device, max_grid, device="cuda"):
    """Full bootstrap against the library `L` (ctypes handle of libspmv_b200.so or a test double)."""
    handle = (C.c_ubyte * HANDLE_BYTES)()
    rc = L.b200_mgpu_init_rank(rank, world, local_device, max_grid, handle)
    if rc:
        raise RuntimeError("b200_mgpu_init_rank rc=%d" % rc)
    raw = all_gather_handles(dist, bytes(handle), device)
    blob = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
    rc = L.b200_mgpu_connect(blob)
    if rc:
        raise RuntimeError("b200_mgpu_connect rc=%d" % rc)
    dist.barrier()
    return raw
