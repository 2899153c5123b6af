"""CPU, world_size 2 (gloo): the one-process-per-GPU bootstrap used by bench.py under torchrun --
handle all-gather in rank order, row-band partition, halo neighbour wiring -- with a test double
in place of the CUDA library calls."""
import ctypes as C
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class FakeLib:
    """records what the bootstrap hands to the library"""

    def __init__(self):
        self.connected = None

    def b200_mgpu_init_rank(self, rank, world, dev, max_grid, handle):
        for i in range(64):
            handle[i] = (rank * 64 + i) % 251
        self.args = (rank, world, dev, max_grid)
        return 0

    def b200_mgpu_connect(self, blob):
        self.connected = bytes(blob)
        return 0


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    import mgpu_bootstrap as mb
    import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = FakeLib()
        raw = mb.connect(lib, dist, rank, world, 0, n, device="cpu")
        expect = b"".join(bytes((r * 64 + i) % 251 for i in range(64)) for r in range(world))
        ok = raw == expect == lib.connected and lib.args == (rank, world, 0, n)
        nl, off = mb.partition(n * n, world, rank)
        ok = ok and (nl, off) == orc.partition(n * n, world, rank)
        # halo wiring: what I send "next" is what my neighbour expects as "prev"
        plo, phi, nlo, nhi = orc.halo_ranges(nl, n, rank, world)
        import torch
        t = torch.tensor([off, off + nl, off + nlo if rank < world - 1 else -1], dtype=torch.int64)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        for r in range(world - 1):
            ok = ok and int(allt[r][1]) == int(allt[r + 1][0])          # bands are contiguous
            ok = ok and int(allt[r][2]) == int(allt[r + 1][0]) - n      # last n rows of r border r+1
        ok = ok and int(allt[0][0]) == 0 and int(allt[-1][1]) == n * n
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 81), (2, 64)])
def test_bootstrap_world2_gloo(world, n):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_partition_rule_matches_oracle():
    import mgpu_bootstrap as mb
    import orc
    for N in (81 * 81, 20000 * 20000, 17, 56576 * 56576):  # the last one: 3.2e9 rows, beyond 32 bits
        for P in (1, 2, 3, 4, 8):
            cover = 0
            for r in range(P):
                assert mb.partition(N, P, r) == orc.partition(N, P, r)
                cover += mb.partition(N, P, r)[0]
            assert cover == N
