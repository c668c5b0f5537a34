"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: the NCCL-id
broadcast plumbing and the row partition.  The data path itself (NCCL inside
libgsi_b200.so) is exercised on GPUs by tests/dist_randsvd_check.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gsi_b200 as gsi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0's id reaches every rank unchanged (fake generator: no NCCL/GPU on this box)
    fake = bytes((7 * i + 3) % 256 for i in range(128))
    uid = gsi.dist.broadcast_unique_id(make_id=lambda: fake)
    # row partition: contiguous, ordered by rank, covers [0, n), 64-row aligned blocks
    out = {"rank": rank, "uid_ok": uid == fake, "parts": {}}
    for n in (1, 63, 64, 65, 1000, 17472, 200704, 10 ** 6):
        out["parts"][n] = gsi.partition_rows(n, world, rank)
    # a multi-rank context without the id must be rejected on the host side
    try:
        gsi.Context(0, rank, world, None)
        out["ctx_error"] = "none"
    except ValueError:
        out["ctx_error"] = "ValueError"
    except Exception as e:     # pragma: no cover
        out["ctx_error"] = type(e).__name__
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_bootstrap_and_partition_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r["uid_ok"] for r in res)
    assert all(r["ctx_error"] == "ValueError" for r in res)
    for n in res[0]["parts"]:
        nxt = 0
        for r in res:
            r0, ml = r["parts"][n]
            assert r0 == nxt and ml >= 0
            assert r0 % 64 == 0 or ml == 0
            nxt = r0 + ml
        assert nxt == n


def test_partition_rows_single_rank():
    import gsi_b200 as gsi
    assert gsi.partition_rows(12345, 1, 0) == (0, 12345)
