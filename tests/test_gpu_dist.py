"""Row-sharded randsvd on 2 GPUs: spawns `torch.distributed.run` on
tests/dist_randsvd_check.py, which checks -- against the CPU oracle and across ranks -- the
matrix-free operator (arithmetic and lattice-table), the q = 0 path, the dense row-sharded
operator (all-reduce of A'Q) and that a rank's block equals the rows of the gathered result.
Skipped on a single-GPU box."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.timeout(600)
def test_sharded_randsvd_two_ranks():
    ndev = _device_count()
    if ndev < 2:
        pytest.skip(f"needs 2 GPUs, this box has {ndev}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "dist_randsvd_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, f"rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
    res = json.loads(lines[-1])
    assert res["ok"] and res["world"] == 2
    for name, c in res["results"].items():
        assert c["sv_rel"] < 1e-10 and c["sine"] < 1e-8, (name, c)
