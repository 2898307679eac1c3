"""Multi-GPU parity (needs >= 2 B200s, e.g. `gpurun --gpus 2`; skipped on a single-GPU box): the row-partitioned
path -- both the in-kernel peer-memory variant and the NCCL baseline -- against the single-GPU kernel AND the compiled oracle.
The host-side partition logic and data flow are covered on CPU by tests/test_partition_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_row_partitioned_paths_match_single_gpu(lib):
    n = 2 if _ngpus() < 4 else 4
    env = dict(os.environ, QPB_DIST_MODES="peer,nccl")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29617",
                          os.path.join(ROOT, "scripts", "dist_check.py"), "0.01", "0"],
                         env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count('"all_ranks_ok": true') == 6, out.stdout[-2000:]
