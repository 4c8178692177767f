"""GPU, >= 2 devices: the NCCL path of poppunk-mod_b200/distributed.py (batches dealt to ranks; one evaluation with
the points sharded and exchanged band-wise) through a real two-rank torchrun.  Skipped on single-GPU boxes, where
tests/test_band_gpu.py covers the same C-ABI steps with a virtual world on one device."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import poppunk_mod_b200 as pk

    return pk._native.device_count()


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_two_ranks_over_nccl():
    env = dict(os.environ, MGPU_SMALL="1")
    env.pop("KDEJS_DEVICE", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29641",
                          os.path.join(ROOT, "tools", "mgpu_check.py")],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "mgpu_check ok" in out.stdout
