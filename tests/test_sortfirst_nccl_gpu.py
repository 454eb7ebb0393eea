"""The real N-GPU composite: rank bands rendered on separate B200s, composited onto rank 0 by the library's
own exchange step, compared on rank 0 with the golden fixture (reference build) and with the single-GPU frame. Skips below
two GPUs (the driver's 1-GPU box); `gpurun --gpus 2 -- python -m pytest tests/test_sortfirst_nccl_gpu.py -m gpu` runs it."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["peer", "native"])
@pytest.mark.parametrize("scene", ["c2_heightfield_small", "c1_cube_def03", "soup_odd_size"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_composite_on_real_gpus_matches_golden(scene, world, mode, tmp_path, built):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    out = tmp_path / "result.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_nccl_worker.py"), scene, str(out), mode]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert out.read_text().startswith("ok"), out.read_text()
