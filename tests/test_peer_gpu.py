"""ps3d_peer_* on one GPU (include/ps3d.h): the export blob has the documented size, a world of one imports and composites as a
no-op, an import can be undone and done again, and handles that cannot be mapped fail loudly and leave the pipe usable."""
import numpy as np
import pytest

from _compare import render_all
from _scenes_small import SMALL
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline

pytestmark = pytest.mark.gpu


def test_peer_import_reset_import(cuda_lib):
    sc = SMALL["c2_heightfield_small"]()
    want = render_all(cuda_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    blob = pipe.peerExport()
    assert len(blob) == pipe.PEER_BLOB
    pipe.peerImport(0, 1, blob)
    pipe.peerReset()
    pipe.peerReset()                                 # undoing nothing is fine
    pipe.peerImport(0, 1, blob)
    up = scenes.upload(pipe, sc)
    scenes.replay(pipe, sc, up, finish=False)
    pipe.compositePeer()
    pipe.finish()
    assert np.array_equal(pipe.readColour().view(np.uint32), want["colour"].view(np.uint32))


def test_peer_import_of_garbage_handles_fails_loudly(cuda_lib):
    pipe = PuresoftPipeline(64, 64, lib=cuda_lib)
    blob = pipe.peerExport()
    bad = bytes(len(blob)) + bytes(b ^ 0x5A for b in blob)   # rank 1 of 2 given a zeroed "rank 0" blob
    with pytest.raises(Exception):
        pipe.peerImport(1, 2, bad)
    pipe.peerReset()
    pipe.peerImport(0, 1, blob)                      # the pipe is still usable


def test_a_pipe_without_communicators_never_loads_nccl():
    """ps3d_destroy used to ask for libnccl.so.2 (to destroy communicators it did not have): in a process that imports torch LATER
    that puts the system's NCCL in front of the one torch bundles, and `import torch` fails on a missing symbol."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = ("from puresoft3d_b200.pipeline import PuresoftPipeline\n"
            "p = PuresoftPipeline(64, 64, device=0)\np.clearColour(0)\np.finish()\np.close()\n"
            "import torch\nassert torch.cuda.is_available()\nprint('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, PYTHONPATH=ROOT))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
