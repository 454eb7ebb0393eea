"""Sort-first sharding on CPU: band arithmetic, and a world_size-2/3 frame over gloo where each rank
renders its raster-row band and the product's compositor gathers the bands onto rank 0. No pixel is touched by two
ranks, so the composite must equal a single-rank frame bit for bit and the per-rank fragment counts must add up."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from puresoft3d_b200 import sortfirst


@pytest.mark.parametrize("height", [1, 15, 16, 17, 131, 240, 1080, 2160, 4320])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_row_bands_partition_the_frame(height, world):
    bands = sortfirst.row_bands(height, world)
    assert len(bands) == world
    assert bands[0][0] == 0 and bands[-1][1] == height
    for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
        assert a1 == b0 and a0 <= a1
    for r0, r1 in bands:
        assert r0 % sortfirst.TILE == 0 or r0 == height   # bands start on tile rows: a tile is never shared by two ranks
        m0, m1 = sortfirst.memory_rows((r0, r1), height)
        assert m1 - m0 == r1 - r0 and 0 <= m0 <= m1 <= height


@pytest.mark.parametrize("units", [0, 1, 7, 36, 3007584])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_units_cover_the_stream(units, world):
    per, rem = sortfirst.shard_units(units, world)
    assert per * world + rem == units and 0 <= rem < world


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,scene", [(2, "c1_cube_def01"), (2, "c4_blend_overdraw"), (3, "soup_odd_size")])
def test_sort_first_frame_over_gloo(world, scene, tmp_path, built):
    out = tmp_path / "result.txt"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
    procs = []
    for rank in range(world):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), scene, str(out)], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for pr in procs:
        try:
            o, _ = pr.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
        assert pr.returncode == 0, o[-2000:]
    assert out.read_text() == "ok", out.read_text() + "\n" + "\n".join(l[-500:] for l in logs)


@pytest.mark.parametrize("mode", ["ok", "export_fails", "import_fails"])
def test_peer_composite_agreement_over_gloo(mode, tmp_path):
    """sortfirst.init_peer_composite with a stand-in pipe: whatever fails on whichever rank, every rank ends with the same answer
    (and a rank that had already mapped rank 0's targets lets go of them)."""
    out = tmp_path / "result.txt"
    world = 3
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_peer_worker.py"), mode, str(out)],
                              env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for rank in range(world)]
    for pr in procs:
        try:
            o, _ = pr.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert pr.returncode == 0, o[-2000:]
    assert out.read_text() == "ok", out.read_text()
