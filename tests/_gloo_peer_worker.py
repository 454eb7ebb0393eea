"""Worker of tests/test_sortfirst_gloo.py::test_peer_composite_agreement_over_gloo: sortfirst.init_peer_composite's agreement
logic on CPU (gloo) with a stand-in pipe — every rank must come to the same answer whatever fails where:
    mode "ok"            every export and import succeeds                 -> True everywhere
    mode "export_fails"  rank 1 cannot export its handles                 -> False everywhere, nothing imported
    mode "import_fails"  rank 1 cannot map rank 0's targets               -> False everywhere, every rank that imported resets"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from puresoft3d_b200 import sortfirst  # noqa: E402


class FakePipe:
    PEER_BLOB = 256

    def __init__(self, rank, mode):
        self.rank, self.mode, self.imported, self.resets, self.blobs = rank, mode, False, 0, None

    def peerExport(self):
        if self.mode == "export_fails" and self.rank == 1:
            raise RuntimeError("no IPC handle")
        return bytes([self.rank + 1]) * self.PEER_BLOB

    def peerImport(self, rank, world, blobs):
        if self.mode == "import_fails" and self.rank == 1:
            raise RuntimeError("cudaIpcOpenMemHandle: no peer access")
        self.imported, self.blobs = True, blobs

    def peerReset(self):
        self.imported = False
        self.resets += 1


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    mode, out_path = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    os.environ.pop("PS3D_SORTFIRST_COMPOSITE", None)
    pipe = FakePipe(rank, mode)
    got = sortfirst.init_peer_composite(pipe, rank, world, torch.device("cpu"))
    ok = got == (mode == "ok")
    if mode == "ok":
        want = b"".join(bytes([r + 1]) * FakePipe.PEER_BLOB for r in range(world))      # every rank's blob, rank 0's first
        ok = ok and pipe.imported and pipe.blobs == want and pipe.resets == 0
    elif mode == "export_fails":
        ok = ok and not pipe.imported and pipe.resets == 0
    else:
        ok = ok and not pipe.imported and pipe.resets == 1
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(out_path, "w") as f:
            f.write("ok" if int(flag.item()) == 1 else "some rank disagreed (mode %s)" % mode)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
