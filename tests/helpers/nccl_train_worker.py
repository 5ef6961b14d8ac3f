"""torchrun worker: sharded training over NCCL must reproduce the single-GPU result and the oracle's."""
import os
import pathlib
import sys

import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import _bootstrap  # noqa: F401,E402
from oracle import oracle  # noqa: E402
from transformer_lm_b200 import _lib  # noqa: E402
from transformer_lm_b200.sharded import train_bpe_sharded  # noqa: E402
from transformer_lm_b200.train import train_bpe  # noqa: E402


def main():
    out = pathlib.Path(sys.argv[1])
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    path = ROOT / "tests" / "fixtures" / "corpus.en"
    got = train_bpe_sharded(path, 1000, ["<|endoftext|>"], ctx=ctx)
    single = train_bpe(path, 1000, ["<|endoftext|>"], ctx=ctx)
    want = oracle.train_bpe(path, 1000, ["<|endoftext|>"])
    assert got[1] == single[1] == want[1], "merges differ on rank %d" % rank
    assert got[0] == single[0] == want[0]
    # a file with carriage returns takes the unsharded path on every rank
    p2 = out / ("crlf.%d.txt" % rank)
    p2.write_bytes(b"hello world\r\nthe quick brown fox\r" * 500)
    got2 = train_bpe_sharded(p2, 300, [], ctx=ctx)
    assert got2 == oracle.train_bpe(p2, 300, [])
    dist.barrier()
    (out / ("ok.%d" % rank)).write_text("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
