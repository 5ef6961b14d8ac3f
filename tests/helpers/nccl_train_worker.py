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
    # ---- bulk encode, sharded (transformer_lm_b200/sharded_encode.py): the ranks' ids in rank order == the oracle's encode ----
    import numpy as np
    from transformer_lm_b200.synth import synth_host
    from transformer_lm_b200.tokenizer import Tokenizer
    world = dist.get_world_size()
    tok = Tokenizer(dict(got[0]), list(got[1]), ["<|endoftext|>"], ctx=ctx)
    text = synth_host("owt", 4322, 3 << 20)
    want_ids = oracle.OracleTokenizer(dict(got[0]), list(got[1]), ["<|endoftext|>"]).encode_bytes(text.tobytes())
    ids, first, total = tok.encode_sharded(text, np.int32)
    assert total == want_ids.size and np.array_equal(ids, want_ids[first: first + ids.size]), "sharded ids differ on rank %d" % rank
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([ids.size], dtype=torch.int64, device="cuda"))
    assert sum(int(s) for s in sizes) == total and first == sum(int(s) for s in sizes[:rank])
    # the file driver, distributed: one output file assembled from the ranks' segments
    src = out / "enc_in.txt"
    if rank == 0:
        src.write_bytes(text.tobytes())
    dist.barrier()
    from transformer_lm_b200.encode_file import encode_file
    n = encode_file(tok, src, out / "enc_out.bin", np.uint16, piece_bytes=400000, distributed=True)
    assert n == total
    if rank == 0:
        assert np.array_equal(np.fromfile(out / "enc_out.bin", dtype="<u2").astype(np.int64), want_ids)
    dist.barrier()
    (out / ("ok.%d" % rank)).write_text("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
