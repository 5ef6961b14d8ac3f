"""GPU (one device): the per-rank half of sharded training on the shard rank 0 of N would own.  usage: time_shard.py <N> [bytes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, torch
import bench
from transformer_lm_b200 import _lib, sharded
N = int(sys.argv[1]); total = int(float(sys.argv[2])) if len(sys.argv) > 2 else int(11e9)
total = total // 4096 * 4096
ctx = _lib.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
text, meta = bench.make_shard(torch, ctx, "owt", 4321, total, 0, N)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = sharded.DeviceCounter(ctx)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    assert c.add(None, meta["own_begin"], meta["own_end"], meta["at_start"], meta["at_end"], device_ptr=text.data_ptr(), n_bytes=meta["n"]) is None
    torch.cuda.synchronize(); t2 = time.perf_counter()
    pt = c.pair_table(["<|endoftext|>"])
    torch.cuda.synchronize(); t3 = time.perf_counter()
    blob, offs, counts = c.export()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("iter %d: begin %.2f ms, add %.2f, pair_table %.2f, export %.2f (%d words, %d bytes)" % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, counts.numel(), blob.numel()))
