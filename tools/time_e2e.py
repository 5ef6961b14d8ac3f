"""GPU: time bpe_train (C ABI) from page-locked host memory.  usage: python tools/time_e2e.py <shape> <seed> <bytes> <vocab> [reps]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
shape, seed, n, vocab = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3])) // 4096 * 4096, int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
import _bootstrap, torch, numpy as np
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
ctx = _lib.default_context(0)
L = _lib.lib()
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device(shape, seed, n, t.data_ptr(), ctx=ctx)
host = _lib.PinnedBuffer(n)
torch.from_numpy(host.array).copy_(t)
del t
torch.cuda.synchronize()
n_merges = vocab - 257
pairs = np.zeros((n_merges, 2), dtype=np.int32)
sp_blob, sp_offs = _lib.pack_blobs([b"<|endoftext|>"])
for it in range(reps):
    n_done = C.c_int(0); stats = _lib.TrainStats()
    t0 = time.time()
    ctx.check(L.bpe_train(ctx.handle, _lib.ptr(host.array), n, _lib.ptr(sp_blob), _lib.ptr(sp_offs), 1, n_merges, _lib.ptr(pairs), C.byref(n_done), C.byref(stats)))
    dt = time.time() - t0
    print("bpe_train from pinned host: %.1f ms" % (dt * 1e3), {k: round(v, 1) for k, v in stats.as_dict().items() if k.startswith("ms_")})
