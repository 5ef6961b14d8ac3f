"""GPU parity of the training-batch feeder (bpe_batch_windows_dev behind models.util.load_batch) and of the batched decode
(bpe_decode_batch behind Tokenizer.decode_batch) -- SURVEY 8f row 4.  The checker for load_batch is the reference's own
few lines (models/util.py:37-57) restated with numpy: same torch generator, same draws, same windows."""
import numpy as np
import pytest
import torch

import _bootstrap  # noqa: F401
from models.util import load_batch
from oracle import oracle
from tests.adapters import get_tokenizer
from tests.common import load_gpt2_fixture
from transformer_lm_b200 import _lib

pytestmark = pytest.mark.gpu


def reference_load_batch(dataset, batch_size, context_length, generator=None):
    # models/util.py:44-57
    inputs = np.zeros((batch_size, context_length))
    targets = np.zeros((batch_size, context_length))
    limit = len(dataset) - context_length
    start_idx = torch.randint(limit, (batch_size,), generator=generator)
    for row, idx in enumerate(start_idx):
        inputs[row] = dataset[idx: idx + context_length]
        targets[row] = dataset[idx + 1: idx + context_length + 1]
    return torch.tensor(inputs, dtype=torch.long), torch.tensor(targets, dtype=torch.long)


@pytest.mark.parametrize("dtype", [np.uint16, np.int32, np.int64])
@pytest.mark.parametrize("batch,context", [(1, 1), (4, 7), (32, 256), (3, 1000)])
def test_load_batch_matches_reference(dtype, batch, context):
    rng = np.random.default_rng(batch * 1000 + context)
    hi = 65535 if dtype == np.uint16 else 100000
    data = rng.integers(0, hi, size=5000, dtype=np.int64).astype(dtype)
    g1, g2 = torch.Generator().manual_seed(1234), torch.Generator().manual_seed(1234)
    for _ in range(3):                         # consecutive draws advance the generator like the reference's
        x, y = load_batch(data, batch, context, "cuda:0", generator=g1)
        wx, wy = reference_load_batch(data, batch, context, generator=g2)
        assert x.dtype == torch.long and y.dtype == torch.long and x.device.type == "cuda"
        assert torch.equal(x.cpu(), wx) and torch.equal(y.cpu(), wy)


def test_load_batch_edges_and_errors():
    data = np.arange(10, dtype=np.uint16)
    # context = len - 1: the only legal start is 0 (randint(1))
    x, y = load_batch(data, 2, 9, "cuda:0")
    assert x.tolist() == [list(range(9))] * 2 and y.tolist() == [list(range(1, 10))] * 2
    with pytest.raises(RuntimeError):          # torch.randint(0, ...) raises for context >= len, like the reference
        load_batch(data, 1, 10, "cuda:0")
    with pytest.raises(_lib.BpeError):
        load_batch(data, 1, 4, "cpu")
    # the C ABI refuses windows that leave the array
    ctx = _lib.default_context(0)
    t = torch.arange(10, dtype=torch.int32, device="cuda")
    x = torch.empty(4, dtype=torch.long, device="cuda")
    y = torch.empty(4, dtype=torch.long, device="cuda")
    import ctypes as C
    for bad in (7, -1):
        starts = np.array([bad], dtype=np.int64)
        rc = _lib.lib().bpe_batch_windows_dev(ctx.handle, C.c_void_p(t.data_ptr()), _lib.DTYPE_I32, 10, _lib.ptr(starts), 1, 4,
                                              C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()))
        assert rc == _lib.ERR_ARG
    starts = np.array([5], dtype=np.int64)
    rc = _lib.lib().bpe_batch_windows_dev(ctx.handle, C.c_void_p(t.data_ptr()), _lib.DTYPE_I32, 10, _lib.ptr(starts), 1, 4,
                                          C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()))
    assert rc == _lib.BPE_OK and x.tolist() == [5, 6, 7, 8] and y.tolist() == [6, 7, 8, 9]


def test_load_batch_large_resident_array():
    # 64 M tokens stay resident between calls; windows checked against numpy slices
    n = 64 << 20
    data = (np.arange(n, dtype=np.uint32) * 2654435761 >> 16).astype(np.uint16)
    g = torch.Generator().manual_seed(7)
    g2 = torch.Generator().manual_seed(7)
    for _ in range(2):
        x, y = load_batch(data, 64, 1024, "cuda:0", generator=g)
        starts = torch.randint(n - 1024, (64,), generator=g2).numpy()
        xs, ys = x.cpu().numpy(), y.cpu().numpy()
        for r in (0, 17, 63):
            assert np.array_equal(xs[r], data[starts[r]: starts[r] + 1024].astype(np.int64))
            assert np.array_equal(ys[r], data[starts[r] + 1: starts[r] + 1025].astype(np.int64))


def test_decode_batch_matches_decode():
    vocab, merges = load_gpt2_fixture()
    tok = get_tokenizer(vocab, merges, ["<|endoftext|>"])
    texts = ["", "Hello, how are you?", "Héllò hôw are ü? 🙃", "<|endoftext|>", "a" * 5000, "once upon a time\n\n"]
    seqs = [tok.encode(t) for t in texts]
    seqs.append([8582, 247])                   # a split emoji: U+FFFD replacement per sequence (SURVEY A-17)
    seqs.append([])
    otok = oracle.OracleTokenizer(dict(vocab), list(merges), ["<|endoftext|>"])
    assert tok.decode_batch(seqs) == [otok.decode(s) for s in seqs]      # the oracle's b"".join(...).decode("utf-8", "replace")
    assert tok.decode_batch(seqs) == [tok.decode(s) for s in seqs]
    assert tok.decode_batch([]) == []
    assert tok.decode_batch([[], []]) == ["", ""]
    with pytest.raises(KeyError) as ei:
        tok.decode_batch([[1, 2], [3, 999999]])
    assert ei.value.args == (999999,)


def test_get_batch_like_the_reference_test():
    # tests/test_data.py:11-60 of the reference, on the GPU: shapes, y = x + 1 on an arange dataset, start indices cover
    # [0, len - context) uniformly (mean +/- 5 sigma)
    import math
    from collections import Counter
    from tests.adapters import run_get_batch
    dataset = np.arange(0, 100)
    context_length, batch_size, num_iters = 7, 32, 1000
    starting_indices = Counter()
    for _ in range(num_iters):
        x, y = run_get_batch(dataset=dataset, batch_size=batch_size, context_length=context_length, device="cuda:0")
        assert x.shape == (batch_size, context_length) and y.shape == (batch_size, context_length)
        assert torch.equal(x + 1, y)
        starting_indices.update(x[:, 0].tolist())
    n_starts = len(dataset) - context_length
    assert max(starting_indices) == n_starts - 1 and min(starting_indices) == 0
    expected = (num_iters * batch_size) / n_starts
    sigma = math.sqrt((num_iters * batch_size) * (1 / n_starts) * (1 - (1 / n_starts)))
    for count in starting_indices.values():
        assert expected - 5 * sigma < count < expected + 5 * sigma


def test_load_batch_device_without_index_hits_the_resident_copy():
    """The reference's trainer passes torch.device("cuda") (train.py:226), which is != torch.device("cuda:0"): the resident
    copy must still be found, the result must be ordered with torch's stream, and models.util must export what train.py:17 imports."""
    from models.util import load_checkpoint, save_checkpoint  # noqa: F401
    from transformer_lm_b200 import batch as B
    data = np.arange(0, 5000, dtype=np.uint16)
    x, y = load_batch(data, 8, 16, torch.device("cuda"), generator=torch.Generator().manual_seed(1))
    n_cached = len(B._resident)
    tokens_a = B._resident[(id(data), torch.cuda.current_device())][1]
    for dev in ("cuda", "cuda:0", torch.device("cuda")):
        x, y = load_batch(data, 8, 16, dev, generator=torch.Generator().manual_seed(1))
        assert len(B._resident) == n_cached and B._resident[(id(data), torch.cuda.current_device())][1] is tokens_a
        assert torch.equal(x + 1, y) and x.device.type == "cuda"
    # on a side stream with queued work: the gather is ordered with torch's current stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        junk = torch.zeros(1 << 24, device="cuda").cumsum(0)
        x, y = load_batch(data, 64, 128, "cuda", generator=torch.Generator().manual_seed(2))
        z = (x + 1 - y).abs().sum()
    s.synchronize()
    assert int(z) == 0 and junk.numel()
