"""bench.py's host-side helpers on CPU: the committed ncu traffic table resolves for the bench configurations, the id checksum is
additive over shards (what the N > 1 lines rely on), the merge digest is order-sensitive, and ClockSampler degrades without nvidia-smi."""
import importlib.util
import pathlib

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_traffic_table_resolves_for_the_bench_configurations():
    b = _bench()
    train = {"corpus_bytes": 10999996416, "vocab_size": 32000, "shape": "owt", "seed": 4321, "n_gpus": 1}
    enc = {"text_bytes": 9999998976, "n_gpus": 1}
    for k in ("k_pretok_flags", "k_count_pretokens", "k_merge_loop"):
        tr, src = b.ncu_traffic(k, train)
        assert tr and tr > 1e9 and "ncu --set full" in src, k
    flags_train, _ = b.ncu_traffic("k_pretok_flags", train)
    stage, src = b.ncu_traffic_sum(["k_pretok_flags", "k_special_candidates", "k_special_resolve", "k_popc_words16"], enc)
    flags_enc, _ = b.ncu_traffic("k_pretok_flags", enc)
    assert flags_enc != flags_train and stage > flags_enc and "k_special_candidates" in src       # the two configurations stay apart
    whole, _ = b.ncu_traffic_sum(["k_enc_lookup", "k_pretok_flags", "k_special_candidates", "k_special_resolve", "k_popc_words16",
                                  "k_enc_bpe_short", "k_enc_bpe", "k_enc_scan_emit"], enc)
    lookup, _ = b.ncu_traffic_sum(["k_enc_lookup"], enc)
    assert whole > stage + lookup
    assert b.ncu_traffic("k_enc_lookup", {"text_bytes": 9999998976, "n_gpus": 8}) == (None, None)      # no capture of that configuration
    assert b.ncu_traffic_sum(["k_enc_lookup"], {"text_bytes": 1, "n_gpus": 1}) == (None, None)
    assert b.ncu_traffic_sum(["k_no_such_kernel", "k_enc_lookup"], enc) == (None, None)              # the stage's dominant kernel must be there


def test_ids_checksum_is_additive_over_shards_and_position_sensitive():
    b = _bench()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 32000, (10007,), generator=g, dtype=torch.int32)
    whole = b.ids_checksum(torch, ids, ids.numel(), 0)
    for cuts in ([0, 10007], [0, 1, 5000, 10007], [0, 3333, 3333, 9999, 10007]):
        parts = sum(b.ids_checksum(torch, ids[lo:hi], hi - lo, lo) for lo, hi in zip(cuts, cuts[1:]))
        assert parts & 0xFFFFFFFFFFFFFFFF == whole
    swapped = ids.clone()
    swapped[[10, 11]] = swapped[[11, 10]]
    assert bool(ids[10] != ids[11]) and b.ids_checksum(torch, swapped, swapped.numel(), 0) != whole
    assert b.ids_checksum(torch, ids[:-1], ids.numel() - 1, 0) != whole


def test_merges_sha_depends_on_order_and_on_where_a_pair_is_split():
    b = _bench()
    m = [(b"a", b"b"), (b"ab", b"c")]
    assert b.merges_sha(m) != b.merges_sha(m[::-1])
    assert b.merges_sha([(b"ab", b"c")]) != b.merges_sha([(b"a", b"bc")])
    assert len(b.merges_sha(m)) == 16


def test_clock_sampler_without_nvidia_smi(monkeypatch):
    b = _bench()
    monkeypatch.setenv("PATH", "/nonexistent")
    out = b.ClockSampler(0).start().stop()
    assert out["sm_mhz"] is None and out["reasons"]
