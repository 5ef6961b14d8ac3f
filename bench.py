#!/usr/bin/env python3
"""Benchmark of the byte-level BPE hot path on B200 (contract and definitions: DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|train-tiny|encode] [--bytes B] [--vocab V]
    torchrun --nproc-per-node N bench.py --gpus N ...       # one rank per GPU, NCCL
    python bench.py --impl reference ...                    # CPU port of the reference on a bounded sample

Default workload: BPE training on the 11 GB synthetic OWT-shaped corpus, vocab 32000, special token
<|endoftext|> (BASELINE.json configs[3]; the shape of perf/bpe/owt.py:4-8 in the reference).  One "step" is one
complete train_bpe over the corpus (pretokenise, count, exchange when N > 1, 31 743 merges).
  value  = corpus MB per second of a whole training run, text already resident in HBM, CUDA-event time
  e2e    = the same through the public API with the text in pinned HOST memory (H2D inside the timed region)
The line also carries an "encode" object: bulk encode of the 10 GB OWT-shaped text with the vocab just trained
(BASELINE.json configs[4]); `--workload encode` makes that the headline instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SPECIALS = ["<|endoftext|>"]
L2_BYTES = 126 << 20
BLOCK = 4096

WORKLOADS = {
    # name: (shape, seed, default bytes, default vocab, description)
    "train": ("owt", 4321, 11e9, 32000, "BPE train, 11 GB synthetic OWT-shape corpus, vocab 32000, special <|endoftext|>"),
    "train-tiny": ("tinystories", 1234, 2 * 2**30, 10000, "BPE train, 2 GiB synthetic TinyStories-shape corpus, vocab 10000, special <|endoftext|>"),
    "encode": ("owt", 4322, 10e9, 32000, "bulk encode to uint16, 10 GB synthetic OWT-shape text, 32K vocab trained on the OWT-shape train corpus"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=list(WORKLOADS))
    ap.add_argument("--bytes", type=float, default=None, help="corpus size in bytes (default: the BASELINE config)")
    ap.add_argument("--encode-bytes", type=float, default=None, help="size of the encode text (default 10e9)")
    ap.add_argument("--vocab-train-bytes", type=float, default=1e9, help="--workload encode: slice of the train corpus the vocab is trained on")
    ap.add_argument("--vocab", type=int, default=None)
    ap.add_argument("--ref-sample-bytes", type=float, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encode", action="store_true", help="train workloads: skip the secondary encode measurement")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (C port of the reference) on a bounded sample, host cores only
# ----------------------------------------------------------------------------------------------
def cpu_train_sample(shape, seed, sample_bytes, vocab):
    import _bootstrap  # noqa: F401
    from oracle import oracle
    from transformer_lm_b200.synth import synth_host
    n = max(int(sample_bytes) // BLOCK * BLOCK, BLOCK)
    data = synth_host(shape, seed, n).tobytes()
    t0 = time.perf_counter()
    vocab_d, merges = oracle.train_bpe_on_bytes(data, vocab, SPECIALS)
    return n, time.perf_counter() - t0, vocab_d, merges


def cpu_encode_sample(tok_vocab, tok_merges, shape, seed, sample_bytes):
    import _bootstrap  # noqa: F401
    from oracle import oracle
    from transformer_lm_b200.synth import synth_host
    tok = oracle.OracleTokenizer(dict(tok_vocab), list(tok_merges), SPECIALS)
    n = max(int(sample_bytes) // BLOCK * BLOCK, BLOCK)
    data = synth_host(shape, seed, n).tobytes()
    t0 = time.perf_counter()
    ids = tok.encode_bytes(data)
    return n, time.perf_counter() - t0, len(ids)


def run_reference(args):
    """--impl reference: the reference's algorithm (its C port, oracle/bpe_oracle.c -- the reference itself is pure
    Python and does not travel to the GPU box) on the box's host cores.  The reference is single-threaded
    (SURVEY 2.1), so this uses one core.  Each step is a bounded sample of the workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    shape, seed, nbytes, vocab, desc = WORKLOADS[args.workload]
    vocab = args.vocab or vocab
    times = []
    if args.workload == "encode":
        n_t, _, tv, tm = cpu_train_sample("owt", 4321, 1 << 20, min(vocab, 4000))
        sample = args.ref_sample_bytes or 16e6
        for i in range(args.warmup + args.steps):
            n, dt, n_ids = cpu_encode_sample(tv, tm, shape, seed, sample)
            if i >= args.warmup:
                times.append(dt)
        metric = "bpe_encode_MBps"
        sample_desc = "first %.1f MB of the encode text; tokenizer = port-trained vocab %d on a 1 MiB slice of the train corpus" % (n / 1e6, min(vocab, 4000))
    else:
        sample = args.ref_sample_bytes or (1 << 19)
        for i in range(args.warmup + args.steps):
            n, dt, _, merges = cpu_train_sample(shape, seed, sample, vocab)
            if i >= args.warmup:
                times.append(dt)
        metric = "bpe_train_MBps"
        sample_desc = "first %.2f MB of the corpus (same generator and seed), full vocab %d => %d merges" % (n / 1e6, vocab, len(merges))
    ms = 1e3 * sum(times) / len(times)
    value = n / 1e6 / (ms / 1e3)
    line = {
        "impl": "reference", "metric": metric, "value": round(value, 4), "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": {"workload": desc, "sample_bytes": n},
        "cpu_baseline": {"value": round(value, 4), "unit": "MB/s", "cores": 1, "kind": "port", "sample": sample_desc,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": round(value, 4), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C port of the reference's algorithm (oracle/bpe_oracle.c: regex-equivalent matcher, dict counts, O(pairs) max() "
                "scan per merge), one host thread like the reference",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
class DeviceShards:
    """Shard source for sharded_count over text this rank already holds in HBM (synthetic corpus blocks
    [b0 - left halo, b1 + right halo) of the global corpus)."""

    def __init__(self, ptr, n_bytes, own_begin, own_end, at_start, at_end, base):
        self.d = dict(data=None, device_ptr=ptr, n_bytes=n_bytes, own_begin=own_begin, own_end=own_end, at_start=at_start, at_end=at_end, base=base)

    def load(self, rank, world, halo_right):
        return self.d

    def raise_decode_error(self, offset):
        raise UnicodeDecodeError("utf-8", b"", 0, 1, "synthetic corpus invalid at byte %d" % offset)


def make_shard(torch, ctx, shape, seed, total_bytes, rank, world, halo_blocks=64):
    from transformer_lm_b200.synth import synth_device
    nb = total_bytes // BLOCK
    b0, b1 = nb * rank // world, nb * (rank + 1) // world
    r0, r1 = max(0, b0 - 1), min(nb, b1 + halo_blocks)
    n = (r1 - r0) * BLOCK
    t = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    synth_device(shape, seed, n, t.data_ptr(), ctx=ctx, first_block=r0)
    torch.cuda.synchronize()
    return t, dict(n=n, own_begin=(b0 - r0) * BLOCK, own_end=(b1 - r0) * BLOCK, at_start=r0 == 0, at_end=r1 == nb, base=r0 * BLOCK)


def encode_shard_cut(torch, t, meta, is_first, is_last):
    """Encode shards are cut at special-token occurrences (Tokenizer.segment splits there first, tokenizer.py:63-66, so
    the cut is exact): the owned range starts/ends at the first <|endoftext|> at or after the nominal block boundary."""
    pat = torch.tensor(list(SPECIALS[0].encode()), dtype=torch.uint8, device=t.device)

    def first_special(frm):
        win = t[frm: frm + (1 << 20)]
        m = torch.ones(win.numel() - pat.numel() + 1, dtype=torch.bool, device=t.device)
        for i in range(pat.numel()):
            m &= win[i: i + m.numel()] == pat[i]
        idx = torch.nonzero(m)
        if idx.numel() == 0:
            raise RuntimeError("no <|endoftext|> within 1 MiB of the shard boundary")
        return frm + int(idx[0])

    lo = meta["own_begin"] if is_first else first_special(meta["own_begin"])
    hi = meta["own_end"] if is_last else first_special(meta["own_end"])
    return lo, hi


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import _bootstrap  # noqa: F401
    from transformer_lm_b200 import _lib, sharded
    from transformer_lm_b200.tokenizer import Tokenizer
    from transformer_lm_b200.train import train_bpe_on_bytes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.use_stream(stream.cuda_stream)                           # library work, NCCL and the timing events share one stream
    L = _lib.lib()
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the stream everything runs on; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = []
        for _ in range(steps):
            out.append(fn())
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        return max_over_ranks(e0.elapsed_time(e1)) / steps, max_over_ranks(wall * 1e3) / steps, out

    primary = args.workload
    is_train = primary != "encode"
    shape, seed, nbytes, vocab_size, desc = WORKLOADS[primary]
    vocab_size = args.vocab or vocab_size
    peak, peak_src = measured_peak_gbs()
    line = {}

    # ======================= training =======================
    train_res = None
    if is_train:
        nbytes = int(args.bytes or nbytes) // BLOCK * BLOCK
        text, meta = make_shard(torch, ctx, shape, seed, nbytes, rank, world)

        def train_step():
            if world == 1:
                return train_bpe_on_bytes(None, vocab_size, SPECIALS, ctx=ctx, return_stats=True, device_ptr=text.data_ptr(), n_bytes=meta["n"])
            counter = sharded.DeviceCounter(ctx)
            src = DeviceShards(text.data_ptr(), meta["n"], meta["own_begin"], meta["own_end"], meta["at_start"], meta["at_end"], meta["base"])
            t_a = time.perf_counter()
            out = sharded.sharded_count(counter, src, SPECIALS, None, True)
            assert out == "ok"
            torch.cuda.synchronize()
            t_b = time.perf_counter()
            v, m, st = counter.finish(vocab_size, SPECIALS, return_stats=True)
            st["ms_count_exchange_wall"] = (t_b - t_a) * 1e3
            if sharded.LAST_TIMES:
                st["exchange_profile_ms"] = {k: round(v, 2) for k, v in sharded.LAST_TIMES.items()}
            return v, m, st

        for _ in range(args.warmup):
            train_res = train_step()
        launches0 = L.bpe_launch_count()
        sampler = ClockSampler(local_rank).start()
        ms_dev, ms_wall, outs = timed(train_step, args.steps)
        clocks = sampler.stop()
        launches = L.bpe_launch_count() - launches0
        train_res = outs[-1]
        merges = train_res[1]
        if world > 1:
            d = sharded.merges_digest(merges)
            lo, hi = torch.tensor([d], device=dev), torch.tensor([d], device=dev)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert int(lo) == int(hi), "ranks disagree on the merge list"
        stats = [o[2] for o in outs]

        def avg(key):
            vals = [s[key] for s in stats if key in s]
            return sum(vals) / len(vals) if vals else 0.0

        st = stats[-1]
        stages = {k: round(avg(k), 3) for k in ("ms_h2d", "ms_pretok", "ms_count", "ms_build", "ms_merge", "ms_total", "ms_count_exchange_wall")
                  if any(k in s for s in stats)}
        if "exchange_profile_ms" in st:
            stages["exchange_profile_ms"] = st["exchange_profile_ms"]
        value = nbytes / 1e6 / (ms_dev / 1e3)
        # roofline of the dominant kernel, the persistent merge loop.  Algorithmic bytes: the reference's
        # max() reads every live pair-table entry (16 B: packed pair + count) at every step, plus the rewritten symbols.
        merge_ms = avg("ms_merge")
        alg_merge = 16.0 * st["sum_live_pairs"] + 8.0 * st["log_records"]
        ach_merge = alg_merge / 1e9 / (merge_ms / 1e3) if merge_ms > 0 else 0.0
        local_bytes = meta["own_end"] - meta["own_begin"]
        pretok_ms, count_ms = avg("ms_pretok"), avg("ms_count")
        # DRAM traffic of that kernel from the committed ncu capture of this exact configuration (null otherwise)
        traffic, traffic_src = None, None
        try:
            tj = json.loads((ROOT / "profiles" / "r1_k_merge_loop_11GB.json").read_text())
            if world == 1 and all(tj["config"][k] == v for k, v in (("corpus_bytes", nbytes), ("vocab_size", vocab_size), ("shape", shape), ("seed", seed))):
                traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "kernel": "k_merge_loop (one persistent cooperative launch, %d merges)" % len(merges),
            "achieved": round(ach_merge, 1), "peak": peak, "unit": "GB/s", "frac": round(ach_merge / peak, 4), "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_source": peak_src, "frac_of_nominal_8000_GBps": round(ach_merge / 8000.0, 4), "algorithmic_bytes_per_launch": alg_merge,
            "algorithmic_bytes_definition": "16 B x sum over merges of live pair-table keys (%d) + 8 B x index records (%d)" % (st["sum_live_pairs"], st["log_records"]),
            "share_of_step": round(merge_ms / ms_dev, 4) if ms_dev else None,
            "us_per_merge": round(1e3 * merge_ms / max(len(merges), 1), 3),
            "other_kernels": {
                "k_pretok_flags": {"algorithmic_bytes": local_bytes * 1.125, "ms": round(pretok_ms, 3),
                                   "achieved_GBps": round(local_bytes * 1.125 / 1e9 / (pretok_ms / 1e3), 1) if pretok_ms else None,
                                   "frac": round(local_bytes * 1.125 / 1e9 / (pretok_ms / 1e3) / peak, 4) if pretok_ms else None,
                                   "note": "stage time: includes the UTF-8 / CR error read-back"},
                "k_count_pretokens": {"algorithmic_bytes": local_bytes * 1.125, "ms": round(count_ms, 3),
                                      "achieved_GBps": round(local_bytes * 1.125 / 1e9 / (count_ms / 1e3), 1) if count_ms else None,
                                      "frac": round(local_bytes * 1.125 / 1e9 / (count_ms / 1e3) / peak, 4) if count_ms else None,
                                      "note": "stage time: includes hash-table growth / rehash kernels"}},
        }
        line = {
            "metric": "bpe_train_MBps", "value": round(value, 2), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_dev, 3), "wall_ms_per_step": round(ms_wall, 3), "train_wall_s": round(ms_wall / 1e3, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc if not args.bytes else desc + " [--bytes %d]" % nbytes, "corpus_bytes": nbytes, "vocab_size": vocab_size,
                       "merges": len(merges), "shape": shape, "seed": seed,
                       "l2_policy": "per-rank input (%.2f GB) larger than L2 (126 MB)" % (local_bytes / 1e9) if local_bytes > L2_BYTES else "input smaller than L2",
                       "unicode_tables": L.bpe_unicode_table_source().decode(),
                       "parallelism": "1 GPU" if world == 1 else "%d ranks: byte-range shards + halo, NCCL all-gather of count tables, pair-table all-reduce check, replicated merge loop" % world},
            "stages_ms": stages,
            "counts": {k: st[k] for k in ("n_pretokens", "n_unique", "n_symbols", "n_pairs_initial", "n_pairs_final", "log_records", "sum_live_pairs")},
            "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
        }

        # ---- end to end: text in pinned host memory, H2D inside the timed region, merges read back ----
        if not args.no_e2e:
            n_loc = meta["n"]
            host = _lib.PinnedBuffer(n_loc)
            torch.from_numpy(host.array).copy_(text[:n_loc])
            torch.cuda.synchronize()

            def e2e_step():
                if world == 1:
                    return train_bpe_on_bytes(host.array, vocab_size, SPECIALS, ctx=ctx, return_stats=True)
                counter = sharded.DeviceCounter(ctx)

                class HostShards(DeviceShards):
                    pass
                src = HostShards(None, n_loc, meta["own_begin"], meta["own_end"], meta["at_start"], meta["at_end"], meta["base"])
                src.d["data"] = host.array
                assert sharded.sharded_count(counter, src, SPECIALS, None, True) == "ok"
                return counter.finish(vocab_size, SPECIALS, return_stats=True)

            e2e_step()
            _, ms_wall2, outs2 = timed(e2e_step, args.steps)
            assert outs2[-1][1] == merges
            line["e2e"] = {"value": round(nbytes / 1e6 / (ms_wall2 / 1e3), 2), "unit": "MB/s", "h2d_bytes_per_step": int(n_loc),
                           "d2h_bytes_per_step": 8 * len(merges), "ms_per_step": round(ms_wall2, 3),
                           "api": "train_bpe_on_bytes(pinned host buffer) -> bpe_train (C ABI); wall clock, includes building the python vocab/merges"}
            host.free()
        del text
        torch.cuda.empty_cache()

    # ======================= encoding =======================
    if (not is_train) or not args.no_encode:
        e_shape, e_seed, e_bytes, _, e_desc = WORKLOADS["encode"]
        e_bytes = int(args.encode_bytes or (args.bytes if not is_train and args.bytes else e_bytes)) // BLOCK * BLOCK
        if train_res is None:
            vt_bytes = int(args.vocab_train_bytes) // BLOCK * BLOCK
            tt = torch.empty(vt_bytes, dtype=torch.uint8, device="cuda")
            from transformer_lm_b200.synth import synth_device
            synth_device("owt", 4321, vt_bytes, tt.data_ptr(), ctx=ctx)
            train_res = train_bpe_on_bytes(None, vocab_size, SPECIALS, ctx=ctx, return_stats=True, device_ptr=tt.data_ptr(), n_bytes=vt_bytes)
            del tt
            vocab_src = "trained on the first %.2f GB of the OWT-shape train corpus" % (vt_bytes / 1e9)
        else:
            vocab_src = "the vocab trained above"
        tok = Tokenizer(dict(train_res[0]), list(train_res[1]), SPECIALS, ctx=ctx)
        h = tok._device_tok()
        etext, emeta = make_shard(torch, ctx, e_shape, e_seed, e_bytes, rank, world, halo_blocks=256)
        lo, hi = encode_shard_cut(torch, etext, emeta, rank == 0, rank == world - 1)
        n_loc = hi - lo
        out = torch.empty(max(n_loc, 1), dtype=torch.uint16, device="cuda")
        src_ptr = etext.data_ptr() + lo
        enc_stats = []

        def encode_step():
            # every step does the whole job: the pretoken -> ids cache is dropped first, so all BPE merges are recomputed
            ctx.check(L.bpe_tok_cache_reset(h))
            n_out, stt = C.c_uint64(0), _lib.EncodeStats()
            ctx.check(L.bpe_encode_dev(h, C.c_void_p(src_ptr), n_loc, _lib.DTYPE_U16, C.c_void_p(out.data_ptr()), out.numel(), C.byref(n_out), C.byref(stt)))
            enc_stats.append(stt.as_dict())
            return n_out.value

        for _ in range(max(args.warmup, 1)):
            encode_step()
        enc_stats.clear()
        l0 = L.bpe_launch_count()
        sampler = ClockSampler(local_rank).start()
        ems_dev, ems_wall, n_tok = timed(encode_step, args.steps)
        eclocks = sampler.stop()
        e_launches = L.bpe_launch_count() - l0
        tokens_local = n_tok[-1]
        tok_total = tokens_local
        bytes_total = n_loc
        if world > 1:
            t = torch.tensor([tokens_local, n_loc], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            tok_total, bytes_total = int(t[0]), int(t[1])
            assert bytes_total == e_bytes, (bytes_total, e_bytes)

        def eavg(key):
            return sum(s[key] for s in enc_stats) / len(enc_stats)

        alg = n_loc + 2.0 * tokens_local                       # text read once + one uint16 per token
        dom = max(("ms_pretok", "ms_lookup", "ms_bpe", "ms_emit"), key=eavg)
        enc = {
            "metric": "bpe_encode_MBps", "value": round(e_bytes / 1e6 / (ems_dev / 1e3), 2), "unit": "MB/s", "ms_per_step": round(ems_dev, 3),
            "config": {"workload": e_desc if e_bytes == 10e9 else e_desc + " [%d bytes]" % e_bytes, "text_bytes": e_bytes, "vocab": vocab_src,
                       "tokens": tok_total, "bytes_per_token": round(e_bytes / max(tok_total, 1), 3), "cache": "pretoken cache reset at the start of every step",
                       "parallelism": "1 GPU" if world == 1 else "%d ranks, shards cut at <|endoftext|>, no data collective" % world},
            "stages_ms": {k: round(eavg(k), 3) for k in ("ms_h2d", "ms_pretok", "ms_lookup", "ms_bpe", "ms_emit", "ms_total")},
            "new_unique_pretokens": int(eavg("cache_new_unique")), "pretokens": int(eavg("n_pretokens")),
            "roofline": {"bound": "hbm", "kernel": "whole encode pipeline (flags, lookup, bpe, fused scan+emit); slowest stage: " + dom,
                         "achieved": round(alg / 1e9 / (ems_dev / 1e3), 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg / 1e9 / (ems_dev / 1e3) / peak, 4), "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_definition": "N text bytes read once + 2 B per token written"},
            "gpu_launches": int(e_launches), "clocks": eclocks,
        }
        if not args.no_e2e:
            hbuf = _lib.PinnedBuffer(n_loc)
            torch.from_numpy(hbuf.array).copy_(etext[lo:hi])
            hout = _lib.PinnedBuffer(2 * (tokens_local + 1024))
            torch.cuda.synchronize()
            out_view = hout.array.view(np.uint16)

            def e2e_encode():
                ctx.check(L.bpe_tok_cache_reset(h))
                n_out = C.c_uint64(0)
                ctx.check(L.bpe_encode(h, _lib.ptr(hbuf.array), n_loc, _lib.DTYPE_U16, _lib.ptr(out_view), out_view.size, C.byref(n_out), None))
                return n_out.value

            e2e_encode()
            _, ems2, n2 = timed(e2e_encode, args.steps)
            assert n2[-1] == tokens_local
            assert np.array_equal(out_view[:4096], out[:4096].cpu().numpy())
            enc["e2e"] = {"value": round(e_bytes / 1e6 / (ems2 / 1e3), 2), "unit": "MB/s", "h2d_bytes_per_step": int(n_loc),
                          "d2h_bytes_per_step": int(2 * tokens_local), "ms_per_step": round(ems2, 3),
                          "api": "bpe_encode (C ABI) with pinned host text in, pinned host uint16 ids out; wall clock"}
            hbuf.free(); hout.free()
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            n_s, dt, n_ids = cpu_encode_sample(train_res[0], train_res[1], "owt", 4322, args.ref_sample_bytes if not is_train and args.ref_sample_bytes else 32e6)
            enc["cpu_baseline"] = {"value": round(n_s / 1e6 / dt, 4), "unit": "MB/s", "cores": 1, "kind": "port",
                                   "sample": "oracle/bpe_oracle.c Tokenizer.encode port (1 thread like the reference) on the first %.1f MB of the same text "
                                             "with the same vocab: %.1f s, %d ids" % (n_s / 1e6, dt, n_ids), "host_cores_available": os.cpu_count()}
        if is_train:
            line["encode"] = enc
        else:
            line = {"metric": enc["metric"], "value": enc["value"], "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": enc["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
                    "data": "synthetic", "config": enc["config"], "stages_ms": enc["stages_ms"], "roofline": enc["roofline"],
                    "gpu_launches": enc["gpu_launches"], "clocks": enc["clocks"]}
            if "e2e" in enc:
                line["e2e"] = enc["e2e"]

    # ======================= CPU baseline (rank 0, N = 1 only) =======================
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if is_train:
            sample = args.ref_sample_bytes or (1 << 20)
            n_s, dt, _, mm = cpu_train_sample(shape, seed, sample, vocab_size)
            line["cpu_baseline"] = {"value": round(n_s / 1e6 / dt, 4), "unit": "MB/s", "cores": 1, "kind": "port",
                                    "sample": "oracle/bpe_oracle.c (C port of the reference's train_bpe, 1 thread like the reference) on the first "
                                              "%.2f MB of the same corpus with the full vocab %d (%d merges): %.1f s" % (n_s / 1e6, vocab_size, len(mm), dt),
                                    "host_cores_available": os.cpu_count()}
        elif "cpu_baseline" in enc:
            line["cpu_baseline"] = enc["cpu_baseline"]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
