#!/usr/bin/env python3
"""Benchmark of the byte-level BPE hot path on B200 (contract: see DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|encode] [--bytes B] [--vocab V]
    python bench.py --impl reference ...          # CPU port of the reference on a bounded sample

Default workload (N=1): BPE training on the 11 GB synthetic OWT-shaped corpus, vocab 32000,
special token <|endoftext|> (BASELINE.json configs[3]; the shape of perf/bpe/owt.py:4-8 in the reference).
One "step" = one complete train_bpe over the corpus (pretokenise, count, 31 743 merges).
`value` = corpus MB per second of a whole training run with the text already resident in HBM;
`e2e` = the same through the C-ABI call with the text in pinned HOST memory (H2D inside the timed region).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "encode", "train-tiny"])
    ap.add_argument("--bytes", type=float, default=None, help="corpus size in bytes (default: the BASELINE config)")
    ap.add_argument("--vocab", type=int, default=None)
    ap.add_argument("--ref-sample-bytes", type=float, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


WORKLOADS = {
    # name: (shape, seed, default bytes, default vocab, description)
    "train": ("owt", 4321, 11e9, 32000, "BPE train, synthetic OWT-shape corpus, vocab 32000, special <|endoftext|>"),
    "train-tiny": ("tinystories", 1234, 2 * 2**30, 10000, "BPE train, synthetic TinyStories-shape corpus, vocab 10000"),
    "encode": ("owt", 4322, 10e9, 32000, "bulk encode, synthetic OWT-shape text, 32K vocab trained on the OWT-shape train corpus"),
}


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference on a bounded sample, host cores only
# ----------------------------------------------------------------------------------------------
def cpu_train_sample(shape, seed, sample_bytes, vocab):
    import _bootstrap  # noqa: F401
    from oracle import oracle
    from transformer_lm_b200.synth import synth_host
    n = int(sample_bytes) // 4096 * 4096
    data = synth_host(shape, seed, n).tobytes()
    t0 = time.perf_counter()
    _, merges = oracle.train_bpe_on_bytes(data, vocab, SPECIALS)
    dt = time.perf_counter() - t0
    return n, dt, len(merges)


def cpu_encode_sample(shape, seed, sample_bytes, vocab, train_bytes):
    import _bootstrap  # noqa: F401
    from oracle import oracle
    from transformer_lm_b200.synth import synth_host
    tv, tm = oracle.train_bpe_on_bytes(synth_host("owt", 4321, int(train_bytes) // 4096 * 4096).tobytes(), vocab, SPECIALS)
    tok = oracle.OracleTokenizer(tv, tm, SPECIALS)
    n = int(sample_bytes) // 4096 * 4096
    data = synth_host(shape, seed, n).tobytes()
    t0 = time.perf_counter()
    ids = tok.encode_bytes(data)
    dt = time.perf_counter() - t0
    return n, dt, len(ids)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shape, seed, nbytes, vocab, desc = WORKLOADS[args.workload]
    vocab = args.vocab or vocab
    times = []
    if args.workload == "encode":
        sample = args.ref_sample_bytes or 4e6
        for i in range(args.warmup + args.steps):
            n, dt, _ = cpu_encode_sample(shape, seed, sample, min(vocab, 2000), 1 << 20)
            if i >= args.warmup:
                times.append(dt)
        metric = "bpe_encode_MBps"
        sample_desc = "first %.1f MB of the OWT-shape encode text, tokenizer = oracle-trained vocab %d on a 1 MiB slice" % (n / 1e6, min(vocab, 2000))
    else:
        sample = args.ref_sample_bytes or 1 << 20
        for i in range(args.warmup + args.steps):
            n, dt, nm = cpu_train_sample(shape, seed, sample, vocab)
            if i >= args.warmup:
                times.append(dt)
        metric = "bpe_train_MBps"
        sample_desc = "first %.2f MB of the corpus (same generator + seed), full vocab %d => %d merges" % (n / 1e6, vocab, nm)
    ms = 1e3 * sum(times) / len(times)
    value = n / 1e6 / (ms / 1e3)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": {"workload": desc, "sample_bytes": n},
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": 1, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (pure Python, single-threaded) cannot travel to the GPU box; this is its C port oracle/bpe_oracle.c, "
                "same algorithm (O(pairs) argmax scan per merge), 1 host thread like the reference",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import _bootstrap  # noqa: F401
    from transformer_lm_b200 import _lib
    from transformer_lm_b200.synth import synth_device
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
    L = _lib.lib()

    shape, seed, nbytes, vocab, desc = WORKLOADS[args.workload]
    nbytes = int(args.bytes or nbytes) // 4096 * 4096
    vocab = args.vocab or vocab

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "encode":
        from bench_encode import run_encode_bench
        return run_encode_bench(args, ctx, world, rank, local_rank, nbytes, vocab, desc, barrier, ClockSampler, measured_peak_gbs)

    if world > 1:
        from bench_multi import run_train_multi
        return run_train_multi(args, ctx, world, rank, local_rank, shape, seed, nbytes, vocab, desc, barrier, ClockSampler,
                               measured_peak_gbs)

    # ---- inputs resident in HBM ----
    text_dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    synth_device(shape, seed, nbytes, text_dev.data_ptr(), ctx=ctx)
    torch.cuda.synchronize()

    def step_dev():
        return train_bpe_on_bytes(None, vocab, SPECIALS, ctx=ctx, return_stats=True, device_ptr=text_dev.data_ptr(), n_bytes=nbytes)

    for _ in range(args.warmup):
        res = step_dev()
    launches0 = L.bpe_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    stats_acc = []
    for _ in range(args.steps):
        res = step_dev()
        stats_acc.append(res[2])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = L.bpe_launch_count() - launches0
    ms_per_step = 1e3 * wall / args.steps
    value = nbytes / 1e6 / (ms_per_step / 1e3)
    merges = res[1]

    def avg(key):
        return sum(s[key] for s in stats_acc) / len(stats_acc)

    stages = {k: round(avg(k), 3) for k in ("ms_h2d", "ms_pretok", "ms_count", "ms_build", "ms_merge", "ms_total")}
    st = stats_acc[-1]
    peak, peak_src = measured_peak_gbs()
    # dominant HBM-streaming kernel of the step: the pretoken count kernel (reads the text once: N bytes)
    # and the flags kernel (N read + N/8 written); the merge loop is latency-bound and reported separately.
    flags_bytes = nbytes + nbytes / 8
    count_bytes = nbytes + nbytes / 8
    dom = "count" if stages["ms_count"] >= stages["ms_pretok"] else "flags"
    dom_ms = stages["ms_count"] if dom == "count" else stages["ms_pretok"]
    dom_bytes = count_bytes if dom == "count" else flags_bytes
    achieved = dom_bytes / 1e9 / (dom_ms / 1e3) if dom_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_count_pretokens" if dom == "count" else "k_pretok_flags", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes,
                "other_kernels": {
                    "k_pretok_flags_GBps": round(flags_bytes / 1e9 / (stages["ms_pretok"] / 1e3), 1) if stages["ms_pretok"] else None,
                    "k_count_pretokens_GBps": round(count_bytes / 1e9 / (stages["ms_count"] / 1e3), 1) if stages["ms_count"] else None,
                    "k_merge_loop_us_per_merge": round(1e3 * stages["ms_merge"] / max(len(merges), 1), 3),
                    "note": "stage times are CUDA-event times on the library's stream (bpe_train_stats); the merge loop is one "
                            "persistent cooperative launch, latency-bound (2 grid syncs per merge), not an HBM stream"}}

    # ---- end to end: text in pinned host memory, H2D inside the timed region, merges read back ----
    e2e = None
    if not args.no_e2e:
        host = _lib.PinnedBuffer(nbytes)
        host_t = torch.from_numpy(host.array)
        host_t.copy_(text_dev)
        torch.cuda.synchronize()

        def step_host():
            return train_bpe_on_bytes(host.array, vocab, SPECIALS, ctx=ctx, return_stats=True)

        for _ in range(min(args.warmup, 3)):
            r2 = step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r2 = step_host()
        barrier()
        wall2 = time.perf_counter() - t0
        assert r2[1] == merges
        e2e = {"value": round(nbytes / 1e6 / (wall2 / args.steps), 2), "unit": "MB/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": 8 * len(merges), "ms_per_step": round(1e3 * wall2 / args.steps, 3)}
        del host_t
        host.free()

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sample = args.ref_sample_bytes or 1 << 20
        n_s, dt, nm = cpu_train_sample(shape, seed, sample, vocab)
        cpu_baseline = {"value": round(n_s / 1e6 / dt, 4), "unit": "MB/s", "cores": 1, "kind": "port",
                        "sample": "oracle/bpe_oracle.c (C port of the reference's train_bpe, 1 thread) on the first %.2f MB of the same "
                                  "corpus with the full vocab %d (%d merges): %.1f s" % (n_s / 1e6, vocab, nm, dt),
                        "host_cores_available": os.cpu_count()}

    line = {
        "metric": "bpe_train_MBps", "value": round(value, 2), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 3), "wall_s_per_train": round(ms_per_step / 1e3, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "corpus_bytes": nbytes, "vocab_size": vocab, "merges": len(merges), "shape": shape, "seed": seed,
                   "l2_policy": "inputs (%.1f GB) larger than L2 (126 MB)" % (nbytes / 1e9) if nbytes > L2_BYTES else "input smaller than L2",
                   "unicode_tables": L.bpe_unicode_table_source().decode(), "parallelism": "1 GPU"},
        "stages_ms": stages,
        "counts": {k: st[k] for k in ("n_pretokens", "n_unique", "n_symbols", "n_pairs_initial", "n_pairs_final", "log_records")},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
