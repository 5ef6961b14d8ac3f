"""Checker-backed stand-in for transformer_lm_b200.sharded.DeviceCounter, used by the CPU (gloo) tests of the
multi-rank host logic.  It follows the contract of bpe_count_add_shard (include/bpe_sm100.h): only pretokens that
START in the owned range are counted, UTF-8 errors are reported only when they start in the owned range, start
positions in the last 16 bytes of a shard cut mid-file are not trusted, a carriage return refuses the shard."""
from __future__ import annotations

import numpy as np
import torch

from oracle import oracle


class OracleCounter:
    def __init__(self):
        self.device = torch.device("cpu")
        self.table: dict[bytes, int] = {}
        self.adds = 0

    def restart(self):
        self.table = {}

    def add(self, data, own_begin, own_end, at_file_start, at_file_end, device_ptr=None, n_bytes=None):
        data = bytes(data)
        self.adds += 1
        n = len(data)
        bad = oracle.utf8_validate(data[own_begin:])
        if bad >= 0 and own_begin + bad < own_end:
            return ("utf8", own_begin + bad)
        if b"\r" in data:
            return ("newline", 0)
        if oracle.utf8_validate(data) >= 0:
            return None                          # an error some other rank owns: the job fails there
        starts = oracle.pretokenize(data) + [n]
        trust_end = n if at_file_end else max(n - 16, 0)
        for i in range(len(starts) - 1):
            s, e = starts[i], starts[i + 1]
            if own_begin <= s < own_end:
                if e > trust_end:
                    return ("halo", 0)
                w = data[s:e]
                self.table[w] = self.table.get(w, 0) + 1
        return None

    def export(self):
        words = list(self.table.keys())
        blob = np.frombuffer(b"".join(words) or b"\0", dtype=np.uint8).copy()
        offs = np.zeros(len(words) + 1, dtype=np.int64)
        if words:
            np.cumsum([len(w) for w in words], out=offs[1:])
        counts = np.array([self.table[w] for w in words], dtype=np.int64)
        nb = int(offs[-1])
        return torch.from_numpy(blob)[:nb], torch.from_numpy(offs), torch.from_numpy(counts)

    def import_(self, blob, offs, counts):
        raw = bytes(blob.numpy().tobytes())
        o = offs.tolist()
        for i, c in enumerate(counts.tolist()):
            w = raw[o[i]:o[i + 1]]
            self.table[w] = self.table.get(w, 0) + c

    def pair_table(self, special_tokens):
        sp = {s.encode("utf-8") for s in special_tokens}
        dense = np.zeros(65536, dtype=np.int64)
        for w, c in self.table.items():
            if w in sp:
                continue
            for a, b in zip(w, w[1:]):
                dense[a * 256 + b] += c
        return torch.from_numpy(dense)


class DeferredCheckCounter(OracleCounter):
    """Like DeviceCounter: takes the all-reduced pair table and checks it later (DeviceCounter.finish does that against the
    table the merge phase builds)."""

    def __init__(self):
        super().__init__()
        self.expected = None

    def expect_pair_table(self, dense):
        self.expected = dense
