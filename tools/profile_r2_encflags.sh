#!/bin/bash
# Round-2 evidence, addendum: ncu --set full capture of the encoder's pretokenise stage (k_pretok_flags<1> and the special-token
# passes run ONCE per encode call over the whole text, so the per-batch capture window of tools/profile_r2.sh never saw them).
set -x
timeout 170 ncu --set full --import-source on --clock-control none -k "regex:k_pretok_flags|k_special_candidates|k_special_resolve|k_popc_words16" -c 8 \
    -o gpurun_out/prof_r2_encflags python tools/enc_probe.py 1e9 1e10 > gpurun_out/ncu_full_r2_encflags.log 2>&1
ls -la gpurun_out/prof_r2_encflags.ncu-rep
