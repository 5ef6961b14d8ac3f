#!/bin/bash
# Round-2 evidence run (one B200): plain bench first (must exit 0 without ncu), then the ncu launch list of the same command and
# ncu --set full captures of the dominant kernels at the bench configuration.  Outputs under gpurun_out/ (summaries are copied to profiles/).
set -x
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-slices --no-tiny --no-files"
python bench.py $ARGS > gpurun_out/plain_r2.json 2> gpurun_out/plain_r2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py $ARGS > gpurun_out/ncu_list_r2.log 2>&1
ncu --set full --import-source on --clock-control none -k "regex:k_merge_loop|k_pretok_flags|k_build_words|k_init_pair_counts|k_csr_fill" -c 5 -o gpurun_out/prof_r2_train python tools/time_train.py owt 4321 1.1e10 32000 1 > gpurun_out/ncu_full_r2_train.log 2>&1
ncu --set full --import-source on --clock-control none -k "regex:k_count_pretokens" --launch-skip 30 -c 2 -o gpurun_out/prof_r2_count python tools/time_train.py owt 4321 1.1e10 32000 1 > gpurun_out/ncu_full_r2_count.log 2>&1
ncu --set full --import-source on --clock-control none -k "regex:k_enc_lookup|k_enc_scan_emit|k_enc_bpe_short|k_pretok_flags|k_special" --launch-skip 120 -c 8 -o gpurun_out/prof_r2_encode python tools/enc_probe.py 1e9 1e10 > gpurun_out/ncu_full_r2_encode.log 2>&1
ls -la gpurun_out/prof_r2_*.ncu-rep
