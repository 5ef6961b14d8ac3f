#!/bin/bash
# Round-1 capture I: steady-state batches (third 256 MB batch of a 1 GB run) of k_count_pretokens and k_enc_lookup.
set -u
ARGS="--bytes 1.024e9 --encode-bytes 1.024e9 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_r1i.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r1i.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_count_pretokens --launch-skip 2 -c 1 -o gpurun_out/prof_r1i_count python bench.py $ARGS > gpurun_out/ncu_full_r1i_count.log 2>&1
echo "count capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_enc_lookup --launch-skip 2 -c 1 -o gpurun_out/prof_r1i_lookup python bench.py $ARGS > gpurun_out/ncu_full_r1i_lookup.log 2>&1
echo "lookup capture rc=$?"
