#!/bin/bash
# Round-1 profiling recipe (B200_PROFILING.md): plain run first, then the launch list and one --set full capture.
set -u
ARGS="--bytes 2.56e8 --encode-bytes 2.56e8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
TAG=${1:-r1b}
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
tail -c 3000 gpurun_out/plain_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_merge_loop|k_count_pretokens|k_enc_lookup|k_enc_bpe|k_enc_emit|k_enc_count|k_pretok_flags' -c 16 \
    -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -5 gpurun_out/ncu_full_$TAG.log
