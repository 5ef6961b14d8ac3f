#!/bin/bash
# Round-1 capture G (after the merge-loop / count-kernel work of the last session): default bench line, launch list of a
# 256 MB bench step, full capture of the streaming kernels at 256 MB and of the merge loop at 11 GB.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default_r1g.json 2> gpurun_out/bench_default_r1g.err || { echo "default bench failed"; tail -20 gpurun_out/bench_default_r1g.err; exit 1; }
tail -c 600 gpurun_out/bench_default_r1g.json
ARGS="--bytes 2.56e8 --encode-bytes 2.56e8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_r1g.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1g.csv python bench.py $ARGS > gpurun_out/ncu_list_r1g.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_merge_loop|k_count_pretokens|k_enc_lookup|k_enc_bpe|k_enc_emit|k_enc_ntok|k_pretok_flags' -c 16 \
    -o gpurun_out/prof_r1g python bench.py $ARGS > gpurun_out/ncu_full_r1g.log 2>&1
echo "full capture rc=$?"
ARGS2="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-encode"
python bench.py $ARGS2 > gpurun_out/plain_r1h.log 2>&1 || { echo "plain 11GB run failed"; exit 1; }
ncu --set full --clock-control none -k regex:k_merge_loop -c 1 -o gpurun_out/prof_r1h_merge_11GB python bench.py $ARGS2 > gpurun_out/ncu_full_r1h.log 2>&1
echo "merge capture rc=$?"
