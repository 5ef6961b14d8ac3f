#!/bin/bash
# Round-1 capture J (end of round): default bench line, launch list + full capture of a 256 MB bench step, merge loop at 11 GB.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default_r1j.json 2> gpurun_out/bench_default_r1j.err || { echo "default bench failed"; tail -20 gpurun_out/bench_default_r1j.err; exit 1; }
tail -c 400 gpurun_out/bench_default_r1j.json
ARGS="--bytes 2.56e8 --encode-bytes 2.56e8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_r1j.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1j.csv python bench.py $ARGS > gpurun_out/ncu_list_r1j.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_merge_loop|k_count_pretokens|k_enc_lookup|k_enc_bpe|k_enc_scan_emit|k_pretok_flags' -c 16 \
    -o gpurun_out/prof_r1j python bench.py $ARGS > gpurun_out/ncu_full_r1j.log 2>&1
echo "full capture rc=$?"
ARGS2="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-encode"
python bench.py $ARGS2 > gpurun_out/plain_r1k.log 2>&1 || { echo "plain 11GB run failed"; exit 1; }
ncu --set full --clock-control none -k regex:k_merge_loop -c 1 -o gpurun_out/prof_r1k_merge_11GB python bench.py $ARGS2 > gpurun_out/ncu_full_r1k.log 2>&1
echo "merge capture rc=$?"
python bench.py --workload train-tiny --no-encode --no-cpu-baseline > gpurun_out/bench_tiny_r1j.json 2> gpurun_out/bench_tiny_r1j.err; echo "tiny rc=$?"
