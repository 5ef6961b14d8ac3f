#!/bin/bash
# Round-1 capture C: streaming kernels only (the merge loop is captured separately), 256 MB slices.
set -u
ARGS="--bytes 2.56e8 --encode-bytes 2.56e8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
TAG=${1:-r1c}
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k 'regex:k_count_pretokens|k_enc_lookup|k_enc_bpe|k_enc_emit|k_enc_ntok|k_pretok_flags|k_starts_to_offsets|k_scan_apply|k_build_words|k_special' --launch-skip 12 -c 24 \
    -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full_$TAG.log
