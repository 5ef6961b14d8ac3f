#!/bin/bash
# tools/build_variant.sh <name> "<extra nvcc flags>": builds the library with the flags and copies it to tools/bin/lib<name>.so
set -e
cd "$(dirname "$0")/.."
BPE_EXTRA_NVCC_FLAGS="$2" python transformer-lm_b200/_build.py -f > /dev/null
cp transformer-lm_b200/libbpe_sm100.so tools/bin/lib$1.so
echo built tools/bin/lib$1.so "$2"
