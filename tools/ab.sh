#!/bin/bash
# A/B of prebuilt library variants on ONE box (boxes of the pool differ by up to 12 %): tools/ab.sh "<lib files in tools/bin>" "<command>" [repeats]
# e.g. tools/build_variant.sh k15 "-DMG_BATCH=15u"; gpurun -- 'tools/ab.sh "libk15.so libbase.so" "python tools/time_train.py owt 4321 1.1e10 32000" 2'
for r in $(seq 1 ${3:-1}); do for l in $1; do echo -n "$l: "; BPE_LIB_PATH=$PWD/tools/bin/$l timeout 600 $2 2>&1 | tail -1; done; done
