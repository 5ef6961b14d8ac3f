#!/bin/bash
# A/B of prebuilt library variants on ONE box (boxes of the pool differ by up to 12 %), optionally with per-variant environment: tools/ab_env.sh "lib1.so[:VAR=x[,VAR2=y]] lib2.so ..." "<command>" [repeats]
for r in $(seq 1 ${3:-1}); do for spec in $1; do l=${spec%%:*}; envs=""; [ "$spec" != "$l" ] && envs=$(echo "${spec#*:}" | tr ',' ' '); echo -n "$spec: "; env $envs BPE_LIB_PATH=$PWD/tools/bin/$l timeout 600 $2 2>&1 | tail -1; done; done
