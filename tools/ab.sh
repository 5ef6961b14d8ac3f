# A/B of prebuilt libraries in one box: tools/ab.sh "<libs>" <bytes> [repeats]
for r in $(seq 1 ${3:-2}); do for l in $1; do echo -n "$l: "; BPE_LIB_PATH=$PWD/tools/bin/$l timeout 200 python tools/prof_merge.py ${2:-1.1e10} 2>&1 | head -1; done; done
