for b in 1e9 1.1e10 2.56e8; do for l in $1; do echo -n "$b $l: "; BPE_LIB_PATH=$PWD/tools/bin/$l timeout 200 python tools/prof_merge.py $b 2>&1 | head -1 | cut -c80-200; done; done
