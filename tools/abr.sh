for b in $2; do for r in $1; do echo -n "$b minrec=$r: "; BPE_MERGE_MINREC=$r BPE_LIB_PATH=$PWD/tools/bin/$3 timeout 200 python tools/prof_merge.py $b 2>&1 | head -1 | cut -c80-200; done; done
