for g in $1; do echo -n "G=$g: "; BPE_MERGE_G=$g BPE_LIB_PATH=$PWD/tools/bin/$2 timeout 200 python tools/prof_merge.py ${3:-1.1e10} 2>&1 | head -1 | cut -c100-200; done
