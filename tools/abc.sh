for c in $1; do echo -n "count ctas/SM=$c: "; BPE_COUNT_CTAS=$c BPE_LIB_PATH=$PWD/tools/bin/$2 timeout 200 python tools/prof_merge.py 1.1e10 2>&1 | head -1 | cut -c1-110; done
