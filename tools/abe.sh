for l in $1; do echo "== $l"; BPE_LIB_PATH=$PWD/tools/bin/$l timeout 300 python tools/enc_probe.py 1e9 ${2:-1e10} 2>&1 | grep "encode iter\|parity\|Error\|error" | cut -c1-330; done
