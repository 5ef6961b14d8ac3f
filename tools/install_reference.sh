#!/bin/bash
# Installs the UNMODIFIED reference (gashon/transformer-lm, /root/reference) into baseline/_ref so that bench.py can time the
# reference's own Python train_bpe / Tokenizer.encode beside the GPU (bench.py --impl reference, cpu_baseline).
#
# The reference's setup.py:27 reads transformer/VERSION, a file the repository does not contain, so `pip install /root/reference`
# fails while generating metadata.  The install therefore runs from a copy under /tmp with that one data file added
# (no Python source is touched; the installed modules are byte-identical to /root/reference/models/**).
# baseline/_ref is git-ignored (reference sources never enter this repository's history) and travels with gpurun.
set -e
cd "$(dirname "$0")/.."
REF=${1:-/root/reference}
[ -d "$REF/models/tokenizer" ] || { echo "no reference at $REF"; exit 1; }
TMP=$(mktemp -d /tmp/refcopy.XXXXXX)
cp -r "$REF"/. "$TMP"/
mkdir -p "$TMP/transformer" && echo "0.0.0" > "$TMP/transformer/VERSION"
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref "$TMP" > /tmp/install_reference.log 2>&1 || { tail -20 /tmp/install_reference.log; exit 1; }
rm -rf "$TMP"
for f in train.py tokenizer.py vocab.py; do cmp "$REF/models/tokenizer/$f" "baseline/_ref/models/tokenizer/$f"; done
echo "installed baseline/_ref (models/tokenizer identical to $REF)"
