#!/bin/bash
# Installs the UNMODIFIED reference (read-only at /root/reference) into baseline/_ref (git-ignored, travels to the GPU box):
#   python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the reference>
# From a copy under /tmp, because (1) the build writes into the source tree and /root/reference is read-only, and
# (2) setup.py's data_files names two weight files the checkout does not contain (data/weights/value_1.pt, policy_0.pt:
# `.MISSING_LARGE_BLOBS`), so the copy gets EMPTY placeholders for them (no source file is touched; the installed
# bokego/*.py are byte-identical to /root/reference/bokego/*.py -- checked below).  --no-deps because torch / numpy / pandas
# are already in the image and the wheelhouse holds no torch wheel.  The callers the north-star names that setup.py does not
# package (bin/selfplay.py, boke.py) and the second shipped weight file are added beside the package.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/bokego" ] || { echo "no reference at $REF"; exit 1; }
TMP="$(mktemp -d /tmp/bokego_ref.XXXXXX)"
cp -r "$REF/." "$TMP/"
chmod -R u+w "$TMP"
for f in value_1.pt policy_0.pt; do [ -e "$TMP/data/weights/$f" ] || : > "$TMP/data/weights/$f"; done
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP"
# data_files land under <target>/bokego/ next to the modules; drop the empty placeholders again
find "$ROOT/baseline/_ref" -name 'value_1.pt' -size 0 -delete
find "$ROOT/baseline/_ref" -name 'policy_0.pt' -size 0 -delete
mkdir -p "$ROOT/baseline/_ref/bin" "$ROOT/baseline/_ref/data/weights"
cp "$REF/bin/selfplay.py" "$ROOT/baseline/_ref/bin/selfplay.py"
cp "$REF/boke.py" "$ROOT/baseline/_ref/boke.py"
cp "$REF/data/weights/policy_17.pt" "$REF/data/weights/policy_19.pt" "$ROOT/baseline/_ref/data/weights/"
for f in __init__ go gtp mcts nnet; do cmp "$REF/bokego/$f.py" "$ROOT/baseline/_ref/bokego/$f.py"; done
rm -rf "$TMP"
echo "reference installed in baseline/_ref (unmodified)"
