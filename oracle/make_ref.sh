#!/bin/bash
# TEST INFRASTRUCTURE.  Snapshots the UNMODIFIED reference files of the hot path from /root/reference into the
# git-ignored oracle/_ref/ so that the GPU box (where /root/reference does not exist) can run the verbatim reference
# as the checker (tests/, smoke()) and as bench.py's `--impl reference` / cpu_baseline leg (kind "reference").
# Nothing under oracle/_ref/ is ever committed (.gitignore) and the product package never imports it.
#   usage: bash oracle/make_ref.sh [reference_root]      (default /root/reference)
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$REF/baseline_code" ]; then
  echo "make_ref: $REF/baseline_code not found (GPU box?) - keeping whatever is in $DST"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/baseline_code/models" "$DST/baseline_code/sampling" "$DST/conf/models"
for f in config.py d_model.py flow_model.py models/__init__.py models/bsrnn.py models/bsrnn_flowse.py models/odes.py \
         sampling/__init__.py sampling/odesolvers.py; do
  cp "$REF/baseline_code/$f" "$DST/baseline_code/$f"
done
cp "$REF"/conf/models/*.yaml "$DST/conf/models/"
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/COMMIT"
( cd "$DST" && find . -type f ! -name SHA256SUMS | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "make_ref: $(find "$DST" -name '*.py' | wc -l) reference files -> $DST"
