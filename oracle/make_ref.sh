#!/bin/bash
# oracle/make_ref.sh — places the UNMODIFIED reference's hot-path modules (pure Python + NumPy; nothing to compile)
# under oracle/_ref/ so that they travel to the GPU box, where /root/reference does not exist.  oracle/_ref/ is
# git-ignored: reference sources are never committed.  Test infrastructure / CPU baseline only — see bench.py
# (`--impl reference`, cpu_baseline) — the product path never imports it.
set -e
SRC=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
DST=$HERE/_ref
[ -d "$SRC/layers" ] || { echo "make_ref: $SRC is not present; keeping $DST as it is"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST/layers"
for f in layer.py mlp.py conv.py attentions.py normalizations.py activations.py transformer.py __init__.py; do
  cp "$SRC/layers/$f" "$DST/layers/$f"
done
for f in optimizer.py loss.py train.py; do cp "$SRC/$f" "$DST/$f"; done
echo "make_ref: $(find "$DST" -name '*.py' | wc -l) reference modules under $DST"
