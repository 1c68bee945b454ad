#!/usr/bin/env bash
# Installs the UNMODIFIED reference (Sompote/sam3_lora) under baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to
# the GPU box).  Used by: bench.py's reference arms, the a9 bridge (sam3_bridge.py) in tests / bench, the golden generators.
#   tools/install_reference.sh [/path/to/reference]      (default /root/reference)
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
# the reference's setup.py writes into its source tree: build from a copy
cp -r "$SRC/." "$TMP/src"
chmod -R u+w "$TMP/src"
rm -f "$TMP/src"/*.zip
rm -rf "$DST"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$TMP/src"
# setup.py lists packages only: the root-level modules the CLI imports and the tokenizer vocabulary are copied next to them
cp "$SRC/lora_layers.py" "$SRC/train_sam3_lora_native.py" "$DST/"
mkdir -p "$DST/sam3/assets"
cp "$SRC/sam3/assets/bpe_simple_vocab_16e6.txt.gz" "$DST/sam3/assets/"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
du -sh "$DST"
