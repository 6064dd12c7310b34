#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, shipped to the GPU box by gpurun):
#   * the `vitef` package with pip (offline, no deps: every dependency the hot path needs is already in the image);
#   * the `apps/` tree next to it (a namespace package that pyproject.toml does not install: apps/vit/utils.py holds
#     freeze_model, apps/vit/analysis.py holds distance).
# Run in the build container only (/root/reference does not exist on the GPU box). Recorded in DESIGN.md section 6.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
ref="${1:-/root/reference}"
rm -rf "$here/_ref"
python -m pip install --quiet --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$here/_ref" "$ref"
cp -r "$ref/apps" "$here/_ref/apps"
find "$here/_ref" -name "__pycache__" -type d -prune -exec rm -rf {} +
echo "reference installed under $here/_ref: $(ls "$here/_ref" | tr '\n' ' ')"
