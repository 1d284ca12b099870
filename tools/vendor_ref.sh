#!/usr/bin/env bash
# Vendor the UNMODIFIED reference into baseline/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
#
# The reference is pure Python with no setup.py / pyproject.toml, so `pip install` does not apply; its importable
# surface is the `models` and `utils` packages plus the two evaluation scripts.  They are copied byte for byte — nothing
# under baseline/_ref is ever edited, and nothing on the product path imports it.  Users of the copy:
#   * tests/test_gpu_dropin.py — runs the reference's own load_model / compute_auroc / evaluate_video.evaluate against
#     the drop-in classes (and against the reference classes on the CPU as the ground truth);
#   * bench.py --impl reference / cpu_baseline — times the reference's own CPU path (`cpu_baseline.kind = "reference"`).
# Run in the build container (needs /root/reference); __graft_entry__.build() calls it when the reference is present.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$REF/models" ]; then
  echo "vendor_ref: $REF not found (nothing to do; baseline/_ref must already be in place)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$REF/models" "$REF/utils" "$DST/"
cp "$REF/evaluate.py" "$REF/evaluate_video.py" "$REF/LICENSE" "$DST/"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
( cd "$REF" && sha256sum models/*.py utils/*.py evaluate.py evaluate_video.py ) > "$DST/SHA256SUMS"
echo "vendored $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
