#!/usr/bin/env bash
# Stage the UNMODIFIED reference (pure Python, no build system) where the GPU box can see it:
#   /root/reference/{src,tests,pyproject.toml}  ->  baseline/_ref/   (git-ignored, NOT gpurun-ignored)
# Used by bench.py (cpu_baseline / --impl reference, kind "reference") and by tests/test_gpu_reference_suite.py
# (the reference's own acceptance tests run on top of this package's core).  Nothing under baseline/_ref is
# tracked or edited; re-run this script to refresh it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${1:-/root/reference}"
DST="$HERE/baseline/_ref"
if [ ! -d "$REF/src" ]; then
  echo "stage_reference: $REF/src not found (nothing staged)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$REF/src" "$DST/src"
cp -r "$REF/tests" "$DST/tests"
[ -f "$REF/pyproject.toml" ] && cp "$REF/pyproject.toml" "$DST/pyproject.toml"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$REF" && find src tests -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "staged $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
