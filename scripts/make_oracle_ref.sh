#!/bin/bash
# Builds oracle/_ref: a git-ignored copy of the reference's own Python package (the hot path's files only), so that the
# UNMODIFIED reference travels to the GPU box with gpurun and `bench.py --impl reference` / `cpu_baseline` can time the
# real thing on the box's host cores (kind "reference") instead of the oracle port.  The reference is pure Python: there
# is nothing to compile, "building" is a copy of the package directory.  Never committed (see .gitignore); the
# GPU box has no /root/reference, so bench.py falls back to the oracle port when oracle/_ref is missing.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$HERE/oracle/_ref"
if [ ! -d "$REF/speechcatcher" ]; then echo "no reference checkout at $REF: oracle/_ref not built" >&2; exit 0; fi
rm -rf "$OUT"
mkdir -p "$OUT/speechcatcher"
# the streaming decode path: facade, beam search, model (frontend / encoder / decoder / attention / layers / ctc / loader)
cp "$REF/speechcatcher/__init__.py" "$REF/speechcatcher/speech2text_streaming.py" "$OUT/speechcatcher/"
cp -r "$REF/speechcatcher/beam_search" "$REF/speechcatcher/model" "$OUT/speechcatcher/"
find "$OUT" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$OUT/REFERENCE_COMMIT"
echo "oracle/_ref: $(find "$OUT" -name '*.py' | wc -l) files copied from $REF"
