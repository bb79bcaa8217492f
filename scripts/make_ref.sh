#!/bin/bash
# Stage the UNMODIFIED reference (pure Python) under oracle/_ref/ so that it travels to the GPU box with the
# gpurun snapshot (oracle/_ref/ is git-ignored, NOT gpurun-ignored): the box has no /root/reference.
#   oracle/_ref/optiland/   the package, byte for byte, minus two catalogue folders of the refractive-index
#                           database (data-nk/organic, data-nk/other: 15 MB of materials no sample system or
#                           reference test on this path uses)
#   oracle/_ref/tests/      the reference's own test files (the sweep of scripts/ref_sweep.sh / the -m gpu sweep)
# Nothing from here enters git history; `oracle/ref_import.py` picks it up when /root/reference is absent.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${OPTILAND_REFERENCE_ROOT:-/root/reference}"
DST="$REPO/oracle/_ref"
if [ ! -d "$SRC/optiland" ]; then
  echo "[make_ref] $SRC/optiland not found; keeping whatever is in $DST" >&2
  exit 0
fi
mkdir -p "$DST"
if command -v rsync >/dev/null 2>&1; then
  rsync -a --delete --exclude '__pycache__' --exclude 'database/data-nk/organic' --exclude 'database/data-nk/other' \
    "$SRC/optiland" "$DST/"
  rsync -a --delete --exclude '__pycache__' "$SRC/tests" "$DST/"
else
  rm -rf "$DST/optiland" "$DST/tests"
  cp -r "$SRC/optiland" "$DST/optiland"
  cp -r "$SRC/tests" "$DST/tests"
  rm -rf "$DST/optiland/database/data-nk/organic" "$DST/optiland/database/data-nk/other"
  find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
fi
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$DST/REFERENCE_COMMIT"
du -sh "$DST" | sed 's/^/[make_ref] staged /'
