#!/usr/bin/env bash
# Builds oracle/_ref/gglasso: the UNMODIFIED reference package (fabian-sp/GGLasso, pure Python + numba), copied from
# the reference tree where it lies, with the one shim this image needs: numba 0.65 rejects keyword arguments of
# np.arange inside @njit code, so the four `np.arange(start=a, stop=b)` calls of solver/ggl_helper.py (lines 58, 169,
# 199, 242) are rewritten positionally -- semantics unchanged (route 1).
#
# Test infrastructure only (parity checker, golden-vector generation, bench.py's CPU reference arm and the
# reference's own input generators).  oracle/_ref/ is git-ignored (no reference sources in the history) but travels
# to the GPU box with the gpurun snapshot.  Usage: oracle/make_ref.sh [reference root, default /root/reference]
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$REF/src/gglasso"
DST="$HERE/_ref/gglasso"
if [ ! -d "$SRC" ]; then
    if [ -d "$DST" ]; then echo "make_ref: $SRC absent, keeping the existing $DST"; exit 0; fi
    echo "make_ref: reference tree not found at $SRC and no prebuilt oracle/_ref" >&2; exit 1
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC" "$DST"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
python3 - "$DST/solver/ggl_helper.py" <<'PY'
import re, sys
f = sys.argv[1]
s = open(f).read()
s2, n = re.subn(r"np\.arange\(start\s*=\s*([^,]+?)\s*,\s*stop\s*=\s*([^)]+?)\)", r"np.arange(\1, \2)", s)
assert n == 4, n
open(f, "w").write(s2)
PY
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$HERE/_ref/REF_COMMIT"
echo "make_ref: wrote $DST ($(find "$DST" -name '*.py' | wc -l) files, 4 np.arange calls rewritten)"
