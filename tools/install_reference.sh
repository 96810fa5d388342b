#!/bin/bash
# Installs the UNMODIFIED reference (/root/reference, read-only) into baseline/_ref (git-ignored, travels with gpurun):
#   pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy under /tmp>
# (--no-deps: setup.py pins numpy==1.16.4 / dominate / dill / future, absent offline and unused on the generator path;
#  a copy because the build writes into the source tree).
# The reference's setup.py uses find_packages(), which skips ctu/models/pix2pixHD_networks/ -- that directory has no
# __init__.py (a namespace package when run from a checkout, as its README does). The wheel therefore lacks it; the
# second step completes the install with exactly those .py files, unmodified.
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
rm -rf "$TMP/ref/datasets"
rm -rf "$REPO/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$REPO/baseline/_ref" "$TMP/ref" 2>&1 | tail -2
(cd "$REF" && find ctu -name '*.py' | while read -r f; do
    if [ ! -f "$REPO/baseline/_ref/$f" ]; then
        mkdir -p "$REPO/baseline/_ref/$(dirname "$f")"
        cp "$f" "$REPO/baseline/_ref/$f"
        echo "completed install with $f"
    fi
done)
rm -rf "$TMP"
find "$REPO/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
(cd "$REF" && find ctu -name '*.py' -exec sha256sum {} + | sort -k2) > "$TMP.ref.sha"
(cd "$REPO/baseline/_ref" && find ctu -name '*.py' -exec sha256sum {} + | sort -k2) > "$TMP.inst.sha"
if cmp -s "$TMP.ref.sha" "$TMP.inst.sha"; then echo "baseline/_ref/ctu is byte-identical to $REF/ctu"; else echo "MISMATCH"; exit 1; fi
rm -f "$TMP.ref.sha" "$TMP.inst.sha"
