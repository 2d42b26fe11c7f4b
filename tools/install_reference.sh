#!/bin/bash
# Installs the UNMODIFIED reference (julbean/describealign) into baseline/_ref with pip, offline.
#
# The reference's checkout does not build as it lies: setuptools' automatic discovery stops at
# "Multiple top-level packages discovered in a flat-layout: ['Package', 'readme_media']" (two
# directories of installer scripts and README images).  /root/reference is read-only anyway, so
# the install goes through a copy under /tmp from which those two non-code directories are removed;
# describealign.py itself is byte-identical to the reference's (checked below).
# baseline/_ref is git-ignored but not gpurun-ignored: it travels to the GPU box, where
# `bench.py --impl reference` loads it (oracle/ref_loader.py) and times it.
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
[ -f "$REF/describealign.py" ] || { echo "no reference at $REF"; exit 0; }
TMP=$(mktemp -d /tmp/describealign_ref.XXXXXX)
cp -r "$REF/." "$TMP/"
chmod -R u+w "$TMP"
rm -rf "$TMP/Package" "$TMP/readme_media"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$ROOT/baseline/_ref" "$TMP"
cmp "$ROOT/baseline/_ref/describealign.py" "$REF/describealign.py"
rm -rf "$TMP"
echo "installed: $ROOT/baseline/_ref/describealign.py"
