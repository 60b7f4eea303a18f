#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Installs the UNMODIFIED reference (mpasha3/trips-py, mounted read-only at /root/reference in the
# build container) into the git-ignored directory oracle/_ref/ with the one offline install the task allows:
#   pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>
# (--no-deps: pylops / astra-toolbox / h5py / python-resize-image are not installable offline; oracle/ref_loader.py
# registers stand-ins for them.  The copy under /tmp is needed because setuptools writes build/ into the source tree and
# /root/reference is read-only.)  oracle/_ref/ is listed in .gitignore but not in .gpurunignore, so it travels to the
# GPU box, where /root/reference does not exist: bench.py --impl reference and the drop-in tests import it from there.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${TRIPS_REFERENCE_SRC:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/trips" ]; then
  if [ -d "$DST/trips" ]; then echo "make_ref: $SRC absent, keeping existing $DST"; exit 0; fi
  echo "make_ref: $SRC not found and no $DST" >&2; exit 1
fi
TMP="$(mktemp -d /tmp/trips_ref_src.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$SRC"/. "$TMP"/
chmod -R u+w "$TMP"
rm -rf "$DST"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$TMP" \
  || { echo "make_ref: pip install failed, copying the package directory instead" >&2; mkdir -p "$DST"; cp -r "$SRC/trips" "$DST/trips"; }
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "make_ref: installed $(find "$DST/trips" -name '*.py' | wc -l) reference modules into $DST"
