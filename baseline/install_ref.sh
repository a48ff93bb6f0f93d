#!/usr/bin/env bash
# Installs the UNMODIFIED reference (qibo 0.3.5, pure Python) into baseline/_ref (git-ignored).
# 1) tries the contract's offline pip install; 2) if the build backend (poetry-core) is missing from the
# wheelhouse, performs the equivalent of a pure-Python wheel install: package dir + dist-info metadata.
# Also drops a stub `openqasm3` module (reference imports it eagerly; it is absent from this image).
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
DST="$HERE/_ref"
[ -d "$REF/src/qibo" ] || { echo "no reference at $REF"; exit 0; }
mkdir -p "$DST"
if python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
      --target "$DST" "$REF" >/tmp/qibo_pip.log 2>&1; then
  echo "pip install ok"
else
  echo "pip install failed (see /tmp/qibo_pip.log); doing a plain pure-python install"
  rm -rf "$DST/qibo"
  cp -r "$REF/src/qibo" "$DST/qibo"
  mkdir -p "$DST/qibo-0.3.5.dist-info"
  printf 'Metadata-Version: 2.1\nName: qibo\nVersion: 0.3.5\n' > "$DST/qibo-0.3.5.dist-info/METADATA"
  printf 'manual\n' > "$DST/qibo-0.3.5.dist-info/INSTALLER"
fi
python -c "import openqasm3" 2>/dev/null || { mkdir -p "$DST/openqasm3"; printf 'parser = None\nast = None\n' > "$DST/openqasm3/__init__.py"; }
PYTHONPATH="$DST" python -c "import qibo; print('qibo', qibo.__version__, 'importable from', qibo.__file__)"
