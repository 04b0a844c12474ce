#!/bin/bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)
# so that bench.py can time the reference's own predict() there.  Test / baseline infrastructure only.
#
# The reference's setup.py lists `pytest-runner` and `flake8` as setup_requires (setup.py:5); neither is in the offline
# wheelhouse and neither is needed to install a pure-Python package, so two empty dist-info stubs on PYTHONPATH satisfy
# setuptools for the duration of this command.  /root/reference is read-only: the build runs in a copy under /tmp.
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${OSML_REF_SRC:-/root/reference}"
[ -f "$SRC/setup.py" ] || { echo "no reference at $SRC"; exit 0; }
[ -f "$REPO/baseline/_ref/osml/predict.py" ] && exit 0
TMP="$(mktemp -d)"
cp -r "$SRC/." "$TMP/src"
mkdir -p "$TMP/stubs/pytest_runner-6.0.1.dist-info" "$TMP/stubs/flake8-7.0.0.dist-info"
printf 'Metadata-Version: 2.1\nName: pytest-runner\nVersion: 6.0.1\n' > "$TMP/stubs/pytest_runner-6.0.1.dist-info/METADATA"
printf 'Metadata-Version: 2.1\nName: flake8\nVersion: 7.0.0\n' > "$TMP/stubs/flake8-7.0.0.dist-info/METADATA"
touch "$TMP/stubs/pytest_runner-6.0.1.dist-info/RECORD" "$TMP/stubs/flake8-7.0.0.dist-info/RECORD"
cd "$TMP/src"
PYTHONPATH="$TMP/stubs" python -m pip install -q --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$REPO/baseline/_ref" "$TMP/src"
rm -rf "$TMP"
