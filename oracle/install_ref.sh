#!/bin/sh
# Install the UNMODIFIED reference (mmuckley/torchkbnufft) into oracle/_ref/ (git-ignored, travels to the GPU box
# with the gpurun snapshot) so that tests and bench.py can time / run the stock package beside the engine:
#   oracle/_ref/torchkbnufft/      the package, pip-installed from a scratch copy of /root/reference
#   oracle/_ref/reference_tests/   the reference's own test files + golden pickles, run against the engine by
#                                  tests/test_reference_suite.py
# Nothing under oracle/_ref/ is committed; the product never imports it.  Usage: sh oracle/install_ref.sh [/root/reference]
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DEST="$HERE/_ref"
[ -d "$SRC/torchkbnufft" ] || { echo "reference checkout not found at $SRC" >&2; exit 1; }
TMP="$(mktemp -d /tmp/tkbn_ref.XXXXXX)"
cp -r "$SRC/." "$TMP/"          # the reference tree is read-only; the build writes _version.py into its source
rm -rf "$DEST"
mkdir -p "$DEST"
SETUPTOOLS_SCM_PRETEND_VERSION=1.4.0 python -m pip install --quiet --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$DEST" "$TMP" || {
  # the build back end needs setuptools-scm; without it fall back to a plain copy of the pure-Python package
  echo "pip install failed; copying the pure-Python package instead" >&2
  cp -r "$SRC/torchkbnufft" "$DEST/torchkbnufft"
}
mkdir -p "$DEST/reference_tests"
cp -r "$SRC/tests" "$DEST/reference_tests/tests"
rm -rf "$TMP"
echo "reference installed under $DEST"
