#!/bin/sh
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with gpurun) so that
# tests/test_gpu_install.py can run it on the B200 frontend through track_analyser_b200.install().  The reference's own test
# files are placed next to it (baseline/_ref/_reference_tests, also git-ignored) for the same purpose.  Nothing from the
# reference enters the repository's history.  /root/reference is read-only, hence the copy under /tmp.
set -e
REPO=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
cp -r /root/reference "$TMP/ref"
cd "$TMP"   # pip --target is taken relative to the cwd
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$REPO/baseline/_ref" "$TMP/ref"
rm -rf "$REPO/baseline/_ref/_reference_tests"
cp -r /root/reference/tests "$REPO/baseline/_ref/_reference_tests"
rm -rf "$TMP"
echo "installed: $(ls "$REPO/baseline/_ref")"
