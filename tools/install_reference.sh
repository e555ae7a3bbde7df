#!/bin/bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)
# so that `bench.py --impl reference` can time the reference's own code on the box's host cores.  Build container only.
# The reference tree is read-only: pip builds from a scratch copy.  PYPOWER and gym are not installable here: the
# reference runs on oracle/shims (restated PYPOWER 5.1.4 slice + gym.spaces), as it does for the golden fixtures.
# The 75 MB of shipped CSV chronics are dropped from the install: the bench plays synthetic chronics, written in the
# reference's on-disk format at run time (oracle/ref_folder.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${PYPOWNET_REFERENCE:-/root/reference}
rm -rf /tmp/pypownet_refcopy "$ROOT/baseline/_ref"
cp -r "$REF" /tmp/pypownet_refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" /tmp/pypownet_refcopy
rm -rf "$ROOT/baseline/_ref/parameters" "$ROOT/baseline/_ref/bin"
du -sh "$ROOT/baseline/_ref"
