#!/bin/bash
# Runs the UNMODIFIED reference's own test suite (/root/reference/tests, 26 integration tests) on top of the oracle
# shims (oracle/shims: gym.spaces + restated PYPOWER) and records the outcome in tests/golden/REFERENCE_TESTS.txt.
# Build container only (/root/reference does not travel).  The tests use relative paths and write files into the
# working directory, so they run from a scratch copy of the reference tree's layout.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${PYPOWNET_REFERENCE:-/root/reference}
W=/tmp/pypownet_reference_tests
rm -rf $W && mkdir -p $W
cp -r $REF/tests $W/tests
ln -s $REF/pypownet $W/pypownet
ln -s $REF/parameters $W/parameters
cd $W
OUT=$ROOT/tests/golden/REFERENCE_TESTS.txt
{
  echo "The reference's own suite (pytest $REF/tests, reference commit $(cd $REF && git rev-parse --short HEAD 2>/dev/null || echo 8839d90))"
  echo "run UNMODIFIED on oracle/shims (gym.spaces + restated PYPOWER 5.1.4), python $(python -c 'import sys; print(sys.version.split()[0])'), numpy $(python -c 'import numpy; print(numpy.__version__)'), scipy $(python -c 'import scipy; print(scipy.__version__)')"
  echo "command: PYTHONPATH=oracle/shims:<scratch copy> python -m pytest tests -q -p no:cacheprovider   (tools/run_reference_tests.sh)"
  echo
  PYTHONPATH=$ROOT/oracle/shims:$W python -m pytest tests -v -p no:cacheprovider -W ignore 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed" | sed "s#$W/##"
} > $OUT
tail -3 $OUT
