#!/usr/bin/env bash
# Runs the reference's own pytest files (copied into the git-ignored baseline/_ref/tests by baseline/install_ref.sh)
# against the qibo_b200 backend on a GPU box.  Usage: scripts/run_reference_tests.sh [pytest args / file names]
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH="$PWD/baseline/_ref:$PWD${PYTHONPATH:+:$PYTHONPATH}"
export QIBO_LOG_LEVEL=3
FILES=("$@")
if [ ${#FILES[@]} -eq 0 ]; then
  FILES=(baseline/_ref/tests)
fi
cd baseline/_ref
python -m pytest -p no:cacheprovider -o addopts="" -q -k "qibo_b200 or not numpy" --tb=line "${FILES[@]/#baseline\/_ref\//}"
