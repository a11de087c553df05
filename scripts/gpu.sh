#!/bin/bash
# Rebuild every native artefact, then ship the tree to the GPU box:  scripts/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m deformationpyramid_b200.build
python oracle/build.py > /dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
