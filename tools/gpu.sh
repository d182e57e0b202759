#!/bin/bash
# build everything that travels to the GPU box (CUDA library, oracle checkers, overlay demo), then run a command there
#   tools/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
exec /usr/local/graft/bin/gpurun "$@"
