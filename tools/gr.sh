#!/bin/bash
# Build first; only go to the GPU when the library links. usage: tools/gr.sh <timeout> '<command>'
set -e
cd "$(dirname "$0")/.."
if ! make -C vettore_b200/csrc -j8 > /tmp/vb_build.log 2>&1; then
  grep -E "error" /tmp/vb_build.log | head -20
  echo "BUILD FAILED - not calling gpurun"
  exit 1
fi
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
