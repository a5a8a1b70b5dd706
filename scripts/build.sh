#!/bin/bash
# Rebuild libits_b200.so from the repo root; prints BUILD OK / BUILD FAILED.
cd "$(dirname "$0")/.." || exit 1
if python -c "import __graft_entry__ as g; g.build()" > /tmp/its_build.log 2>&1; then
  echo "BUILD OK $(date +%T)"
else
  grep -E "error" /tmp/its_build.log | head -20
  echo "BUILD FAILED"
  exit 1
fi
