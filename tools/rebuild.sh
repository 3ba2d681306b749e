#!/bin/bash
# Development aid: rebuild librobchar_b200.so in tree and refresh the source hash (any cwd).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
bash "$ROOT/code-robchar_b200/csrc/build.sh" 8 2>&1 | grep -v "^nvcc warning" | tail -3
cd "$ROOT" && python -c "import __graft_entry__ as g; open(g.HASH_FILE,'w').write(g._source_hash())"
