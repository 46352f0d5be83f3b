#!/bin/bash
# Rebuild librtkv.so in-tree; non-zero exit on any compile error.
set -e
cd "$(dirname "$0")/../realtime-kv-cache-compression_b200"
python build.py > /tmp/rtkv_build.log 2>&1 || { tail -30 /tmp/rtkv_build.log; exit 1; }
if grep -l " error" build/*.log 2>/dev/null; then grep -h " error" build/*.log | head; exit 1; fi
tail -1 /tmp/rtkv_build.log
