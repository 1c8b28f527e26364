#!/bin/bash
cd "$(dirname "$0")/.."
for t in 1024 512 256; do echo "=== GG_ATTN_TC_MIN_TK=$t"; GG_ATTN_TC_MIN_TK=$t python tools/bench_attn.py 2>&1 | grep -E "T256|T1024|T2048 "; done
