#!/bin/bash
for v in 0 1; do echo "spin=$v"; B4CP_ATTN_SPIN=$v timeout 120 python scripts/time_attention.py 4096 52 2 32 2>&1 | tail -1 | cut -c1-110; done
