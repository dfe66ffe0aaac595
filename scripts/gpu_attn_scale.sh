#!/bin/bash
for B in 512 1024 2048 4096 8192 16384; do timeout 120 python scripts/time_attention.py $B 52 2 32 2>&1 | tail -1 | cut -c1-60; done
echo "grid 148:"; B4CP_ATTN_GRID=148 timeout 120 python scripts/time_attention.py 4096 52 2 32 2>&1 | tail -1 | cut -c1-60
B4CP_ATTN_GRID=148 timeout 120 python scripts/time_attention.py 16384 52 2 32 2>&1 | tail -1 | cut -c1-60
