#!/bin/bash
for b in 65536 16384 4096 1024 256; do echo "== MPTG_ORDER_BINS=$b"; MPTG_ORDER_BINS=$b timeout 300 python tools/knn_small_wave.py 2>&1 | grep -E "Q= *(8192|16384|65536)"; done | tee gpurun_out/order_bins.txt
