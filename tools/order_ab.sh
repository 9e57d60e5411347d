#!/bin/bash
# chunk order x carve-out hint of the index kernel: tools/order_ab.py under GKM_IDX_CARVEOUT = 1 (default) and 0
for c in 1 0; do echo "== GKM_IDX_CARVEOUT=$c"; GKM_IDX_CARVEOUT=$c python tools/order_ab.py 2>&1 | grep -v "^DEBUG\|^INFO\|^WARN"; done
