#!/bin/bash
# A/B builds of K1 / K3 for profiles/sweep_k1.py and profiles/time_inference.py (run here, then `gpurun -- python profiles/sweep_k1.py`):
#   epi   KT_EPI_BATCH      y targets of all rows on the fast path, one fix-up branch per thread, label fetched by a select
#   empty KT_EMPTY_WARP     warps no table reaches: x targets per column, y targets per row (4 lanes + shuffles), no per-anchor matching state
#   ax    KT_AEXACT         anchors per cell a compile-time 9 in the common instantiation
#   k3pf  K3_PREFETCH_ROWS  K3 requests a candidate's regression row (prefetch.global.L2) when its score passes the threshold
# Every variant must print the same checksums as librn_b200.so; remove the .so files afterwards (they travel with gpurun).
set -e
cd "$(dirname "$0")/.."
python retinanet-for-table-detection_b200/build.py --variant epi KT_EPI_BATCH
python retinanet-for-table-detection_b200/build.py --variant empty KT_EMPTY_WARP
python retinanet-for-table-detection_b200/build.py --variant ax KT_AEXACT
python retinanet-for-table-detection_b200/build.py --variant all KT_EPI_BATCH KT_EMPTY_WARP KT_AEXACT
python retinanet-for-table-detection_b200/build.py --variant k3pf K3_PREFETCH_ROWS
