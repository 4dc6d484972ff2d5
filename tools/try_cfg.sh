#!/bin/bash
# usage: tools/try_cfg.sh GROUPS GROUP_WARPS PRODUCERS STORE "chunks..."   (dev helper)
set -e
cd /root/repo
F=matterport_maskrcnn_with_tensorflow_serving_b200/csrc/expand_ws.cu
sed -i "s/^constexpr int kGroups = [0-9]*;/constexpr int kGroups = $1;/; s/^constexpr int kGroupWarps = [0-9]*;/constexpr int kGroupWarps = $2;/; s/^constexpr int kProducers = [0-9]*;/constexpr int kProducers = $3;/; s/^constexpr int kStoreWarps = [0-9]*;/constexpr int kStoreWarps = $4;/" $F
python -m matterport_maskrcnn_with_tensorflow_serving_b200.build --force | tail -1
/usr/local/graft/bin/gpurun --timeout 300 -- "for c in $5; do echo -n 'G=$1x$2 P=$3 S=$4 chunk='\$c' '; timeout 40 python tools/quick_bench.py --chunks \$c --iters 20 2>&1 | tail -1 | cut -c1-110; done" 2>&1 | grep "G="
