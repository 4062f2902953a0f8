#!/bin/bash
# A/B of programmatic dependent launch on ONE box: the product library against a copy built with -DB2D_NO_PDL, bench runs
# interleaved (boxes differ by ~3 % in power-capped clocks, so only same-box pairs mean anything).
#   here:        bash tools/pdl_ab.sh build
#   on the box:  bash tools/pdl_ab.sh run
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  d=$(mktemp -d); cp finetrainers_b200/csrc/*.cu finetrainers_b200/csrc/*.cuh finetrainers_b200/csrc/*.h $d/
  sed -i "s#\"../../include/b2d.h\"#\"$PWD/include/b2d.h\"#" $d/*
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       -diag-suppress 177 -shared -cudart static -DB2D_NO_PDL -o tools/micro/libb2d_nopdl.so $d/*.cu
  echo built tools/micro/libb2d_nopdl.so
else
  cp finetrainers_b200/libb2d.so /tmp/b2d_pdl.so
  for i in 1 2; do
    for v in pdl nopdl; do
      [ $v = pdl ] && cp /tmp/b2d_pdl.so finetrainers_b200/libb2d.so || cp tools/micro/libb2d_nopdl.so finetrainers_b200/libb2d.so
      python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
    done
  done
  cp /tmp/b2d_pdl.so finetrainers_b200/libb2d.so
fi
