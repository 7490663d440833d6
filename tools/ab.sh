#!/bin/bash
# A/B timing of library variants in ONE gpurun call (box-to-box noise is ~3 %): tools/ab.sh "<bench args>" variants/a.so variants/b.so ...
args="$1"; shift
for rep in 1 2; do
  for v in "$@"; do
    ms=$(SEPT_LIB_PATH=$v timeout 200 python bench.py --steps 30 --warmup 3 --no-extras $args 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f' % d['ms_per_step'])")
    echo "$rep $args $v $ms"
  done
done
