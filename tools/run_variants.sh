#!/bin/bash
# Development aid (GPU box): time every experiment build libhm_matcher_<name>.so present in the package directory
# with tools/time_prepared.py at the C4 shape, the 8-GPU shard shape and three short-CTA shapes.
cd "$(dirname "$0")/.."
SHAPES=${SHAPES:-"2000x8192000 2000x1024000 16384x16384 10000x10000 2000x2000"}
for so in slam_experiments_b200/libhm_matcher.so slam_experiments_b200/libhm_matcher_*.so; do
  [ -f "$so" ] || continue
  case "$so" in *trace*) continue;; esac
  for shape in $SHAPES; do
    HM_MATCHER_SO=$PWD/$so timeout 120 python tools/time_prepared.py ${shape%x*} ${shape#*x} 2>&1 | tail -1
  done
done
