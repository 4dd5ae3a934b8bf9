#!/bin/bash
# precise A/B of the main build (+ experiment libraries) on the given workloads, no tests: bash tools/gpu_ab_only.sh TAG WORKLOADS
OUT=gpurun_out/${1:-ab}
WL=${2:-c3s,c4s}
mkdir -p $OUT
timeout 900 python tools/flavour_ab.py --flavours precise --tag main --workloads $WL > $OUT/ab_main.json 2> $OUT/ab_main.err; cat $OUT/ab_main.json; tail -n 3 $OUT/ab_main.err
for lib in hiprt-path-tracer_b200/libb2pt_*.so; do
  [ -e "$lib" ] || continue
  tag=$(basename $lib .so)
  B2PT_LIB=$lib timeout 900 python tools/flavour_ab.py --flavours precise --tag $tag --workloads $WL > $OUT/ab_$tag.json 2> $OUT/ab_$tag.err; cat $OUT/ab_$tag.json; tail -n 3 $OUT/ab_$tag.err
done
rm -rf gpurun_out/scenes
