#!/bin/bash
# parity of the wavefront forms after a kernel change + precise A/B on the 10M-triangle workloads (+ experiment libraries if any)
OUT=gpurun_out/${1:-r02n}
WL=${2:-c3s,c4s}
mkdir -p $OUT
timeout 1500 python -m pytest ${3:-tests} -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 5 $OUT/pytest_gpu.log
timeout 900 python tools/flavour_ab.py --flavours precise --tag main --workloads $WL > $OUT/ab_main.json 2> $OUT/ab_main.err; cat $OUT/ab_main.json; tail -n 3 $OUT/ab_main.err
for lib in hiprt-path-tracer_b200/libb2pt_*.so; do
  [ -e "$lib" ] || continue
  tag=$(basename $lib .so)
  B2PT_LIB=$lib timeout 900 python tools/flavour_ab.py --flavours precise --tag $tag --workloads $WL > $OUT/ab_$tag.json 2> $OUT/ab_$tag.err; cat $OUT/ab_$tag.json; tail -n 3 $OUT/ab_$tag.err
done
rm -rf gpurun_out/scenes
