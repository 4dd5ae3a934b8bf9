#!/bin/bash
# wavefront MegaKernel: parity against the other forms, then A/B timing against the phased form
OUT=gpurun_out/${1:-wf}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "forms or schedule" > $OUT/pytest_forms.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_forms.log; tail -n 15 $OUT/pytest_forms.log
for form in phased wavefront; do
  B2PT_MEGAKERNEL=$form timeout 900 python tools/flavour_ab.py --flavours ${3:-precise} --tag $form --workloads ${2:-c2,c3s,c4s} > $OUT/ab_$form.json 2> $OUT/ab_$form.err; cat $OUT/ab_$form.json; tail -n 3 $OUT/ab_$form.err
done
rm -rf gpurun_out/scenes
