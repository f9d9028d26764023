# A/B of the tensor-core blur schedule variants (experiments; see profiles/)
for v in 0 3; do for pf in 0 1 2 3 5; do
  echo "=== variant $v pf $pf"
  BGAN_TC_VARIANT=$v BGAN_TC_PF=$pf python tools/prof_blur.py --stds 5 --iters 300 --sets 6
done; done > gpurun_out/ab_variants.log 2>&1
for v in 0 3; do for pf in 1 5; do
  BGAN_TC_VARIANT=$v BGAN_TC_PF=$pf python tools/trace_tc.py
done; done > gpurun_out/ab_trace.log 2>&1
cat gpurun_out/ab_variants.log
