# stage on/off experiments on the two-plane pipeline (results are wrong with a flag set; timing only)
for d in 0 1 3 4 8 16 32 24 28 60 63; do
  echo "=== dbg $d"
  BGAN_TC_VARIANT=12 BGAN_TC_DBG=$d timeout 120 python tools/prof_blur.py --stds 5 --iters 300 --sets 6
done > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log
