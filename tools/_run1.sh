mkdir -p gpurun_out
rm -f gpurun_out/f6_div.log
for shape in "200 625000" "200 100000" "27 10000000" "27 2000000" "4 10000000" "100 1000000" "200 2500000"; do
  for f in 1 0; do
  echo "== $shape fused=$f" >> gpurun_out/f6_div.log
  ABFIT_DEV_DIV_FUSED=$f timeout 120 python tools/bench_divergence.py $shape 2>&1 | tail -3 | head -2 >> gpurun_out/f6_div.log
  done
done
cat gpurun_out/f6_div.log | cut -c1-200
