mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused -s 1 -c 1 -o gpurun_out/f2_fused python tools/bench_divergence.py 200 5000000 > gpurun_out/f2_ncu.log 2>&1
tail -5 gpurun_out/f2_ncu.log
ls -la gpurun_out/
