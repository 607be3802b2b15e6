mkdir -p gpurun_out
for ws in 8 4 2; do for wc in 8 4 2 1; do JTB_W_STRIDED=$ws JTB_W_CONTIG=$wc python scripts/prof_fft3d.py; done; done > gpurun_out/sweep1.log 2>&1
cat gpurun_out/sweep1.log
