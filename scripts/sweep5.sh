for r in 1 2 4 8; do JTB_FAST_REPS=$r python scripts/prof_fft3d.py; done
