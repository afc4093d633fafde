for Q in 8 7 6 4; do echo "qw=$Q 8192: $(PGTT_QUAD_WARPS=$Q python tools/step_time.py stairs 8192 level07 100 1 2>&1 | grep -E 'back' | cut -c1-90)"; done
for Q in 8 7; do echo "qw=$Q 32768: $(PGTT_QUAD_WARPS=$Q python tools/step_time.py stairs 32768 level1 60 0 2>&1 | grep -E 'back' | cut -c1-90)"; done
for Q in 8 7; do echo "qw=$Q 16384: $(PGTT_QUAD_WARPS=$Q python tools/step_time.py stairs 16384 level1 60 0 2>&1 | grep -E 'back' | cut -c1-90)"; done
