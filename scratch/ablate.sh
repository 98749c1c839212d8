#!/bin/bash
# timing ablation of stack_b phases (EXPERIMENT)
for m in 0 1 2 4 8 16 32 24 28 31 63; do
  SILENT_ABLATE=$m KB_TAG="ablate=$m" timeout 120 python scratch/kbench.py 2>&1 | tail -1
done
