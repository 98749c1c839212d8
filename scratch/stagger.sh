for ns in 0 3000 6000 12000 25000; do SILENT_STAGGER_NS=$ns KB_TAG="stagger=$ns" timeout 120 python scratch/kbench.py 2>&1 | tail -1; done
