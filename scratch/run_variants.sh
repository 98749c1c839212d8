python -m pytest tests -m gpu -x -q 2>&1 | tail -3
KB_TAG=r2f python scratch/kbench.py 2>&1 | tail -1
KB_BATCH=1 KB_TAG=r2f_b1 python scratch/kbench.py 2>&1 | tail -1
