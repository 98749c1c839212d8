python -m pytest tests -m gpu -x -q 2>&1 | tail -2
KB_TAG=htab python scratch/kbench.py 2>&1 | tail -1
