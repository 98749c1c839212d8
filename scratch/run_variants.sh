python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for tw in 36 42 48 64 72; do
  SILENT_PAIR_TILEW=$tw KB_TAG=tw$tw python scratch/kbench.py 2>&1 | tail -1
done
KB_BATCH=1 KB_TAG=b1 python scratch/kbench.py 2>&1 | tail -1
