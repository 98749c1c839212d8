python -m pytest tests -m gpu -x -q 2>&1 | tail -4
KB_TAG=ha16 python scratch/kbench.py 2>&1 | tail -1
KB_BATCH=1 KB_TAG=ha16_b1 python scratch/kbench.py 2>&1 | tail -1
SILENT_TILE_HA=22 python -m pysilent_b200.build --force > /dev/null 2>&1
KB_TAG=ha22 python scratch/kbench.py 2>&1 | tail -1
python -m pytest tests -m gpu -x -q -k "fused_stack or baseline_configs" 2>&1 | tail -2
