for hb in 8 12 16; do
  SILENT_TILE_HB=$hb python -m pysilent_b200.build --force > /dev/null 2>&1
  KB_TAG=hb$hb python scratch/kbench.py 2>&1 | tail -1
done
SILENT_TILE_HB=12 python -m pysilent_b200.build --force > /dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
