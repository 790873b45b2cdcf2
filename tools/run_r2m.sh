set -x
mkdir -p gpurun_out
timeout 400 python tools/tc_sweep.py --scale 24 --reps 4 --configs '[{}, {"tile_shift": 88}, {"tile_shift": 87}, {"tile_shift": 86}, {"tile_shift": 89}, {"tile_shift": 90}, {"tile_shift": 88, "item_cost": 131072}, {"tile_shift": 87, "item_cost": 131072}, {"tile_shift": -1}]' > gpurun_out/r2m_tile_sweep.jsonl 2> gpurun_out/r2m_tile_sweep.err
python - <<'P'
import json
for l in open('gpurun_out/r2m_tile_sweep.jsonl'):
    d = json.loads(l)
    if 'cfg' in d: print(d['cfg'], round(d['ms_count'], 2), round(d['ms_bitmap'], 2), d.get('bitmap_items'))
P
tail -3 gpurun_out/r2m_tile_sweep.err
