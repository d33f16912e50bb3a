# usage: bash tools/ab_parts.sh <particles> "ENV=1 ..." tag    one bench line at a per-GPU particle count (development aid)
env $2 python bench.py --particles $1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$3.json 2> gpurun_out/$3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$3.json').read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('$3', round(d['ms_per_step'],3), 'match',round(s['match'],3),'cast',round(s['raycast_cast'],3),'prep',round(s['raycast_prepare'],3),'w',round(s['weight'],3),'plan',round(s['resample_plan'],3),'apply',round(s['resample_apply'],3))
PY
