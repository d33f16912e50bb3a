# usage: bash tools/ab65k.sh "ENV=1 ..." tag     one 65,536-particle bench line with the given environment (development aid)
env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$2.json 2> gpurun_out/$2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$2.json').read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('$2', round(d['ms_per_step'],2), 'match',round(s['match'],2),'cast',round(s['raycast_cast'],2),'prep',round(s['raycast_prepare'],2),'w',round(s['weight'],2),'plan',round(s['resample_plan'],3),'apply',round(s['resample_apply'],3), 'failed', d['config']['match_failed'])
PY
