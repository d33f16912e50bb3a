# usage: bash tools/ab65k_match.sh "ENV=1 ..." tag    like ab65k.sh, plus the matcher's phase split in ms (development aid)
env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$2.json 2> gpurun_out/$2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$2.json').read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
m=s['match']
print('$2', round(d['ms_per_step'],2), 'match',round(m,2),'cast',round(s['raycast_cast'],2),'w',round(s['weight'],2), 'failed', d['config']['match_failed'], 'fpe', round(d['config']['match_full_pass_equivalents_per_update'],1))
print('   ', {k: round(v*m,2) for k,v in d['match_phase_share'].items()})
PY
