# usage: bash tools/run_variants.sh base var1 var2 ...   (variants built by thesis_b200/build.py --variant)
for v in "$@"; do
  if [ $v = base ]; then unset RBPF_LIB; else export RBPF_LIB=$PWD/thesis_b200/_var_$v.so; fi
  python bench.py --particles 16384 --steps 8 --warmup 3 --burnin 25 --no-cpu-baseline $BENCH_EXTRA 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$v', round(d['ms_per_step'],2), 'match',round(s['match'],2),'cast',round(s['raycast_cast'],2),'prep',round(s['raycast_prepare'],2),'w',round(s['weight'],2),'plan',round(s['resample_plan'],3),'apply',round(s['resample_apply'],3),'passes',round(d['config']['match_scoring_passes_per_update'],1),'equiv',round(d['config']['match_full_pass_equivalents_per_update'],1),'runfrac',round(d['config']['match_searches_run_fraction'],3),'failed',d['config']['match_failed'],'ndt',round(d['config']['ndt_evaluations_per_search'],1),round(d['config']['ndt_accepted_fraction'],2))"
done
