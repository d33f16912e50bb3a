"""Print the interesting fields of a bench.py JSON line (development aid)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    print(path, "value %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    print("  stages", {k: round(v, 3) for k, v in d.get("stage_ms_per_step", {}).items()})
    for k in ("match_phase_share", "match_warp_busy_share", "match_visits_share"):
        if k in d:
            print(" ", k, d[k])
    c = d["config"]
    print("  passes %.1f  fpe %.1f  runfrac %.3f  failed %d  uniq %.3f pool_in_use %d" % (
        c["match_scoring_passes_per_update"], c["match_full_pass_equivalents_per_update"],
        c["match_searches_run_fraction"], c["match_failed"], c["unique_subtile_fraction"], c["pool_in_use"]))
