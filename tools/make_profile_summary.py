"""Regenerate profiles/README.md from the ncu CSV exports and the bench JSON kept in profiles/.

    python tools/make_profile_summary.py
"""
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    return ["| %s | %d | %.3f | %.3f | %.1f %% |" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6, 100 * sum(v) / tot)
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]


WANT = [("gpu__time_duration.sum", "ms"), ("smsp__inst_executed.sum", "warp instr"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global st sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global st requests")]


def full(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        vals = []
        for w, label in WANT:
            v = r[hdr.index(w)] if w in hdr else ""
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append("%s=%s" % (label, v))
        out.append("* `%s`: " % name + ", ".join(vals))
    return out


def main():
    b = json.load(open(os.path.join(P, "r1_bench_n1_65536p.json")))
    md = ["# profiles/ -- round 1", "",
          "All captures on a B200 (sm_100a, 148 SMs, 1965 MHz) through `gpurun`. ncu runs use `--clock-control none`;",
          "per-launch times under ncu are cold-cache and serialised, so compare SHARES with the live CUDA-event",
          "numbers, not absolutes.  Regenerate this file with `python tools/make_profile_summary.py`.", "",
          "## Live bench (no profiler): `python bench.py` (65,536 particles x 360 beams, configs[4], N=1)", "",
          "`r1_bench_n1_65536p.json` -- %.0f updates/s, %.2f ms/step; e2e %.0f updates/s; CPU oracle port %.0f updates/s on %d cores."
          % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["cpu_baseline"]["value"], b["cpu_baseline"]["cores"]), "",
          "| stage (CUDA events inside rbpf_step) | ms/step | share |", "|---|---|---|"]
    tot = sum(b["stage_ms_per_step"].values())
    for k, v in b["stage_ms_per_step"].items():
        md.append("| %s | %.3f | %.1f %% |" % (k, v, 100 * v / tot))
    md += ["", "Matcher: %.1f bitmap scoring passes per update after branch and bound (exhaustive = 231)."
           % b["config"]["match_scoring_passes_per_update"],
           "Roofline entry of the JSON line: match_kernel, %.0f GB/s algorithmic = %.1f %% of the measured %.0f GB/s -- the"
           % (b["roofline"]["achieved"], 100 * b["roofline"]["frac"], b["roofline"]["peak"]),
           "kernel is bound by instruction issue / shared-memory lookups (see below), not by HBM.", ""]
    nd = os.path.join(P, "r1_bench_n1_65536p_ndt.json")
    if os.path.exists(nd):
        n = json.load(open(nd))
        md += ["With the NDT refinement stage of the matcher on (`bench.py --refine 1`, `r1_bench_n1_65536p_ndt.json`): %.0f updates/s,"
               % n["value"],
               "%.2f ms/step, match stage %.2f ms; %.1f NDT score evaluations per search, %.0f %% of the refined poses accepted."
               % (n["ms_per_step"], n["stage_ms_per_step"]["match"], n["config"]["ndt_evaluations_per_search"],
                  100 * n["config"]["ndt_accepted_fraction"]),
               "An `ncu --set full` source view of that launch (8,192 particles) puts 43 % of the kernel's warp samples in the NDT",
               "stage: 52 % of those in the per-beam fp64 evaluation, 11 % waiting for the slowest warp at the first barrier, 36 %",
               "waiting for warp 0 (totals, 3x3 Cholesky solves, sin/cos of the next proposal) at the second.", ""]
    sc = []
    for g in (2, 4, 8):
        f = os.path.join(P, "r1_scale_n%d.json" % g)
        if os.path.exists(f):
            sc.append((g, json.load(open(f))))
    if sc:
        md += ["## Strong scaling, 65,536 particles over N GPUs (`torchrun ... bench.py --gpus N --steps 20 --warmup 5`)", "",
               "| GPUs | updates/s | ms/step | efficiency vs N=1 | match | weight | ray-cast | resample + exchange |", "|---|---|---|---|---|---|---|---|",
               "| 1 | %.0f | %.2f | 100 %% | %.2f | %.2f | %.2f | %.2f |"
               % (b["value"], b["ms_per_step"], b["stage_ms_per_step"]["match"], b["stage_ms_per_step"]["weight"],
                  b["stage_ms_per_step"]["raycast_prepare"] + b["stage_ms_per_step"]["raycast_cast"] + b["stage_ms_per_step"]["weight_fallback"],
                  b["stage_ms_per_step"]["resample_plan"] + b["stage_ms_per_step"]["resample_apply"])]
        for g, d in sc:
            st = d["stage_ms_per_step"]
            md.append("| %d | %.0f | %.2f | %.0f %% | %.2f | %.2f | %.2f | %.2f |"
                      % (g, d["value"], d["ms_per_step"], 100 * d["value"] / (g * b["value"]), st["match"], st["weight"],
                         st["raycast_cast"], st["resample_plan"]))
        md += ["", "`r1_scale_n{2,4,8}.json`.  Exchange = NCCL all-gather of the weights, the global plan on every rank, pull of the",
               "remote ancestors' page tables and sub-tiles over NVLink peer mappings, a one-element all-reduce as barrier,",
               "local gather / reference counts (thesis_b200/dist.py, transport \"peer\").", ""]
    w8 = os.path.join(P, "r1_bench_n1_8192p.json")
    s8 = os.path.join(P, "r1_scale_n8.json")
    if os.path.exists(w8) and os.path.exists(s8):
        a1, a8 = json.load(open(w8)), json.load(open(s8))
        md += ["Weak scaling at 8,192 particles per GPU (SURVEY 8d): 1 GPU x 8,192 = %.0f updates/s (%.2f ms/step, `r1_bench_n1_8192p.json`),"
               % (a1["value"], a1["ms_per_step"]),
               "8 GPUs x 8,192 = %.0f updates/s (%.2f ms/step): %.0f %% -- the exchange adds %.2f ms to a rank's scan." %
               (a8["value"], a8["ms_per_step"], 100 * a8["value"] / (8 * a1["value"]), a8["ms_per_step"] - a1["ms_per_step"]), ""]
    fr = os.path.join(P, "r1_bench_n1_65536p_fresh.json")
    if os.path.exists(fr):
        f = json.load(open(fr))
        md += ["Divergence regimes (SURVEY 8d): the headline line is the DIVERGED regime (30 burn-in scans with resampling, unique",
               "sub-tile fraction %.2f: descendants of one ancestor share what they have not written since); FRESH (`bench.py --burnin 0 --warmup 3 --steps 5`, scans 4-8 of a new map, unique fraction %.2f,"
               % (b["config"]["unique_subtile_fraction"], f["config"]["unique_subtile_fraction"]),
               "`r1_bench_n1_65536p_fresh.json`): %.0f updates/s, %.2f ms/step." % (f["value"], f["ms_per_step"]), ""]
    for tag, title in (("final", "final kernels"),
                       ("baseline", "first working version: exhaustive 231-rotation matcher, per-cell closed-form ray-cast")):
        f = os.path.join(P, "r1_launches_8192p_%s.csv" % tag)
        if os.path.exists(f):
            md += ["## Launch list (%s): `ncu --metrics gpu__time_duration.sum` on `bench.py --particles 8192 --steps 3 --warmup 3 --burnin 6`"
                   % title, "", "`%s`" % os.path.basename(f), "", "| kernel | launches | total ms | avg ms | share |",
                   "|---|---|---|---|---|"] + launches(f) + [""]
    md += ["The kernel shares of the final launch list agree with the live per-stage CUDA-event split above.", ""]
    for tag in ("final", "baseline"):
        f = os.path.join(P, "r1_full_8192p_%s_raw.csv" % tag)
        if os.path.exists(f):
            md += ["## `ncu --set full` (%s), one launch per kernel: `%s`" % (tag, os.path.basename(f)), ""] + full(f) + [""]
    md += ["## Reading", "",
           "* `match_kernel`: ~70 % of issue slots busy, DRAM a few %: bound by instruction issue (LOP3 carry-save adders, LDS,",
           "  address arithmetic) of the bit-parallel scoring passes; the levers were fewer passes (exact branch and bound over",
           "  rotation groups, seeding with the rotations around the guess, admissible early abort: 231 -> ~31 full-pass",
           "  equivalents per update) and fewer instructions per pass.",
           "* `raycast_cast_kernel`: byte read-modify-writes along rays.  Baseline: ~15 sectors per store request, bound by L1/L2",
           "  sector traffic; not storing unchanged (saturated) cells, the 8x4-cell sector blocks and the interior fast path",
           "  took it from 29.4 to ~15 ms at 65,536 particles; now ~80 % of issue slots busy, about half of them per-beam set-up.",
           "* `raycast_prepare_kernel`: copy-on-write sub-tile copies, DRAM-bound as intended.",
           "* `weight_kernel`: latency-bound lookups, one warp per particle, 4 lookups in flight per lane."]
    open(os.path.join(P, "README.md"), "w").write("\n".join(md) + "\n")
    print("wrote profiles/README.md")


if __name__ == "__main__":
    main()
