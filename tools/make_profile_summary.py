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
           "  rotation groups, seeding with the rotations around the guess, admissible early abort: 231 -> ~34 full-pass",
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
