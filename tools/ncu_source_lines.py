"""Per-source-line summary of one kernel from an .ncu-rep captured with
`ncu --set full --import-source on` (lines compiled with -lineinfo).

  python tools/ncu_source_lines.py report.ncu-rep [top_n] [--md]

Runs `ncu -i ... --page source --print-source cuda,sass --csv` and prints, per CUDA source line,
the share of warp-stall samples and of executed warp instructions."""
import csv
import io
import subprocess
import sys


def lines_of(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    data, fname, hdr = [], "", None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 6 and r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
        elif hdr and len(r) > 6 and r[0].isdigit():
            try:
                data.append((int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    return data


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
    md = "--md" in sys.argv
    d = lines_of(rep)
    ts, ti = sum(x[0] for x in d) or 1, sum(x[1] for x in d) or 1
    print("samples %d, warp instructions %d, source lines %d" % (ts, ti, len(d)))
    if md:
        print("| samples | instr | line | source |\n|---|---|---|---|")
    for s, i, f, ln, src in sorted(d, reverse=True)[:top]:
        if md:
            print("| %.1f %% | %.1f %% | %s:%d | `%s` |" % (100 * s / ts, 100 * i / ti, f, ln, src[:100].replace("|", "\\|")))
        else:
            print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100 * s / ts, 100 * i / ti, f, ln, src[:120]))
