#!/usr/bin/env python
"""bench.py -- particle-scan updates/s of the RBPF per-scan update on B200.

A "step" is one lidar event of the hot path for the whole particle set: odometry
motion + scan match + likelihood weighting + ray-cast map integration +
systematic resampling (main.py:139-160 of the reference), on the synthetic
workload of BASELINE.json configs[4] (200 m x 200 m world, 360-beam sweeps,
65,536 particles in total, sharded over the GPUs: strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                           the CPU oracle port of the
                                                                 reference path on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REFINE_DEFAULT = 0      # --refine: NDT stage of the matcher on (1) or off (0)
METRIC = "particle_scan_updates_per_sec"
UNIT = "updates/s"
# compulsory cells of the matcher footprint for a 180-degree sweep,
# (240/360) * pi * 11.7^2 / 0.0025 cells, 1 byte each (int8 tenths)
MATCH_BYTES_PER_UPDATE = (240.0 / 360.0) * np.pi * 11.7 ** 2 / 0.0025
# dram__bytes_read.sum + dram__bytes_write.sum of one launch over 8,192 particles, per update (`ncu --set full`,
# profiles/r2_*_raw.csv).  The matcher's is below its algorithmic figure: only blocks under the reference-set mask
# are fetched and particles that share sub-tiles after a resample hit in L2.
MATCH_DRAM_BYTES_PER_UPDATE_NCU = (549.42e6 + 8.45e6) / 8192       # match_kernel, profiles/r2_match_v9_raw.csv
CAST_DRAM_BYTES_PER_UPDATE_NCU = (472.81e6 + 321.70e6) / 8192      # raycast_cast2_kernel, profiles/r2_cast2_v2_raw.csv
WEIGHT_DRAM_BYTES_PER_UPDATE_NCU = (86.91e6 + 2.62e6) / 8192        # weight_kernel, profiles/r2_weight_v3_raw.csv
PREPARE_DRAM_BYTES_PER_UPDATE_NCU = (134.72e6 + 91.77e6) / 8192     # raycast_prepare_kernel, profiles/r2_prepare_raw.csv


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=65536, help="total over all GPUs")
    ap.add_argument("--beams", type=int, default=360)
    ap.add_argument("--burnin", type=int, default=30, help="untimed scans before warm-up (particle divergence)")
    ap.add_argument("--cpu-particles", type=int, default=0, help="CPU baseline sample (0 = 4 per core)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline: stop after this many seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dist-parity", type=int, default=1024,
                    help="N > 1: particles per rank of the untimed sharded-vs-single-set identity check (0 = skip)")
    ap.add_argument("--refine", type=int, default=REFINE_DEFAULT, choices=[0, 1],
                    help="NDT refinement stage of the reference matcher (matchScanCustom.m:32-50) after the grid search")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def port_baseline(work, n_particles, max_seconds, beams, refine):
    """The oracle (C port of the reference path, OpenMP over particles) on a
    bounded sample of the same workload.  Returns (updates/s, cores, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    cores = O.set_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1
    O.set_refine(bool(refine))
    if n_particles <= 0:
        n_particles = 4 * cores
    f = O.Filter(n_particles, beams, 30)
    rng = np.random.default_rng(5)
    f.set_scan(work.ranges[0], work.angles)
    f.integrate()
    f.integrate()
    # one untimed scan so that the first match sees a map and tiles are paged in
    f.motion(1, work.odom[0], work.dt, work.par)
    f.set_scan(work.ranges[1], work.angles)
    f.map_update(rng.standard_normal((n_particles, 30, 3)))
    f.resample(float(rng.random()))
    t0 = time.perf_counter()
    n_scans = 0
    for s in range(2, len(work.ranges)):
        f.motion(1, work.odom[s - 1], work.dt, work.par)
        f.set_scan(work.ranges[s], work.angles)
        f.map_update(rng.standard_normal((n_particles, 30, 3)))
        f.resample(float(rng.random()))
        n_scans += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    val = n_particles * n_scans / dt
    desc = "%d particles x %d scans of the same workload (%d beams), %.1f s" % (n_particles, n_scans, beams, dt)
    return val, cores, desc


def cpu_baseline(work, args):
    """The reference's CPU path on the box's host cores, on a bounded sample of the bench workload:
    the UNMODIFIED Python reference (oracle/_ref, vendored by oracle/vendor_ref.py; MATLAB's
    matchScanCustom answered by the oracle's restated matcher) on one core -- the reference is
    single-threaded, main.py:144,157 -- and fanned over all cores, and beside it the oracle's C port."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_baseline as RB

    cores = os.cpu_count() or 1
    pv, pc, pd = port_baseline(work, args.cpu_particles, args.cpu_seconds, args.beams, args.refine)
    port = {"value": pv, "unit": UNIT, "cores": pc, "kind": "port", "sample": pd}
    if not RB.available() or args.refine:
        return port
    one = RB.run(work.ranges, work.angles, work.odom, work.dt, 2, 3, workers=1)
    allc = RB.run(work.ranges, work.angles, work.odom, work.dt, 2 * cores, 3, workers=cores)
    return {"value": allc["updates_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "unmodified Python reference: %d particles x %d scans of the same workload in %d processes, %.1f s "
                      "(every process its own small filter; MATLAB matchScanCustom answered by the oracle's C matcher)"
                      % (allc["particles"], allc["scans"], cores, allc["seconds"]),
            "single_core": {"value": one["updates_per_s"], "cores": 1,
                            "sample": "%d particles x %d scans, %.1f s" % (one["particles"], one["scans"], one["seconds"])},
            "port": port}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores --
    the unmodified Python reference (oracle/_ref) fanned over all cores, one particle per core and step
    (a bounded sample: the reference needs about a second per particle-scan); the oracle's C port when
    the reference has not been vendored (or with --refine, which the Python reference run does not have)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from thesis_b200 import synth

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import ref_baseline as RB

    cores = os.cpu_count() or 1
    n_scans_total = 2 + args.warmup + args.steps
    work = synth.Workload(n_scans_total + 1, args.beams)
    if RB.available() and not args.refine:
        per = max(1, args.cpu_particles // cores) if args.cpu_particles > 0 else 1
        dt, n_cpu = RB.run_steps(work.ranges, work.angles, work.odom, work.dt, per, args.warmup, args.steps, cores)
        kind = "reference"
        sample = ("each step = %d particles x 1 scan (%d beams) of the same synthetic workload: the unmodified Python reference in "
                  "%d processes (one small filter each; MATLAB matchScanCustom answered by the oracle's C matcher)" % (n_cpu, args.beams, cores))
    else:
        cores = O.set_threads(cores)                  # torchrun exports OMP_NUM_THREADS=1
        O.set_refine(bool(args.refine))
        n_cpu = args.cpu_particles if args.cpu_particles > 0 else 8 * cores
        f = O.Filter(n_cpu, args.beams, 30)
        rng = np.random.default_rng(5)
        f.set_scan(work.ranges[0], work.angles)
        f.integrate()
        f.integrate()

        def one(s):
            f.motion(1, work.odom[s - 1], work.dt, work.par)
            f.set_scan(work.ranges[s], work.angles)
            f.map_update(rng.standard_normal((n_cpu, 30, 3)))
            f.resample(float(rng.random()))

        s = 1
        for _ in range(args.warmup):
            one(s)
            s += 1
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one(s)
            s += 1
        dt = time.perf_counter() - t0
        kind = "port"
        sample = "each step = %d particles x 1 scan (%d beams) of the same synthetic workload, oracle C port with OpenMP" % (n_cpu, args.beams)
    val = n_cpu * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {
        "workload": "configs[4]: synthetic 200 m x 200 m world, %d particles x %d-beam sweeps" % (args.particles, args.beams),
        "particles_total": args.particles, "particles_per_gpu": args.particles // world, "beams": args.beams,
        "samples_per_particle": 30, "cell_m": 0.05, "parallelism": "particles sharded x%d" % world,
        "cell_dtype": "int8 log-odds tenths", "state_dtype": "f64 poses / covariances / weights",
        "l2_policy": "per-step working set (page tables + touched sub-tiles of all particles) exceeds the 126 MB L2; no flush",
        # where the workload departs from SURVEY 8d config 5 (thesis_b200/synth.py): 5 % clutter would leave no free
        # space to range over, so 0.05 %; odometry = truth + N(0, (3 cm, 3 cm, 0.02 rad)) per scan, enough that the
        # zero correction isValidPose rejects is rarely the optimum; a 40 m loop through the door centres
        "world": "200 m x 200 m, walls every 10 m with 1 m doors, clutter fraction 0.0005",
        "odometry_noise": "N(0, (0.03 m, 0.03 m, 0.02 rad)) per scan on the velocity-family increments",
        "trajectory": "closed 40 m x 40 m loop, 0.35 m or pi/10 per scan (the reference's update gate is 0.33 m / pi/9)",
    }


def run_b200(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    from thesis_b200 import synth
    from thesis_b200.particles import ParticleSet

    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner on stdout when the first communicator is created:
        # keep stdout to the one JSON line by pointing fd 1 at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        from thesis_b200.dist import ShardedParticleSet, dist_parity_check
    # untimed, before anything else: is the sharded run the same computation as a single set?
    dist_parity = None
    if world > 1 and args.dist_parity > 0:
        dist_parity = dist_parity_check(args.dist_parity, args.beams, frames=8, device=local_rank)
    n_local = args.particles // world
    n_scans = 1 + args.burnin + args.warmup + args.steps + 2
    work = synth.Workload(n_scans, args.beams)
    pool = int(n_local * 26 + 4096)
    if world > 1:
        ps = ShardedParticleSet(n_local, args.beams, world_tiles=(5, 5), pool_subtiles=pool, device=local_rank, seed=7, ndt_refine=bool(args.refine))
    else:
        ps = ParticleSet(n_local, args.beams, world_tiles=(5, 5), pool_subtiles=pool, device=local_rank, seed=7, ndt_refine=bool(args.refine))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(s):
        ps.motion(1, work.odom[s - 1], work.dt, work.par)
        ps.step(work.ranges[s], work.angles)

    ps.set_scan(work.ranges[0], work.angles)
    ps.integrate()
    ps.integrate()                                          # map seeding, main.py:89-90
    s = 1
    for _ in range(args.burnin + args.warmup):
        one(s)
        s += 1
    ps.synchronize()
    st0 = ps.stats()
    # The timed loop and the end-to-end loop run the SAME scans from the same state: the whole particle
    # set is snapshotted on the device here and rewound in between (rbpf_snapshot / rbpf_restore).
    ps.snapshot()
    s0 = s
    # ---- timed region: device-resident inputs apart from the 360-double sweep ----
    ps.timing_enable(args.steps)
    ph0 = ps.match_phase_clocks() if hasattr(ps, "match_phase_clocks") else None
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one(s)
        s += 1
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    stage_ms, nst = ps.timing_read()
    ph1 = ps.match_phase_clocks() if ph0 is not None else None
    ps.timing_enable(0)
    ps.synchronize()
    st1 = ps.stats()
    # ---- end-to-end through the public API with host buffers, same scans ----
    ps.restore()
    s = s0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2h = 0
    for _ in range(args.steps):
        one(s)
        s += 1
        poses = ps.poses                                      # device -> host, what the Robot views expose
        weights = ps.weights
        d2h = poses.nbytes + weights.nbytes
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    st = ps.stats()
    if getattr(ps, "_prof", None):
        print("rank %d resample phases (ms/step): %s  migrated particles/step %.1f, MB/step %.2f" % (
            rank, {k: round(1e3 * v / ps._prof["n"], 3) for k, v in ps._prof.items() if k != "n"},
            ps.migrated_particles / max(ps._prof["n"], 1), ps.migrated_bytes / 1e6 / max(ps._prof["n"], 1)), file=sys.stderr)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json, sustained copy)" if peaks else "fallback 6.65 TB/s"
    value = args.particles * args.steps / (ms / 1e3)
    e2e = args.particles * args.steps / (ms_e2e / 1e3)
    # ---- roofline entries: ALGORITHMIC bytes per update x updates per launch
    #      / the launch's CUDA-event duration on the handle's stream, against the measured copy bandwidth
    a_r = float(np.mean([work.ray_cells(i) for i in range(s0, s0 + args.steps)]))
    cow = (st1["cow_copies"] - st0["cow_copies"]) / max(args.steps, 1)
    fresh = (st1["fresh_allocs"] - st0["fresh_allocs"]) / max(args.steps, 1)
    per_update = {
        "match_kernel": ("match", MATCH_BYTES_PER_UPDATE, "int8 cells of the 240-degree sector of radius 11.7 m the search can reach",
                         MATCH_DRAM_BYTES_PER_UPDATE_NCU),
        "raycast_cast2_kernel": ("raycast_cast", 2.0 * a_r, "read + write of the int8 cells under the rays (A_r = sum min(r, 15 m) / 5 cm)", CAST_DRAM_BYTES_PER_UPDATE_NCU),
        "raycast_prepare_kernel": ("raycast_prepare", (2.0 * cow + fresh) * 25600.0 / n_local,
                                   "copy-on-write: 2 x 25,600 B per shared sub-tile made private, 25,600 B per fresh one", PREPARE_DRAM_BYTES_PER_UPDATE_NCU),
        "weight_kernel": ("weight", 32.0 * args.beams, "one 32-byte sector per beam (the 30 samples of a beam fall into the same cells)", WEIGHT_DRAM_BYTES_PER_UPDATE_NCU),
    }
    rooflines = {}
    for kname, (stage, bpu, what, dram) in per_update.items():
        ms_k = stage_ms[stage] / max(nst, 1)
        ach = bpu * n_local / (ms_k / 1e3) / 1e9 if ms_k > 0 else 0.0
        rooflines[kname] = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": dram * n_local if dram else None, "algorithmic_bytes_per_update": bpu, "what": what,
                            "updates_per_launch": n_local, "launch_ms": ms_k, "peak_source": peak_src}
    dominant = max(rooflines.values(), key=lambda r: r["launch_ms"])
    for r_ in rooflines.values():
        r_["traffic_source"] = "ncu --set full on an 8,192-particle launch (profiles/r2_*_raw.csv), scaled per update"
    rooflines["match_kernel"]["note"] = ("the correlative search is bound by shared-memory wavefronts and the ALU pipe (ncu: 69 % / 65 % "
                                         "of peak), not by HBM; see DESIGN.md")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "i8", "data": "synthetic",
        "config": dict(workload_config(args, world), burnin_scans=args.burnin,
                       ray_cells_per_scan=a_r, pool_subtiles=pool, cow_copies_per_scan=cow, fresh_subtiles_per_scan=fresh,
                       unique_subtile_fraction=1.0 - st["shared_refs"] / max(st["total_refs"], 1),
                       pool_in_use=st["pool_in_use"], match_failed=st["match_failed"], match_failed_zero_correction=st["match_failed_zero"],
                       match_failed_fraction=st["match_failed"] / max(1, n_local * (s - 1)), resamples=st["resamples"],
                       match_searches_run_fraction=st["match_runs"] / max(1, n_local * (s - 1)),
                       match_scoring_passes_per_update=st["match_evals"] / max(1, st["match_runs"]),
                       match_exhaustive_passes_per_update=231, ndt_refine=bool(args.refine),
                       ndt_evaluations_per_search=st["ndt_evals"] / max(1, st["match_runs"]),
                       ndt_accepted_fraction=st["ndt_accepted"] / max(1, st["match_runs"]),
                       match_full_pass_equivalents_per_update=st["match_visits"] / max(1, st["match_points"])),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(2 * args.beams * 8 + 4 * 8 + 4 * 8), "d2h_bytes_per_step": int(d2h),
                "note": "the same K scans as the device-timed loop, from the same device-side snapshot of the particle set; every step "
                        "copies the sweep from pinned host memory and reads all poses and weights back"},
        # kernels of _librbpf.so per step (ncu launch list profiles/r2_launches_65536p.csv): motion, match, match_copy_dups,
        # weight x2 (samples + fallback), raycast prepare + cast, resample plan + ancestors, gather, refs = 11; sharded runs
        # add the five pull kernels (claim, alloc, copy, place, release); NCCL's own kernels and memsets are not counted
        "gpu_launches": int(args.steps * (11 if world == 1 else 16)),
        "roofline": dominant,                               # the kernel with the longest launch
        "rooflines": rooflines,
        "stage_ms_per_step": {k: v / max(nst, 1) for k, v in stage_ms.items()},
    }
    if dist_parity is not None:
        line["dist_parity"] = dist_parity
    if ph1 is not None:
        d = {k: ph1[k] - ph0[k] for k in ph1}
        cta = sum(v for k, v in d.items() if k.startswith("cta_")) or 1
        line["match_phase_share"] = {k[4:]: round(v / cta, 4) for k, v in d.items() if k.startswith("cta_")}
        line["match_warp_busy_share"] = {k[5:]: round(d[k] / (12.0 * d["cta_" + c] or 1), 4) for k, c in (
            ("warp_seeds", "seeds_bounds"), ("warp_bounds", "seeds_bounds"), ("warp_members", "members"))}
        line["match_visits_share"] = {k[7:]: round(d[k] / (sum(d[q] for q in d if q.startswith("visits_")) or 1), 4)
                                      for k in d if k.startswith("visits_")}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(work, args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
