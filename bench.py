#!/usr/bin/env python
"""bench.py — headline benchmark of the PTZ-Calib NLS hot path on B200 (contract in the task description, ④).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference-equivalent CPU path (oracle port), host cores
  python bench.py --gpus N --config cfg5                   BASELINE cfg 5 at its named size (V = 10,000, 5e7 observations in total)

Workload (config.workload): BASELINE cfg 4 "scaled synthetic PTZ-BA: 1,000 views x 2M observations, full LM with
Schur + PCG" on one GPU.  With N > 1 the scene grows with N (V = 1000 N views, ~2M observations per rank, tracks
sharded by observation, NCCL all-reduce of the camera blocks): weak scaling, cfg-5-shaped; --config cfg5 runs cfg 5 itself
(fixed total size, sharded N ways).
A STEP is one Levenberg-Marquardt iteration (stages 1-4).  Steps are drawn from complete solves run with the
reference's own tolerances: when a solve converges the problem is reset and solved again until K steps are done.
metric = observations x LM iterations per second (Mobs/s); lm_iters_per_sec, the small configurations (cfg 1 / cfg 2, LM
iterations per second), the batched-reloc and the track-building figures ride along.

roofline: the kernel with the largest share of the step (stage 3, k_cg: L2-latency / grid-barrier bound, reported against the HBM
peak all the same), then one entry per streaming kernel in `roofline_kernels`.  `traffic` comes from hardware counters read IN THIS
RUN: after the timed region rank 0 runs a few LM iterations of the same scene under ncu (tests/scripts/ncu_counters.py; nothing
timed there) and takes dram__bytes_read + dram__bytes_write per launch; null when ncu is not available.
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

RJ_BYTES_PER_OBS = {0: 160, 1: 176, 2: 192}  # read 16 + write r 16 + write J (16/18/20 doubles)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


NCU_METRICS = ("dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,"
               "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
NCU_KERNELS = "regex:k_cg|k_resjac|k_obs_what|k_track_factor|k_track_backsub|k_track_accum|k_schur_offdiag|k_schur_diag|k_cost|k_reloc"


def hardware_counters(args):
    """Per-launch DRAM bytes and fp64 operations of the hot kernels, from ncu, measured on this box in this run (side run, untimed):
    {kernel: {launches, dram_bytes, fp64_flops}} or None when ncu is missing / fails."""
    import csv
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    log = os.path.join(ROOT, "gpurun_out", "bench_ncu_counters.csv")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    cmd = [ncu, "--metrics", NCU_METRICS, "--clock-control", "none", "-k", NCU_KERNELS, "-c", "200", "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "tests", "scripts", "ncu_counters.py"), str(args.scale), str(args.factor_type), str(min(args.reloc_queries, 20000))]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        subprocess.run(cmd, check=True, timeout=600, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
        rows = [l for l in open(log) if not l.startswith("==")]
        agg = {}
        for r in csv.DictReader(rows):
            name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("ptz::", "").strip()
            k = agg.setdefault(name, dict(ids=set(), dram=0.0, dfma=0.0, dadd=0.0, dmul=0.0))
            k["ids"].add(r["ID"])
            v = float(r["Metric Value"].replace(",", ""))
            m = r["Metric Name"]
            if m.startswith("dram__bytes"):
                unit = r.get("Metric Unit", "byte").lower()
                v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
                k["dram"] += v
            elif "dfma" in m:
                k["dfma"] += v
            elif "dadd" in m:
                k["dadd"] += v
            elif "dmul" in m:
                k["dmul"] += v
        out = {}
        for name, k in agg.items():
            n = max(len(k["ids"]), 1)
            out[name] = dict(launches=n, dram_bytes=int(k["dram"] / n), fp64_flops=int((2 * k["dfma"] + k["dadd"] + k["dmul"]) / n))
        return out
    except Exception as e:  # noqa: BLE001
        return dict(error=str(e)[:200])


def counters_for(counters, prefix):
    """the entry of the first kernel whose name starts with `prefix` (template arguments vary with the factor type)"""
    if not counters or "error" in counters:
        return None
    for name, v in counters.items():
        if name.startswith(prefix):
            return v
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_scene(args, rank, world):
    from ptz_calib_b200 import synth

    if args.config == "cfg5":
        # BASELINE cfg 5 at its named size: V = 10,000 views, 1e7 tracks / 5e7 observations IN TOTAL, every rank generates its own
        # 1/world of the tracks over the same views
        V, P = max(8, int(10000 * args.scale)), int(10000000 * args.scale / world)
        return synth.make_ba_scene(V, P, "band", factor_type=args.factor_type, seed=synth.SEEDS[5], track_seed=950001 + rank)
    scale = args.scale
    V = max(8, int(1000 * scale)) * world
    P = int(400000 * scale)  # tracks generated per rank; >= 4 visible views survive
    return synth.make_ba_scene(V, P, "band", factor_type=args.factor_type, seed=synth.SEEDS[4], track_seed=900001 + rank)


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import ptz_calib_b200 as ptz
    from ptz_calib_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ptz.nccl_init_from_torch()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    prob = make_scene(args, rank, world)
    M_total = allsum(float(prob.M))
    K, W = args.steps, args.warmup
    opt = ptz.default_options(max_num_iterations=200, pcg_rel_tolerance=args.pcg_tol)

    # ---------------- device-resident timed region: K LM iterations drawn from complete solves ----------------
    h = ptz.BAHandle(prob, opt)

    def run_steps(n):
        """n LM iterations drawn from complete solves; returns (iterations done, solves completed, last result)"""
        done, solves, last = 0, 0, None
        while done < n:
            want = n - done
            last = h.run(want, want_outputs=False)
            got = last.num_iterations - run_steps.base
            done += got
            run_steps.base = last.num_iterations
            if got < want:  # the solve terminated (convergence / iteration cap): start the next one
                h.reset()
                run_steps.base = 0
                solves += 1
                if got == 0 and last.num_iterations == 0:
                    break  # degenerate problem that converges at iteration 0
        return done, solves, last

    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):  # (diagnostics only: the contract's line carries the clocks)
        sampler.start()
    run_steps.base = 0
    run_steps(W)
    h.reset()
    run_steps.base = 0

    def timed_pass(per_kernel):
        """K LM iterations from complete solves; returns the stage-time difference over the pass"""
        h.set_stage_timing(per_kernel)
        h.reset()
        run_steps.base = 0
        barrier()
        st_a = h.stage_times()
        t0 = time.perf_counter()
        done, solves, last = run_steps(K)
        st_b = h.stage_times()
        barrier()
        wall = time.perf_counter() - t0
        # timings accumulate over the handle's life: the pass is the difference
        st = dict(ms_run=st_b["ms_run"] - st_a["ms_run"], pcg_iterations=st_b["pcg_iterations"] - st_a["pcg_iterations"],
                  lm_iterations=max(st_b["lm_iterations"] - st_a["lm_iterations"], 1), kernels={})
        for name, kb in st_b["kernels"].items():
            ka = st_a["kernels"].get(name, dict(ms=0.0, launches=0))
            if kb["launches"] > ka["launches"]:
                st["kernels"][name] = dict(ms=kb["ms"] - ka["ms"], launches=kb["launches"] - ka["launches"], stage=kb["stage"])
        st["launches_total"] = sum(k["launches"] for k in st["kernels"].values())
        st["ms_kernels_total"] = sum(k["ms"] for k in st["kernels"].values())
        st["deflated_solves"] = st_b["deflated_solves"] - st_a["deflated_solves"]
        return st, st_b, done, solves, wall

    # the headline: the span of the ptzba_run calls with the per-kernel events OFF (two cudaEventRecord per launch open gaps between
    # dependent kernels); then the SAME K steps again with them on, for the per-kernel table and the rooflines.  Solves are
    # bit-reproducible, so both passes execute the same launches.
    st, st_b, done, solves, wall = timed_pass(False)
    st_k, st_b, done_k, _, _ = timed_pass(True)
    st["kernels"] = st_k["kernels"]
    st["ms_kernels_total"] = st_k["ms_kernels_total"]
    st["ms_run_instrumented"] = st_k["ms_run"]
    st["deflated_solves"] = st_k["deflated_solves"]
    # single GPU: bit-reproducible, both passes are the same launches.  Sharded: the order of the NCCL sums may differ from run to run,
    # an LM trajectory may then take one linear solve more or less; the line says so instead of failing
    passes_identical = bool(done_k == done and st_k["launches_total"] == st["launches_total"])
    # (the clock sampler keeps running through the e2e and reloc legs: the BA region alone lasts ~0.3 s)
    dev_ms = allmax(st["ms_run"])  # CUDA events on the solver's stream, max over ranks
    dev_ms_instrumented = allmax(st["ms_run_instrumented"])  # (collectives are called by EVERY rank, never inside `if rank == 0`)
    ms_per_step = dev_ms / max(done, 1)
    value = M_total * done / (dev_ms * 1e-3) / 1e6
    kernels = st["kernels"]
    # ---- roofline: algorithmic bytes per launch / mean event-timed duration, against the measured HBM peak
    peak, peak_src = load_peaks()
    ncl = 4 + (args.factor_type > 0) + (args.factor_type == 2)
    RS, WS = 8 + 2 * ncl, (3 * ncl + 3) & ~3  # record and What sizes in doubles (Dims<NCL> of ba_kernels.cuh)
    nnzb, npairs = st_b["nnz_blocks"], st_b["num_pairs"]
    cg_its_per_launch = st["pcg_iterations"] / max(kernels.get("pcg", dict(launches=1))["launches"], 1)
    # after an accepted step the ray blocks are damped and factored inside k_track_accum (160 B/track: Vh in registers, Lt + diagonal out),
    # k_track_factor then only runs for the solves that follow a rejected step: split its bytes between the two entries accordingly
    n_acc = kernels.get("track_accum", dict(launches=0))["launches"]
    n_sol = max(kernels.get("track_solve", dict(launches=1))["launches"], 1)
    f_fused = min(1.0, n_acc / n_sol)
    alg = {
        # stage 1: 16 B read + r 16 B + J
        "resjac": ("hbm", RJ_BYTES_PER_OBS[args.factor_type] * prob.M, "k_resjac: 16 B in + 16 B r + %d B J per observation" % (RJ_BYTES_PER_OBS[args.factor_type] - 32)),
        "track_accum": ("hbm", 64 * prob.M + (80 + 104) * prob.P, "k_track_accum: r + E (64 B) per observation gathered by track, 80 B per track out, + the damped Cholesky factor of the next solve (104 B per track)"),
        # stage 2: record in, What out; Cholesky factor per track
        "track_solve": ("hbm", 8 * (RS + WS) * prob.M + 160 * prob.P * (1.0 - f_fused),
                        "k_obs_what (+ k_track_factor after rejected steps, %.0f %% of the solves): record %d B in + What %d B out per observation, 160 B per track when the factor runs" % (100 * (1 - f_fused), 8 * RS, 8 * WS)),
        "schur_offdiag": ("l2", 2 * 8 * WS * npairs, "k_schur_offdiag: two What records (%d B) per observation pair, L2-resident gathers" % (8 * WS)),
        # stage 3: per CG iteration every block of S~ (NCL^2 doubles) and the (r, w, s) state of its column (3 NCL doubles)
        "pcg": ("l2-latency + grid barrier", nnzb * (ncl * ncl + 3 * ncl) * 8 * cg_its_per_launch,
                "k_cg: (NCL^2 + 3 NCL) x 8 B per block of S~ per CG iteration x %.1f iterations per launch; S~ lives in shared memory, the state in L2: "
                "latency-bound, not bandwidth-bound" % cg_its_per_launch),
        # stage 4
        "track_backsub": ("hbm", 8 * WS * prob.M + 240 * prob.P, "k_track_backsub: What %d B per observation gathered by track, 240 B per track" % (8 * WS)),
        "cost": ("hbm", 16 * prob.M, "k_cost: 16 B per observation (+ the L2-resident track gather)"),
    }
    prefix = {"resjac": "k_resjac", "track_accum": "k_track_accum", "track_solve": "k_obs_what", "schur_offdiag": "k_schur_offdiag", "pcg": "k_cg",
              "track_backsub": "k_track_backsub", "cost": "k_cost"}
    table = {}
    for name, k in kernels.items():
        avg_us = 1e3 * k["ms"] / k["launches"]
        row = dict(stage=k["stage"], launches=k["launches"], avg_us=round(avg_us, 2), share=round(k["ms"] / max(st["ms_kernels_total"], 1e-9), 4))
        if name in alg:
            row["alg_gbs"] = round(alg[name][1] / (avg_us * 1e-6) / 1e9, 1)
        table[name] = row
    dom = max(table, key=lambda n: table[n]["share"]) if table else None
    rj = table.get("resjac")
    launches = st["launches_total"]
    h.close()
    h = None

    counters = None
    if rank == 0 and world == 1 and not args.no_ncu:
        counters = hardware_counters(args)

    def roof_entry(name):
        row = table[name]
        bound, nbytes, what = alg[name]
        c = counters_for(counters, prefix[name])
        traffic = None
        if c is not None:
            traffic = c["dram_bytes"]
            if name == "track_solve":  # two kernels share the id: add the small one
                c2 = counters_for(counters, "k_track_factor")
                traffic += c2["dram_bytes"] if c2 else 0
        return dict(kernel=name, what=what, bound=bound if bound != "l2" else "l2", achieved=row["alg_gbs"], peak=peak, unit="GB/s",
                    frac=round(row["alg_gbs"] / peak, 4), traffic=traffic, share_of_step=row["share"], avg_us=row["avg_us"],
                    algorithmic_bytes_per_launch=int(nbytes))

    roofs = [roof_entry(n) for n in sorted((n for n in table if n in alg), key=lambda n: -table[n]["share"])]
    roofline = None
    if dom in alg:
        roofline = dict(roof_entry(dom))
        # the contract's enumeration is hbm | tensor: the fraction is taken against the HBM peak; `limiter` says what really bounds it
        roofline["limiter"] = roofline["bound"]
        roofline["bound"] = "hbm"
        roofline["peak_source"] = peak_src
        roofline["traffic_source"] = ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, side run of this bench (first 3 LM iterations of the same scene)"
                                      if roofline["traffic"] is not None else None)
        if dom == "pcg":
            roofline["us_per_cg_iteration"] = round(1e3 * kernels["pcg"]["ms"] / max(st["pcg_iterations"], 1), 3)
            roofline["cg_iterations_per_lm_step"] = round(st["pcg_iterations"] / st["lm_iterations"], 1)
            roofline["deflated_solves"] = st["deflated_solves"]
            roofline["deflation_vectors"] = st_b["deflation_vectors"]


    # ---------------- end to end through the C ABI with host buffers: complete ptzba_solve calls ----------------
    e2e = None
    if not args.no_e2e:
        # the contract's e2e leg copies from PINNED host memory: re-home the input arrays (same values)
        def pin(a):
            return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        for name in ("intr", "ext", "obs_uv", "obs_view", "obs_track", "track_weight"):
            setattr(prob, name, pin(getattr(prob, name)))
        outs = {}

        def pinned_out(name, shape):  # caller-owned result buffers: pinned once, reused by every solve (as a host application would)
            if name not in outs:
                outs[name] = torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()
            return outs[name]

        ptz.ba_solve(prob, opt, alloc=pinned_out)  # warm the memory pool / first-call costs outside the timed region
        barrier()
        n_e2e = max(1, args.e2e_solves)
        t0 = time.perf_counter()
        its = 0
        per_solve = []
        for _ in range(n_e2e):
            t1 = time.perf_counter()
            r = ptz.ba_solve(prob, opt, alloc=pinned_out)
            per_solve.append(round(1e3 * (time.perf_counter() - t1), 2))
            its += r.num_iterations
        barrier()
        dt = allmax(time.perf_counter() - t0)
        h2d = prob.M * (8 + 4 + 4) + prob.P * 8 + prob.V * 15 * 8
        d2h = prob.V * (15 + 21) * 8 + prob.P * 6 * 8
        e2e = dict(value=round(M_total * its / dt / 1e6, 3), unit="Mobs/s", h2d_bytes_per_step=int(h2d * n_e2e / max(its, 1)),
                   d2h_bytes_per_step=int(d2h * n_e2e / max(its, 1)), solves=n_e2e, lm_iterations=its, seconds=round(dt, 4),
                   seconds_setup_per_solve=round(r.seconds_setup, 4), seconds_lm_per_solve=round(r.seconds_solve, 4), ms_per_solve=per_solve)

    # ---------------- batched relocalisation (cfg 3), kernel-only with device-resident inputs and end to end ----------------
    reloc = None
    if not args.no_reloc:
        reloc = bench_reloc(args, rank, world, barrier, allmax, allsum, counters)

    tracks = None
    if rank == 0 and world == 1 and not args.no_tracks:
        tracks = bench_tracks(args, prob)

    small = None
    if rank == 0 and world == 1 and not args.no_small:
        small = bench_small(args)

    clocks = sampler.stop() if rank == 0 else None

    # ---------------- CPU baseline (oracle port) on rank 0, bounded sample ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, prob)

    if rank == 0:
        line = {
            "metric": "ptzba_lm_mobs_per_sec", "value": round(value, 3), "unit": "Mobs/s", "n_gpus": world, "steps": done, "warmup": W,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong" if args.config == "cfg5" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, prob, world, M_total),
            "lm_iters_per_sec": round(done / (dev_ms * 1e-3), 2), "solves_in_timed_region": solves, "wall_seconds": round(wall, 4),
            "ms_per_step_with_per_kernel_events": round(dev_ms_instrumented / max(done_k, 1), 4),
            "ms_per_step_kernels_only": round(st["ms_kernels_total"] / max(done_k, 1), 4), "timed_and_instrumented_pass_identical": passes_identical,
            "pcg_iterations_per_step": round(st["pcg_iterations"] / st["lm_iterations"], 1),
            "us_per_pcg_iteration": round(1e3 * kernels["pcg"]["ms"] / max(st["pcg_iterations"], 1), 3) if "pcg" in kernels else None,
            "rj_mobs_per_sec": round(prob.M / (rj["avg_us"]) , 1) if rj else None,
            "gpu_launches": launches, "kernels": table, "roofline": roofline, "roofline_kernels": roofs, "e2e": e2e, "small_configs": small, "reloc": reloc,
            "tracks": tracks, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if dist is not None:
        ptz.nccl_finalize()
        dist.destroy_process_group()


def bench_reloc(args, rank, world, barrier, allmax, allsum, counters=None):
    import ctypes as C

    import torch

    import ptz_calib_b200 as ptz
    from ptz_calib_b200 import abi, lib, synth

    B = args.reloc_queries
    full = synth.make_reloc_batch(B, factor_type=abi.PTZ_KRT_F)
    b = full.shard(rank, world)
    opt = ptz.default_options()
    # device-resident
    dev = torch.device("cuda")
    t_off = torch.from_numpy(b.match_offset).to(dev)
    t_ur, t_uc = torch.from_numpy(b.uv_ref).to(dev), torch.from_numpy(b.uv_cur).to(dev)
    t_ref, t_init = torch.from_numpy(b.ref_cam).to(dev), torch.from_numpy(b.init_cam).to(dev)
    o_cam = torch.zeros(b.B, 21, dtype=torch.float64, device=dev)
    o_i = [torch.zeros(b.B, dtype=torch.int32, device=dev) for _ in range(4)]
    o_d = [torch.zeros(b.B, dtype=torch.float64, device=dev) for _ in range(3)]
    o_loc = torch.zeros(b.B, 15, dtype=torch.float64, device=dev)
    cb = abi.RelocBatchC()
    cb.factor_type, cb.num_queries, cb.max_iter, cb.max_reproj_error = b.factor_type, b.B, b.max_iter, b.max_reproj_error
    cb.match_offset = C.cast(t_off.data_ptr(), abi.lp)
    cb.uv_ref, cb.uv_cur = C.cast(t_ur.data_ptr(), abi.fp), C.cast(t_uc.data_ptr(), abi.fp)
    cb.ref_cam, cb.init_cam = C.cast(t_ref.data_ptr(), abi.dp), C.cast(t_init.data_ptr(), abi.dp)
    cr = abi.RelocResultC()
    cr.cam = C.cast(o_cam.data_ptr(), abi.dp)
    cr.success, cr.termination, cr.num_iter, cr.iterations = (C.cast(t.data_ptr(), abi.ip) for t in o_i)
    cr.initial_cost, cr.final_cost, cr.final_rms = (C.cast(t.data_ptr(), abi.dp) for t in o_d)
    cr.local_cam15 = C.cast(o_loc.data_ptr(), abi.dp)
    L = lib.load()
    stream = torch.cuda.current_stream()
    for _ in range(3):
        lib.check(L.ptzreloc_solve_batch_dev(C.byref(cb), C.byref(opt), C.byref(cr), C.c_void_p(stream.cuda_stream)), "reloc dev")
    barrier()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        lib.check(L.ptzreloc_solve_batch_dev(C.byref(cb), C.byref(opt), C.byref(cr), C.c_void_p(stream.cuda_stream)), "reloc dev")
    e1.record(stream)
    barrier()
    ms = allmax(e0.elapsed_time(e1) / reps)
    iters = float(o_i[3].double().mean().item())
    succ = allsum(float(o_i[0].sum().item())) / B
    # end to end (pinned host buffers in, host results out)
    for name in ("match_offset", "uv_ref", "uv_cur", "ref_cam", "init_cam"):
        setattr(b, name, torch.from_numpy(np.ascontiguousarray(getattr(b, name))).pin_memory().numpy())
    ptz.reloc_solve_batch(b, opt)  # warm-up
    barrier()
    t0 = time.perf_counter()
    r = ptz.reloc_solve_batch(b, opt)
    barrier()
    dt = allmax(time.perf_counter() - t0)
    out = dict(workload=f"cfg3: {B} queries, {int(allsum(float(b.N)))} matches, factor F, 5% displaced outliers", solves_per_sec=round(B / (ms * 1e-3), 1),
               ms_per_batch=round(ms, 3), mean_lm_iterations=round(iters, 2), success_rate=round(succ, 4), e2e_solves_per_sec=round(B / dt, 1),
               e2e_seconds=round(dt, 4), h2d_bytes=int(b.N * 16 + b.B * (42 * 8 + 8)), d2h_bytes=int(b.B * (21 + 15 + 3) * 8 + b.B * 16))
    # roofline of k_reloc: bound by the fp64 pipe, not by HBM (16 B per match read once).  fp64 operations per query come
    # from the instruction counters of this run's side capture (2 x DFMA + DADD + DMUL), the peak from a DFMA micro-kernel on this device.
    if rank == 0:
        gf = C.c_double(0)
        if L.ptz_measure_fp64_gflops(C.byref(gf)) == 0 and gf.value > 0:
            out["fp64_peak_gflops_measured"] = round(gf.value, 1)
        c = counters_for(counters, "k_reloc")
        if c and c.get("fp64_flops"):
            nq = min(args.reloc_queries, 20000)  # queries of the side run (same generator, a prefix-sized batch)
            fl_q = c["fp64_flops"] * c["launches"] / nq  # (the batch runs as several slices = launches: sum them)
            out["fp64_flops_per_query"] = round(fl_q, 1)
            out["fp64_flops_per_match_per_lm_iteration"] = round(fl_q / (b.N / max(b.B, 1)) / max(iters, 1e-9), 1)
            ach = fl_q * b.B / (ms * 1e-3) / 1e9
            out["roofline"] = dict(kernel="k_reloc<F>", bound="fp64 pipe", achieved=round(ach, 1), peak=out.get("fp64_peak_gflops_measured"), unit="GFLOP/s",
                                   frac=round(ach / gf.value, 4) if gf.value > 0 else None,
                                   hbm_gbs=round((b.N * 16 + b.B * (42 + 39) * 8) / (ms * 1e-3) / 1e9, 1),
                                   traffic=c["dram_bytes"] * c["launches"] * b.B // max(nq, 1))
    return out


def bench_small(args):
    """BASELINE cfg 1 (Synthetic-shaped, V=36, PTZRay; with its georeferencing stage) and cfg 2 (WorldCup14-shaped, V=60, PTZRayDist) at full
    size: latency-bound problems, reported as LM iterations per second -- device-resident (CUDA events over ptzba_run)
    and end to end (ptzba_solve with host buffers)."""
    import ptz_calib_b200 as ptz
    from ptz_calib_b200 import synth

    out = {}
    for name, p in (("cfg1", synth.make_config(1)), ("cfg1_georef", synth.make_config(1, num_pts3d=10)), ("cfg2", synth.make_config(2))):
        # these scenes converge in 2-3 iterations from the standard initial guess; a harder start (2 deg, 8 % focal error) gives the
        # loop something to do, so both are reported
        rows = {}
        for tag, q in (("standard_init", p), ("hard_init", synth.make_config(int(name[3]), rot_noise_deg=2.0, focal_noise=0.08, **({"num_pts3d": 10} if "georef" in name else {})))):
            h = ptz.BAHandle(q, max_num_iterations=200)
            h.set_stage_timing(False)
            h.run(200); h.reset()
            a = h.stage_times()
            reps, its, conv = 20, 0, True
            for _ in range(reps):
                r = h.run(200)
                its += r.num_iterations
                conv = conv and r.converged
                h.reset()
            b = h.stage_times()
            h.close()
            ms = b["ms_run"] - a["ms_run"]
            ptz.ba_solve(q, max_num_iterations=200)
            t0 = time.perf_counter()
            for _ in range(5):
                r = ptz.ba_solve(q, max_num_iterations=200)
            dt = (time.perf_counter() - t0) / 5
            rows[tag] = dict(lm_iterations_per_solve=round(its / reps, 1), lm_iters_per_sec=round(its / (ms * 1e-3), 1), us_per_lm_iteration=round(1e3 * ms / max(its, 1), 1),
                             ms_per_solve_device=round(ms / reps, 3), ms_per_solve_e2e=round(1e3 * dt, 3), converged=bool(conv),
                             kernel_launches_per_iteration=round((b["launches_total"] - a["launches_total"]) / max(its, 1), 1))
        out[name] = dict(views=p.V, tracks=p.P, observations=p.M, annotated_points=p.A, **rows)
    try:
        out["iba_flow"] = bench_iba()
    except Exception as e:  # noqa: BLE001  (no compiler on the box, ...: the line still goes out)
        out["iba_flow"] = dict(error=str(e)[:200])
    return out


def bench_iba():
    """The whole incremental flow (run_ptzba_synthetic.sh: PtzIncrementalOptimizer from unknown cameras -- seed pair, batched KRT
    registrations, a global BA whenever the model grew by 10 %) through the C++ adaptor, on cfg-1-shaped rings with exhaustive pairwise
    matches.  One process per size: an untimed run first (CUDA start-up), then the timed one."""
    import re
    import struct
    import tempfile

    from ptz_calib_b200 import lib, synth

    d = tempfile.mkdtemp()
    exe = os.path.join(d, "iba_check")
    so_dir = os.path.dirname(lib.SO_PATH)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "iba_check.cpp"), "-o", exe, "-L" + so_dir, "-lptzcalib_b200", "-ldl",
                    "-Wl,-rpath," + so_dir], check=True, capture_output=True)
    rows = []
    for scale in (1.0, 2.0):
        p = synth.make_config(1, scale=scale)
        cams = np.zeros((p.V, 21))
        cams[:, 0] = cams[:, 1] = p.gt["f"]
        cams[:, 2:4] = p.gt["c"]
        cams[:, 4:13] = p.gt["R"].reshape(p.V, 9)
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("4i", p.V, p.M, 100, 1))
            for a in (cams, p.obs_view, p.obs_track, p.obs_uv):
                f.write(np.ascontiguousarray(a).tobytes())
        r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600, env=dict(os.environ, IBA_WARMUP="1"))
        m = re.search(r"iba ok=(\d) registered=(\d+)/(\d+) global BAs=(\d+) reloc batches=(\d+) \((\d+) queries\) reproj=([0-9.]+) pairs=(\d+) matches=(\d+)\s+([0-9.]+) s", r.stdout)
        c = re.search(r"cold run[^:]*: ([0-9.]+) s", r.stdout)
        if not m:
            rows.append(dict(views=p.V, error=(r.stdout + r.stderr)[-200:]))
            continue
        rows.append(dict(views=p.V, matches=int(m.group(9)), image_pairs=int(m.group(8)), ok=bool(int(m.group(1))), registered=int(m.group(2)),
                         global_bundle_adjustments=int(m.group(4)), registration_batches=int(m.group(5)), krt_solves=int(m.group(6)),
                         final_reproj_error_px=float(m.group(7)), seconds=float(m.group(10)), seconds_cold_process=float(c.group(1)) if c else None))
    return rows


def bench_tracks(args, prob):
    """Track building on the matches of the bench scene: device-resident (ptztracks_build_dev, CUDA events), end to
    end through ptztracks_build with pinned host buffers, and the CPU restatement of TracksBuilder on a bounded sample of the pairs."""
    import ctypes as C

    import torch

    import ptz_calib_b200 as ptz
    from ptz_calib_b200 import abi, lib, synth
    from ptz_calib_b200 import tracks as T

    m, v, _ = synth.make_matches_from_scene(prob)
    N = m.num_matches
    dev = torch.device("cuda")
    L = lib.load()
    tens = [torch.from_numpy(a).to(dev) for a in (m.pair_src, m.pair_dst, m.match_offset, m.query_idx, m.train_idx)]
    cm = T.MatchesC()
    cm.num_pairs, cm.min_track_length = len(m.pair_src), 4
    cm.pair_src, cm.pair_dst = C.cast(tens[0].data_ptr(), abi.ip), C.cast(tens[1].data_ptr(), abi.ip)
    cm.match_offset = C.cast(tens[2].data_ptr(), abi.lp)
    cm.query_idx, cm.train_idx = C.cast(tens[3].data_ptr(), abi.ip), C.cast(tens[4].data_ptr(), abi.ip)
    o_tid, o_off = torch.zeros(N, dtype=torch.int32, device=dev), torch.zeros(N + 1, dtype=torch.int64, device=dev)
    o_img, o_feat = torch.zeros(2 * N, dtype=torch.int32, device=dev), torch.zeros(2 * N, dtype=torch.int32, device=dev)
    cr = T.TracksC()
    cr.cap_tracks, cr.cap_elems = N, 2 * N
    cr.track_id, cr.track_offset = C.cast(o_tid.data_ptr(), abi.ip), C.cast(o_off.data_ptr(), abi.lp)
    cr.elem_img, cr.elem_feat = C.cast(o_img.data_ptr(), abi.ip), C.cast(o_feat.data_ptr(), abi.ip)
    stream = torch.cuda.Stream()  # a real stream: the library's scratch is pooled per stream (the legacy default stream gets plain cudaMalloc)
    torch.cuda.synchronize()

    def run():
        lib.check(L.ptztracks_build_dev(C.byref(cm), C.c_int64(N), C.byref(cr), C.c_void_p(stream.cuda_stream)), "tracks dev")

    for _ in range(3):
        run()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ok = cr.num_tracks == prob.P and cr.num_elems == prob.M
    # end to end with pinned host buffers
    for name in ("pair_src", "pair_dst", "match_offset", "query_idx", "train_idx"):
        setattr(m, name, torch.from_numpy(np.ascontiguousarray(getattr(m, name))).pin_memory().numpy())
    ptz.build_tracks(m, 4)
    t0 = time.perf_counter()
    t = ptz.build_tracks(m, 4)
    dt = time.perf_counter() - t0
    # CPU restatement (std::set + union-by-rank, as the reference) on the first pairs holding ~1/8 of the matches
    from oracle import oracle as orc

    kp = int(np.searchsorted(m.match_offset, N // 8))
    ms_ = T.Matches(m.pair_src[:kp], m.pair_dst[:kp], m.match_offset[:kp + 1], m.query_idx[:m.match_offset[kp]], m.train_idx[:m.match_offset[kp]])
    t0 = time.perf_counter()
    orc.tracks_build(ms_, 4)
    dc = time.perf_counter() - t0
    # algorithmic bytes: 16 B/match in (two indices + the pair's images), 8 B/track element + 12 B/track out; the build is a chain of
    # radix sorts and scans over 24-32 B per match, so its real DRAM traffic is ~10x that: the fraction says so
    peak, peak_src = load_peaks()
    alg = 16 * N + 8 * len(t.elem_img) + 12 * t.num_tracks
    roof = dict(kernel="ptztracks_build_dev (28 launches: CUB radix sorts / scans + union-find)", bound="hbm", achieved=round(alg / (ms * 1e-3) / 1e9, 1), peak=peak,
                unit="GB/s", frac=round(alg / (ms * 1e-3) / 1e9 / peak, 4), algorithmic_bytes_per_launch=int(alg), peak_source=peak_src, traffic=None)
    return dict(roofline=roof, workload=f"matches of the bench scene: {N} matches in {len(m.pair_src)} image pairs -> {t.num_tracks} tracks, {len(t.elem_img)} elements",
                matches_per_sec=round(N / (ms * 1e-3), 1), ms_per_build=round(ms, 3), tracks_as_scene=bool(ok),
                e2e_matches_per_sec=round(N / dt, 1), e2e_seconds=round(dt, 4), h2d_bytes=int(8 * N + 16 * len(m.pair_src)), d2h_bytes=int(8 * len(t.elem_img) + 12 * t.num_tracks),
                cpu_matches_per_sec=round(ms_.num_matches / dc, 1), cpu_sample=f"first {kp} pairs ({ms_.num_matches} matches), 1 thread, {dc:.1f} s")


def config_of(args, prob, world, M_total):
    """`config` of the JSON line: the workload both arms are quoted on (prob = rank 0's part of the scene)."""
    return {"workload": f"{'cfg5_sharded_ba' if args.config == 'cfg5' else 'cfg4_scaled_ba'} per GPU (V={prob.V}, P={prob.P} local tracks, M={prob.M} local obs, M_total={int(M_total)}); "
                        f"factor {['PTZRay', 'PTZRayDist', 'PTZRayFxfyDist'][args.factor_type]}; LM iterations from complete solves at Ceres-default "
                        f"tolerances, PCG tol {args.pcg_tol:g}; inputs larger than L2 (records {prob.M * 128 / 1e6:.0f} MB)",
            "views": prob.V, "obs_per_gpu": prob.M, "parallelism": (f"tracks/observations sharded x{world} (NCCL all-reduce of camera blocks), rows of the reduced system sharded x{world} "
                                                                     "(in-kernel NVLink peer exchange)") if world > 1 else "single GPU"}


def cpu_baseline(args, prob):
    """The oracle port timed on the box's host cores: a bounded sample of the same workload (first tracks of the scene), all host threads
    (passed explicitly: torchrun exports OMP_NUM_THREADS=1), plus its thread scaling and the Jacobian-only rate on a smaller sample."""
    from oracle import oracle as orc

    import ptz_calib_b200 as ptz

    threads = host_threads()

    def sample_of(T):
        sel = prob.obs_track < T
        return ptz.BAProblem(prob.factor_type, prob.intr, prob.ext, prob.obs_uv[sel], prob.obs_view[sel], prob.obs_track[sel], prob.track_weight[:T])

    kw = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, jacobian_mode=1, linear_solver=1, pcg_rel_tolerance=args.pcg_tol)
    T = min(prob.P, 10 * args.cpu_tracks)  # at the default the WHOLE cfg-4 scene (all tracks): ~5 s per run on 16 host cores, ~15 s with the scaling legs
    sample = sample_of(T)
    iters = args.cpu_iters
    t0 = time.perf_counter()
    rc, r = orc.ba_solve(sample, max_num_iterations=iters, num_threads=threads, **kw)
    dt = time.perf_counter() - t0
    its = max(r.num_iterations, 1)
    # thread scaling on a tenth of that sample: 1 thread vs all
    small = sample_of(max(1000, T // 10))
    t0 = time.perf_counter()
    orc.ba_solve(small, max_num_iterations=2, num_threads=1, **kw)
    d1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.ba_solve(small, max_num_iterations=2, num_threads=threads, **kw)
    dn = time.perf_counter() - t0
    tj = orc.ba_time_jacobian(small, 1, threads, 2)  # seconds per numeric-diff (Ceres CENTRAL) Jacobian evaluation
    ta = orc.ba_time_jacobian(small, 0, threads, 2)  # ... and per exact (analytic-equivalent) one
    return dict(value=round(sample.M * its / dt / 1e6, 4), unit="Mobs/s", cores=threads, kind="port",
                sample=f"{'all' if T == prob.P else 'first'} {T} tracks ({sample.M} obs) of the cfg-4 scene, all {prob.V} views, {its} LM iterations with Ceres-CENTRAL numeric "
                       f"Jacobians (as the reference), block-sparse Schur + block-Jacobi PCG; {dt:.1f} s", lm_iters_per_sec=round(its / dt, 4),
                thread_scaling=dict(threads=threads, speedup=round(d1 / dn, 2), sample_obs=small.M, seconds_1_thread=round(d1, 2), seconds_all_threads=round(dn, 2)),
                jacobian_only_mobs_per_sec=dict(ceres_central_numeric=round(small.M / tj / 1e6, 3), exact=round(small.M / ta / 1e6, 3), threads=threads))


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own algorithm for the path (Ceres-CENTRAL numeric Jacobians, LM, Schur) as
    restated by the oracle port, on this box's host cores.  Ceres/OpenCV/the reference binary cannot be built offline
    (SURVEY.md §8c), so kind = "port".  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import oracle as orc

    import ptz_calib_b200 as ptz

    threads = host_threads()  # explicit: torchrun exports OMP_NUM_THREADS=1, which must not turn the arm into a 1-thread run
    # the workload of our arm at this N (rank 0's part of the scene; the other ranks' parts only for the observation total of `config`)
    prob = make_scene(args, 0, world)
    M_total = prob.M + sum(make_scene(args, r, world).M for r in range(1, world))
    T = min(prob.P, args.cpu_tracks)
    sel = prob.obs_track < T
    sample = ptz.BAProblem(prob.factor_type, prob.intr, prob.ext, prob.obs_uv[sel], prob.obs_view[sel], prob.obs_track[sel], prob.track_weight[:T])
    K, W = args.steps, args.warmup
    # as our arm: LM iterations drawn from complete solves at Ceres-default tolerances (a solve that ends early is followed by the next)
    kw = dict(jacobian_mode=1, linear_solver=1, num_threads=threads, pcg_rel_tolerance=args.pcg_tol)

    def run_steps(n):
        done, dt = 0, 0.0
        while done < n:
            t0 = time.perf_counter()
            rc, r = orc.ba_solve(sample, max_num_iterations=n - done, **kw)
            dt += time.perf_counter() - t0
            if rc != 0 or r.num_iterations <= 0:
                break
            done += r.num_iterations
        return done, dt

    if W > 0:
        run_steps(min(W, 2))  # page the sample and the thread pool in; each LM iteration of the port costs 0.1-0.2 s
    its, dt = run_steps(K)
    its = max(its, 1)
    v = sample.M * its / dt / 1e6
    desc = (f"first {T} tracks ({sample.M} obs) of rank 0's part of the scene (V={prob.V}); {its} LM iterations, Ceres-CENTRAL numeric Jacobians, "
            f"block-sparse Schur + PCG, {threads} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": "ptzba_lm_mobs_per_sec", "value": round(v, 4), "unit": "Mobs/s", "n_gpus": world, "steps": its, "warmup": W,
        "ms_per_step": round(1e3 * dt / its, 3), "higher_is_better": True, "scaling": "strong" if args.config == "cfg5" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_of(args, prob, world, M_total),  # our arm's config; each step of THIS arm is one LM iteration on the bounded sample below
        "cpu_baseline": {"value": round(v, 4), "unit": "Mobs/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": round(v, 4), "unit": "Mobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lm_iters_per_sec": round(its / dt, 4), "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink cfg 4 (tests / smoke only; the default is the named config)")
    ap.add_argument("--factor-type", type=int, default=0)
    ap.add_argument("--pcg-tol", type=float, default=1e-13)
    ap.add_argument("--reloc-queries", type=int, default=100000)
    ap.add_argument("--e2e-solves", type=int, default=3)
    ap.add_argument("--cpu-tracks", type=int, default=40000)
    ap.add_argument("--cpu-iters", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tracks", action="store_true")
    ap.add_argument("--no-reloc", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="skip the hardware-counter side run (roofline.traffic becomes null)")
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"], help="cfg4: 2M observations per GPU (weak scaling); cfg5: 5e7 observations in total")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
