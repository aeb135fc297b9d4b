#!/usr/bin/env python
"""bench.py — headline benchmark of the PTZ-Calib NLS hot path on B200 (contract in the task description, ④).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference-equivalent CPU path (oracle port), host cores

Workload (config.workload): BASELINE cfg 4 "scaled synthetic PTZ-BA: 1,000 views x 2M observations, full LM with
Schur + PCG" on one GPU.  With N > 1 the scene grows with N (V = 1000 N views, ~2M observations per rank, tracks
sharded by observation, NCCL all-reduce of the camera blocks): weak scaling, cfg-5-shaped.
A STEP is one Levenberg-Marquardt iteration (stages 1-4).  Steps are drawn from complete solves run with the
reference's own tolerances: when a solve converges the problem is reset and solved again until K steps are done.
metric = observations x LM iterations per second (Mobs/s); lm_iters_per_sec and the batched-reloc figures ride along.
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


def ncu_traffic(factor_type, M):
    """DRAM bytes per launch of k_resjac from the committed ncu --set full capture (profiles/r1_dram_traffic.json), scaled by
    the observation count when the launch differs from the profiled one; None when no capture covers this factor type."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_dram_traffic.json")) as f:
            rec = json.load(f).get(f"k_resjac<{factor_type}>")
        return None if rec is None else int(rec["dram_bytes_per_launch"] * (M / rec["obs"]))
    except Exception:
        return None


def make_scene(args, rank, world):
    from ptz_calib_b200 import synth

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
    barrier()
    st_a = h.stage_times()
    t0 = time.perf_counter()
    done, solves, last = run_steps(K)
    st_b = h.stage_times()
    barrier()
    wall = time.perf_counter() - t0
    # (the clock sampler keeps running through the e2e and reloc legs: the BA region alone lasts ~0.1 s)
    # timings accumulate over the handle's life: the timed region is the difference
    st = dict(ms_run=st_b["ms_run"] - st_a["ms_run"], pcg_iterations=st_b["pcg_iterations"] - st_a["pcg_iterations"],
              lm_iterations=max(st_b["lm_iterations"] - st_a["lm_iterations"], 1), kernels={})
    for name, kb in st_b["kernels"].items():
        ka = st_a["kernels"].get(name, dict(ms=0.0, launches=0))
        if kb["launches"] > ka["launches"]:
            st["kernels"][name] = dict(ms=kb["ms"] - ka["ms"], launches=kb["launches"] - ka["launches"], stage=kb["stage"])
    st["launches_total"] = sum(k["launches"] for k in st["kernels"].values())
    st["ms_kernels_total"] = sum(k["ms"] for k in st["kernels"].values())
    dev_ms = allmax(st["ms_run"])  # CUDA events on the solver's stream, max over ranks
    ms_per_step = dev_ms / max(done, 1)
    value = M_total * done / (dev_ms * 1e-3) / 1e6
    kernels = st["kernels"]
    # roofline of the dominant kernel (largest share of the step) and of the streaming residual+Jacobian kernel
    peak, peak_src = load_peaks()
    alg_bytes = {
        "resjac": RJ_BYTES_PER_OBS[args.factor_type] * prob.M,
        "cost": 16 * prob.M,
        "track_accum": 64 * prob.M,
        "track_solve": (8 * (8 + 2 * (4 + (args.factor_type > 0) + (args.factor_type == 2))) + 8 * 3 * 4 + 8 * 4) * prob.M,
    }
    table = {}
    for name, k in kernels.items():
        avg_us = 1e3 * k["ms"] / k["launches"]
        row = dict(stage=k["stage"], launches=k["launches"], avg_us=round(avg_us, 2), share=round(k["ms"] / max(st["ms_kernels_total"], 1e-9), 4))
        if name in alg_bytes:
            row["alg_gbs"] = round(alg_bytes[name] / (avg_us * 1e-6) / 1e9, 1)
        table[name] = row
    dom = max(table, key=lambda n: table[n]["share"]) if table else None
    rj = table.get("resjac")
    roof_kernel = "resjac"
    roofline = None
    if rj:
        roofline = dict(kernel="k_resjac (stage 1: residual + analytic Jacobian)", bound="hbm", achieved=rj["alg_gbs"], peak=peak, unit="GB/s",
                        frac=round(rj["alg_gbs"] / peak, 4), traffic=ncu_traffic(args.factor_type, prob.M), peak_source=peak_src,
                        algorithmic_bytes_per_obs=RJ_BYTES_PER_OBS[args.factor_type], dominant_kernel_by_time=dom,
                        dominant_kernel_share=table[dom]["share"] if dom else None)
    launches = st["launches_total"]
    h.close()

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
        reloc = bench_reloc(args, rank, world, barrier, allmax, allsum)

    tracks = None
    if rank == 0 and world == 1 and not args.no_tracks:
        tracks = bench_tracks(args, prob)

    clocks = sampler.stop() if rank == 0 else None

    # ---------------- CPU baseline (oracle port) on rank 0, bounded sample ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, prob)

    if rank == 0:
        line = {
            "metric": "ptzba_lm_mobs_per_sec", "value": round(value, 3), "unit": "Mobs/s", "n_gpus": world, "steps": done, "warmup": W,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg4_scaled_ba per GPU (V={prob.V}, P={prob.P} local tracks, M={prob.M} local obs, M_total={int(M_total)}); "
                                   f"factor {['PTZRay', 'PTZRayDist', 'PTZRayFxfyDist'][args.factor_type]}; LM iterations from complete solves at Ceres-default "
                                   f"tolerances, PCG tol {args.pcg_tol:g}; inputs larger than L2 (records {prob.M * 128 / 1e6:.0f} MB)",
                       "views": prob.V, "obs_per_gpu": prob.M, "parallelism": (f"tracks/observations sharded x{world} (NCCL all-reduce of camera blocks), rows of the reduced system sharded x{world} "
                                                                                       "(in-kernel NVLink peer exchange)") if world > 1 else "single GPU"},
            "lm_iters_per_sec": round(done / (dev_ms * 1e-3), 2), "solves_in_timed_region": solves, "wall_seconds": round(wall, 4),
            "pcg_iterations_per_step": round(st["pcg_iterations"] / st["lm_iterations"], 1),
            "us_per_pcg_iteration": round(1e3 * kernels["pcg"]["ms"] / max(st["pcg_iterations"], 1), 3) if "pcg" in kernels else None,
            "rj_mobs_per_sec": round(prob.M / (rj["avg_us"]) , 1) if rj else None,
            "gpu_launches": launches, "kernels": table, "roofline": roofline, "e2e": e2e, "reloc": reloc, "tracks": tracks, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if dist is not None:
        ptz.nccl_finalize()
        dist.destroy_process_group()


def bench_reloc(args, rank, world, barrier, allmax, allsum):
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
    return dict(workload=f"cfg3: {B} queries, {int(allsum(float(b.N)))} matches, factor F, 5% displaced outliers", solves_per_sec=round(B / (ms * 1e-3), 1),
                ms_per_batch=round(ms, 3), mean_lm_iterations=round(iters, 2), success_rate=round(succ, 4), e2e_solves_per_sec=round(B / dt, 1),
                e2e_seconds=round(dt, 4), h2d_bytes=int(b.N * 16 + b.B * (42 * 8 + 8)), d2h_bytes=int(b.B * (21 + 15 + 3) * 8 + b.B * 16))


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
    # algorithmic bytes: 16 B/match in (two indices + the pair's images), 8 B/track element out
    return dict(workload=f"matches of the bench scene: {N} matches in {len(m.pair_src)} image pairs -> {t.num_tracks} tracks, {len(t.elem_img)} elements",
                matches_per_sec=round(N / (ms * 1e-3), 1), ms_per_build=round(ms, 3), tracks_as_scene=bool(ok),
                e2e_matches_per_sec=round(N / dt, 1), e2e_seconds=round(dt, 4), h2d_bytes=int(8 * N + 16 * len(m.pair_src)), d2h_bytes=int(8 * len(t.elem_img) + 12 * t.num_tracks),
                cpu_matches_per_sec=round(ms_.num_matches / dc, 1), cpu_sample=f"first {kp} pairs ({ms_.num_matches} matches), 1 thread, {dc:.1f} s")


def cpu_baseline(args, prob):
    """The oracle port timed on the box's host cores: a bounded sample of the same workload (first tracks of the scene)."""
    from oracle import oracle as orc

    threads = orc.num_threads()
    T = min(prob.P, 5 * args.cpu_tracks)  # ~10-20 s of CPU work
    sel = prob.obs_track < T
    import ptz_calib_b200 as ptz

    sample = ptz.BAProblem(prob.factor_type, prob.intr, prob.ext, prob.obs_uv[sel], prob.obs_view[sel], prob.obs_track[sel], prob.track_weight[:T])
    iters = args.cpu_iters
    t0 = time.perf_counter()
    rc, r = orc.ba_solve(sample, max_num_iterations=iters, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, jacobian_mode=1,
                         linear_solver=1, num_threads=threads, pcg_rel_tolerance=args.pcg_tol)
    dt = time.perf_counter() - t0
    its = max(r.num_iterations, 1)
    return dict(value=round(sample.M * its / dt / 1e6, 4), unit="Mobs/s", cores=threads, kind="port",
                sample=f"first {T} tracks ({sample.M} obs) of the cfg-4 scene, all {prob.V} views, {its} LM iterations with Ceres-CENTRAL numeric "
                       f"Jacobians (as the reference), block-sparse Schur + block-Jacobi PCG; {dt:.1f} s", lm_iters_per_sec=round(its / dt, 4))


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

    threads = orc.num_threads()
    prob = make_scene(args, 0, 1)
    T = min(prob.P, args.cpu_tracks)
    sel = prob.obs_track < T
    sample = ptz.BAProblem(prob.factor_type, prob.intr, prob.ext, prob.obs_uv[sel], prob.obs_view[sel], prob.obs_track[sel], prob.track_weight[:T])
    K, W = args.steps, args.warmup
    kw = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, jacobian_mode=1, linear_solver=1, num_threads=threads,
              pcg_rel_tolerance=args.pcg_tol)
    if W > 0:
        orc.ba_solve(sample, max_num_iterations=min(W, 1), **kw)
    t0 = time.perf_counter()
    rc, r = orc.ba_solve(sample, max_num_iterations=K, **kw)
    dt = time.perf_counter() - t0
    its = max(r.num_iterations, 1)
    v = sample.M * its / dt / 1e6
    desc = (f"first {T} tracks ({sample.M} obs) of the cfg-4 scene (V={prob.V}); {its} LM iterations, Ceres-CENTRAL numeric Jacobians, "
            f"block-sparse Schur + PCG, {threads} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": "ptzba_lm_mobs_per_sec", "value": round(v, 4), "unit": "Mobs/s", "n_gpus": world, "steps": its, "warmup": W,
        "ms_per_step": round(1e3 * dt / its, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4_scaled_ba (bounded sample): " + desc},
        "cpu_baseline": {"value": round(v, 4), "unit": "Mobs/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": round(v, 4), "unit": "Mobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lm_iters_per_sec": round(its / dt, 4), "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
