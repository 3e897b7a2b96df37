#!/usr/bin/env python3
"""bench.py -- dStar hot path on B200: zone-microphysics evals/s and cooling-curves/s.

One "step" = one pass of the hot path over one batch of synthetic stars:
    seam A  coefficient pass (cell loop + face loop kernels)  -> 2 * nz * 128 evals per star
    seam B  thermal evolution of every star through its whole epoch list (one launch)
`value` times the step with all inputs resident in HBM (CUDA events around the library's own
kernel launches, summed; reported max over ranks); `e2e` times the same step through the C ABI
with HOST buffers: H2D of every input inside the timed region, D2H of the light-curve monitors.

Workload at N = 1: BASELINE.json configs[1], the custom_Tc_Qimp sweep (256 stars, (core Tc, Qimp)
grid, 9 epochs, hooks alt_Qimp/alt_sf, nz ~ 253).  N > 1: each rank owns its own 256-star grid
(weak scaling, independent stars, no data-path collective; only the final gather of monitors).

--impl reference: the reference's CPU implementation of the same path.  The Fortran+MESA original
cannot be built here, so this is the CPU oracle (kind "port"), on all host threads, on a bounded
sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "zone_microphysics_evals_per_sec"
UNIT = "evals/s"
# algorithmic FP64 work per evaluation: cell = EOS + neutrino, face = EOS + conductivity,
# one ion species, neutron channels off; weighted operation counts incl. transcendentals
FLOP_CELL, FLOP_FACE = 6.0e3, 1.6e3
ARRS_STRUCT = ["rho", "rho_bar", "P", "P_bar", "dm", "ionic", "ionic_bar", "Yion", "Yion_bar"]
ARRS_GEOM = ["dm_bar", "area", "ePhi", "e2Phi", "ePhi_bar", "e2Phi_bar"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="custom_Tc_Qimp")
    ap.add_argument("--stars", type=int, default=256, help="stars per GPU")
    ap.add_argument("--resolution", type=float, default=0.05)
    ap.add_argument("--cpu-sample-stars", type=int, default=0, help="stars in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md).

    NVML in-process (a few hundred samples per second, so even a 50 ms region is covered); falls back to an
    `nvidia-smi -lms` child process when the NVML binding is unavailable.
    """
    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            self.index = int(vis.split(",")[index]) if vis else index
        except (ValueError, IndexError):
            self.index = index
        self.sm, self.reasons, self.smax = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.rows = []
        self.source = None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in self._BITS.items():
                    if bits & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.sm.clear()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.source == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl" if args.impl == "b200" else "gloo")
    return rank, world, local, dist


def barrier(dist, local=0):
    if dist is not None:
        import torch
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()


def allmax(dist, x, local):
    if dist is None:
        return x
    import torch
    dev = "cuda:%d" % local if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def build_stars(ds, args, rank):
    from dstar_b200 import workloads
    fn = workloads.BUILDERS[args.workload]
    if args.workload == "basic_run":
        stars = fn(args.resolution)
    else:
        stars = fn(args.stars, args.resolution, offset=rank)
    ds.create_models(stars)
    return stars


def gather_host_inputs(stars):
    """Concatenate the stars' arrays into the flat SoA buffers the C ABI takes."""
    nz = np.array([s.get_int("nz") for s in stars], dtype=np.int32)
    zoff = np.concatenate([[0], np.cumsum(nz)]).astype(np.int32)
    cat = lambda name: np.concatenate([s.array(name) for s in stars])
    H = {k: cat(k) for k in ARRS_STRUCT + ARRS_GEOM + ["lnT"]}
    # the stars carry the fixed-up rho_bar(1); the coefficient pass starts from the undefined one (0)
    H["rho_bar"] = H["rho_bar"].copy()
    H["rho_bar"][zoff[:-1]] = 0.0
    H["Zion"] = stars[0].array("Zion"); H["Aion"] = stars[0].array("Aion")
    try:
        H["Tc_cell"] = cat("Tc_cell"); H["Tc_face"] = cat("Tc_face")
    except Exception:
        H["Tc_cell"] = None; H["Tc_face"] = None
    H["atm_lgTb"] = np.stack([s.array("atm_lgTb") for s in stars])
    H["atm_lgTeff"] = np.stack([s.array("atm_lgTeff") for s in stars])
    H["atm_lgflux"] = np.stack([s.array("atm_lgflux") for s in stars])
    H["eb"] = np.stack([s.array("epoch_boundaries") for s in stars], axis=1)   # (n_epoch+1, n_star)
    H["em"] = np.stack([s.array("epoch_Mdots") for s in stars], axis=1)
    return nz, zoff, H


def workload_ctrl(ds, args):
    mc = ds.MicroCtrl.defaults(); ec = ds.EvolveCtrl.defaults()
    heat = [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 3.0]
    T_atm = 1.0e8
    if args.workload == "custom_Tc_Qimp":
        ec.fix_atmosphere_temperature_when_accreting = 0; ec.turn_on_extra_heating = 1
        heat[6:9] = [27.0, 28.0, 1.7]; T_atm = 2.4e8
    elif args.workload == "basic_run":
        ec.turn_on_extra_heating = 1; heat[6:9] = [27.5, 28.5, 0.1]; T_atm = 2.4e8
    elif args.workload in ("fit_lightcurve", "int16_demo"):
        ec.fix_atmosphere_temperature_when_accreting = 0; ec.turn_on_extra_heating = 1
    elif args.workload == "accretion":
        mc.use_neutron_conductivity = 1; mc.use_superfluid_phonon_conductivity = 1; mc.turn_on_shell_Urca = 1
        ec.turn_on_extra_heating = 1; heat[6:9] = [27.0, 28.0, 1.0]; T_atm = 5.0e8
    return mc, ec, heat, T_atm


def per_star_heat(stars, heat, args):
    h = np.tile(np.array(heat), (len(stars), 1))
    return h


def run_b200(args):
    rank, world, local, dist = dist_setup(args)
    import dstar_b200 as ds
    ds.NScool_init(device=local)
    ctx = ds.Context(local, gaps=("ns", "sfb03" if args.workload != "accretion" else "gc", "t72"))
    stars = build_stars(ds, args, rank)
    nz, zoff, H = gather_host_inputs(stars)
    n_star, total = len(stars), int(zoff[-1])
    pinned = []
    try:   # host inputs in pinned memory (torch is only the allocator here)
        import torch
        for k_, v_ in list(H.items()):
            if isinstance(v_, np.ndarray) and v_.dtype == np.float64:
                t_ = torch.from_numpy(np.ascontiguousarray(v_)).pin_memory()
                pinned.append(t_)
                H[k_] = t_.numpy()
    except Exception:
        pass
    mc, ec, heat, T_atm = workload_ctrl(ds, args)
    heat_arr = per_star_heat(stars, heat, args)
    if args.workload == "fit_lightcurve":
        from dstar_b200 import workloads
        rng = np.random.default_rng(workloads.SEED + 2 + 1000 * rank)
        rng.uniform(np.log10(3e16), np.log10(3e17), n_star)
        heat_arr[:, 8] = rng.uniform(0.0, 3.0, n_star)
    T_atm_arr = np.full(n_star, T_atm)
    evals_cell = int(np.sum(nz)) * 128
    evals = 2 * evals_cell

    def upload_all(b):
        b.upload_structure(*[H[k] for k in ARRS_STRUCT], H["Zion"], H["Aion"], H["Tc_cell"], H["Tc_face"])
        b.upload_evolve(*[H[k] for k in ARRS_GEOM], H["atm_lgTb"], H["atm_lgTeff"], H["atm_lgflux"],
                        np.arange(n_star), heat_arr, T_atm_arr, H["lnT"], H["eb"], H["em"])

    # ---------------- resident measurement
    b = ds.Batch(ctx, zoff, stars[0].get_int("ncharged"))
    upload_all(b)
    h2d = b._h2d_structure + b._h2d_evolve

    def step_resident():
        b.reset_state()
        b.micro_tables(mc, 3)
        t_c, t_f = b.kernel_ms(0), b.kernel_ms(1)
        b.evolve(ec, 0)
        return t_c, t_f, b.kernel_ms(2)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier(dist, local)
    sampler.start()
    tc = tf = te = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a_, b_, c_ = step_resident()
        tc += a_; tf += b_; te += c_
    wall_res = time.perf_counter() - t0
    barrier(dist, local)
    clocks = sampler.stop()
    res = b.download_results(full=False)
    ms_step = (tc + tf + te) / args.steps        # device time of the step's kernels (CUDA events in the library)
    ms_step = allmax(dist, ms_step, local)
    # ---------------- end-to-end through the C ABI with host buffers
    # (the batch handle is reused between steps, as a sweep driver would: device buffers are allocated once,
    #  every input is copied host->device again on every step)
    bb = ds.Batch(ctx, zoff, stars[0].get_int("ncharged"))

    def step_e2e():
        upload_all(bb)
        bb.micro_tables(mc, 3)
        bb.evolve(ec, 0)
        return bb.download_results(full=False)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r2 = step_e2e()
    barrier(dist, local)
    e2e_s = allmax(dist, (time.perf_counter() - t0) / args.steps, local)
    d2h = sum(r2[k].nbytes for k in ("t_monitor", "Teff_monitor", "Qb_monitor", "idid", "counters", "Teff", "Lsurf",
                                     "dt", "tsec", "model"))
    # ---------------- roofline of the dominant kernel (coefficient pass), FP64 pipe
    fp64_peak = ctx.fp64_peak_tflops()
    # dominant kernel = the cell pass (ds_micro_kernel<cell>: EOS + neutrino emissivities, ~60 % of the step); its
    # duration is the library's CUDA-event time on the launching stream, averaged over the timed steps
    cell_ms = tc / args.steps
    micro_ms = (tc + tf) / args.steps
    achieved_tf = FLOP_CELL * evals_cell / (cell_ms * 1e-3) / 1e12
    both_tf = (FLOP_CELL + FLOP_FACE) * evals_cell / (micro_ms * 1e-3) / 1e12
    alg_bytes = 64.0 * evals_cell       # ~64 B / eval (32 B of spline coefficients per knot and quantity)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM bytes of one cell-pass launch from the committed ncu --set full capture of this very configuration
    # (profiles/r01d_ncu_full_summary.md, r01l table: dram__bytes_read.sum + dram__bytes_write.sum); other sizes: null
    traffic = 6.11e9 if (args.workload == "custom_Tc_Qimp" and n_star == 256 and abs(args.resolution - 0.05) < 1e-12) else None
    roofline = {"bound": "fp64", "kernel": "ds_micro_kernel<cell>", "achieved": achieved_tf, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                "traffic_note": "bytes per launch, ncu capture r01l (mostly register-spill frames cycling through L2); "
                                "algorithmic bytes per launch %.3g" % alg_bytes,
                "peak_source": "DFMA-chain probe measured in this run (MEASURED_PEAKS.json has no FP64 figure); the "
                               "reference's formulas may not be contracted into FMAs, so 0.5 is the ceiling of frac",
                "flops_per_eval": {"cell": FLOP_CELL, "face": FLOP_FACE},
                "cell_plus_face": {"achieved": both_tf, "frac": both_tf / fp64_peak if fp64_peak > 0 else None},
                "fp64_pipe_busy_ncu": {"cell": 0.445, "face": 0.44, "source": "profiles/r01d_ncu_full_summary.md"},
                "hbm": {"achieved_gbs": alg_bytes / (cell_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
                "kernel_ms": {"cells": tc / args.steps, "faces": tf / args.steps, "evolve": te / args.steps}}
    # ---------------- CPU baseline (oracle) on rank 0, N = 1
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(stars, H, zoff, args, threads=1, budget_s=15.0)
    ok = bool(np.all(res["idid"] >= 0))
    line = {"metric": METRIC, "value": evals * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d stars/GPU x nz~%d, %d epochs, %s crust, bc09 atmosphere (device-generated table), rebuilt HELM/gap tables"
                       % (args.workload, n_star, int(np.median(nz)), H["em"].shape[0],
                          "abuntime (synthetic)" if args.workload == "int16_demo" else "HZ90"),
                       "stars_per_gpu": n_star, "zones_total_per_gpu": total, "seed": 20261019,
                       "l2_policy": "inputs+tables per step %.0f MB > 126 MB L2" % ((4 * 4096.0 * total + h2d) / 1e6)},
            "curves_per_sec": n_star * world / (ms_step * 1e-3),
            "e2e": {"value": evals * world / e2e_s, "unit": UNIT, "curves_per_sec": n_star * world / e2e_s,
                    "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            # launches per step: zone pre-pass + cell pass (branch-free) + cell redo pass, zone pre-pass + face pass, evolve
            "gpu_launches": 6 * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "all_stars_converged": ok, "wall_s_resident_loop": wall_res}
    if dist is not None:
        # final gather of the light curves (the only collective of the job)
        from dstar_b200.dist import gather_monitors
        full = gather_monitors(dist, res["Teff_monitor"], n_star * world, rank, world, device="cuda:%d" % local)
        if rank == 0:
            line["gathered_monitor_shape"] = list(full.shape)
    if rank == 0:
        print(json.dumps(line))
    bb.close(); b.close(); ctx.close(); ds.NScool_shutdown()
    if dist is not None:
        dist.destroy_process_group()


def oracle_star_inputs(stars, H, zoff, i):
    s = stars[i]
    a, e = int(zoff[i]), int(zoff[i + 1])
    sl = lambda k, w=1: H[k][a * w:e * w]
    return dict(nz=e - a, ncharged=s.get_int("ncharged"), rho=sl("rho"), rho_bar=sl("rho_bar"), P=sl("P"),
                P_bar=sl("P_bar"), dm=sl("dm"), ionic=sl("ionic", 11), ionic_bar=sl("ionic_bar", 11),
                Yion=sl("Yion", s.get_int("ncharged")), Yion_bar=sl("Yion_bar", s.get_int("ncharged")),
                Zion=H["Zion"], Aion=H["Aion"],
                Tc_cell=None if H["Tc_cell"] is None else sl("Tc_cell", 3),
                Tc_face=None if H["Tc_face"] is None else sl("Tc_face", 3))


def cpu_baseline(stars, H, zoff, args, threads, budget_s):
    """Time the CPU oracle on a bounded sample of the same workload: `threads` stars in parallel, each
    doing its full coefficient pass and light curve."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from parity_helpers import load_oracle_gaps, oracle_evolve_star
    from concurrent.futures import ThreadPoolExecutor
    o = Oracle()
    load_oracle_gaps(o, "ns", "sfb03" if args.workload != "accretion" else "gc", "t72")
    import dstar_b200 as ds
    mc, ec, heat, T_atm = workload_ctrl(ds, args)
    cond = (mc.use_neutron_conductivity, mc.use_superfluid_phonon_conductivity, 1, 1, 10.0, 9.0)
    n = args.cpu_sample_stars or threads
    n = min(n, len(stars))

    def one(i):
        kw = oracle_star_inputs(stars, H, zoff, i)
        t0 = time.perf_counter()
        tabs = o.micro_tables(cond_ctrl=cond, shell_Urca=(mc.turn_on_shell_Urca, 1.0, 28.5), **kw)
        t1 = time.perf_counter()
        lnT0 = H["lnT"][zoff[i]:zoff[i + 1]]
        oracle_evolve_star(o, stars[i], lnT0, tabs=tabs, heat=heat, T_atm=T_atm, quirks=1,
                           ctrl=dict(turn_on_extra_heating=ec.turn_on_extra_heating,
                                     fix_atmosphere_temperature_when_accreting=ec.fix_atmosphere_temperature_when_accreting))
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, 2 * kw["nz"] * 128
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        rs = list(ex.map(one, range(n)))
    wall = time.perf_counter() - t0
    evals = sum(r[2] for r in rs)
    return {"value": evals / wall, "unit": UNIT, "cores": min(threads, n), "kind": "port",
            "sample": "%d of %d stars of the same workload (full coefficient pass + light curve each), oracle C++ -O2, %d thread(s)"
                      % (n, len(stars), threads),
            "curves_per_sec": n / wall, "seconds": wall,
            "micro_tables_s_per_star": float(np.mean([r[0] for r in rs])),
            "evolve_s_per_star": float(np.mean([r[1] for r in rs])),
            "note": "the Fortran+MESA reference cannot be built in this image (no Fortran compiler, MESA r15140 absent)"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on all host threads (oracle port)."""
    rank, world, local, dist = dist_setup(args)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    import dstar_b200 as ds
    ds.NScool_init(device=0)          # input generation only (structure of the sample stars); untimed
    threads = os.cpu_count() or 1
    a2 = argparse.Namespace(**vars(args))
    side = int(np.ceil(np.sqrt(min(threads, 64))))        # the sweep builders want a square grid of stars
    a2.stars = side * side if args.workload != "basic_run" else 1
    stars = build_stars(ds, a2, 0)
    nz, zoff, H = gather_host_inputs(stars)
    n_sample = min(len(stars), threads)                   # one star per host thread
    a2.cpu_sample_stars = n_sample
    # one step = the bounded sample (<= 16 stars, at most one per host thread: ~1.3 s), so W + K steps stay
    # within a few minutes for any reasonable K
    for _ in range(args.warmup):
        cpu_baseline(stars, H, zoff, a2, threads=threads, budget_s=0)
    vals, secs = [], []
    for _ in range(max(1, args.steps)):
        r = cpu_baseline(stars, H, zoff, a2, threads=threads, budget_s=0)
        vals.append(r["value"]); secs.append(r["seconds"])
    v = float(np.mean(vals))
    r["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: sample of %d stars (nz~%d) on %d host threads" %
                       (args.workload, n_sample, int(np.median(nz)), threads)},
            "curves_per_sec": r["curves_per_sec"], "cpu_baseline": r,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    ds.NScool_shutdown()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
