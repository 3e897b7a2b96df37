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
# The cell figure apportioned to the two kernels of the cell pass by their executed-flop shares (hardware counters,
# profiles/r02_flop_counts.json: EOS 4858, neutrino 1493 per evaluation = 76.5 % / 23.5 %; SURVEY's own split, 3.7 k +
# 2.3 k, overstates the neutrino part)
FLOP_EOS = 6.0e3 * 4858.0 / (4858.0 + 1493.0)
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
    ap.add_argument("--total-stars", type=int, default=0,
                    help="strong scaling: this many stars in the whole job, split evenly over the GPUs (overrides --stars)")
    ap.add_argument("--resolution", type=float, default=0.05)
    ap.add_argument("--cpu-sample-stars", type=int, default=0, help="stars in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharing", action="store_true", help="skip the shared-tables measurement")
    ap.add_argument("--no-two-in-flight", action="store_true", help="skip the end-to-end measurement with two batches in flight")
    ap.add_argument("--other-workloads", dest="other_workloads", action="store_true", default=None,
                    help="also measure the other BASELINE configs at their named sizes (default: on for the default workload at N = 1)")
    ap.add_argument("--no-other-workloads", dest="other_workloads", action="store_false")
    ap.add_argument("--write-cpu-sample", action="store_true", help="write bench_data/cpu_sample_<workload>.npz (GPU box)")
    a = ap.parse_args()
    if a.other_workloads is None:
        a.other_workloads = (a.workload == "custom_Tc_Qimp" and a.stars == 256 and a.gpus == 1 and a.impl == "b200")
    return a


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
        if args.impl == "b200":
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


TAB_KEYS = ("Tc_cell", "Tc_face")
# distinct stars built through the NScool host layer per workload; larger batches repeat them (the host layer keeps
# every star's 4 MB of tables, so building thousands of distinct stars there would measure the host, not the path)
MAX_DISTINCT = {"custom_Tc_Qimp": 256, "fit_lightcurve": 64, "accretion": 64, "int16_demo": 64, "basic_run": 1}


def build_stars(ds, workload, n_star, resolution, rank):
    from dstar_b200 import workloads
    fn = workloads.BUILDERS[workload]
    if workload == "basic_run":
        stars = fn(resolution)
    else:
        stars = fn(n_star, resolution, offset=rank)
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
    H["ncharged"] = stars[0].get_int("ncharged")
    return nz, zoff, H


def workload_ctrl(ds, workload):
    mc = ds.MicroCtrl.defaults(); ec = ds.EvolveCtrl.defaults()
    heat = [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 3.0]
    T_atm = 1.0e8
    if workload == "custom_Tc_Qimp":
        ec.fix_atmosphere_temperature_when_accreting = 0; ec.turn_on_extra_heating = 1
        heat[6:9] = [27.0, 28.0, 1.7]; T_atm = 2.4e8
    elif workload == "basic_run":
        ec.turn_on_extra_heating = 1; heat[6:9] = [27.5, 28.5, 0.1]; T_atm = 2.4e8
    elif workload in ("fit_lightcurve", "int16_demo"):
        ec.fix_atmosphere_temperature_when_accreting = 0; ec.turn_on_extra_heating = 1
    elif workload == "accretion":
        mc.use_neutron_conductivity = 1; mc.use_superfluid_phonon_conductivity = 1; mc.turn_on_shell_Urca = 1
        ec.turn_on_extra_heating = 1; heat[6:9] = [27.0, 28.0, 1.0]; T_atm = 5.0e8
    return mc, ec, heat, T_atm


def per_star_heat(workload, n_star, heat, rank):
    """(n_star, 9) heating windows: the shallow-heating strength varies per walker exactly as the workload builders
    draw it (dstar_b200/workloads.py: same generator, same order of draws)."""
    from dstar_b200 import workloads
    h = np.tile(np.array(heat), (n_star, 1))
    if workload == "fit_lightcurve":
        rng = np.random.default_rng(workloads.SEED + 2 + 1000 * rank)
        rng.uniform(np.log10(3e16), np.log10(3e17), n_star)
        h[:, 8] = rng.uniform(0.0, 3.0, n_star)
    elif workload == "int16_demo":
        rng = np.random.default_rng(workloads.SEED + 4 + 1000 * rank)
        rng.uniform(0.0, 2.0, n_star)
        h[:, 8] = rng.uniform(0.0, 3.0, n_star)
    return h


def tile_inputs(nz, H, heat_arr, reps):
    """Repeat a batch of distinct stars `reps` times (synthetic large batches)."""
    if reps == 1:
        return nz, np.concatenate([[0], np.cumsum(nz)]).astype(np.int32), H, heat_arr
    nz2 = np.tile(nz, reps)
    zoff2 = np.concatenate([[0], np.cumsum(nz2)]).astype(np.int32)
    H2 = {}
    for k, v in H.items():
        if k in ("Zion", "Aion", "ncharged") or v is None:
            H2[k] = v
        elif k in ("eb", "em"):
            H2[k] = np.tile(v, (1, reps))
        elif k.startswith("atm_"):
            H2[k] = np.tile(v, (reps, 1))
        else:
            H2[k] = np.tile(v, reps)
    return nz2, zoff2, H2, np.tile(heat_arr, (reps, 1))


def prepare_workload(ds, workload, n_star, resolution, rank):
    """Host buffers of one workload: distinct stars through the NScool host layer (create_models builds structure,
    composition, atmosphere tables), repeated up to n_star."""
    cap = MAX_DISTINCT.get(workload, 64)
    n_distinct = min(n_star, cap)
    reps = max(1, -(-n_star // n_distinct))
    stars = build_stars(ds, workload, n_distinct, resolution, rank)
    nz, zoff, H = gather_host_inputs(stars)
    mc, ec, heat, T_atm = workload_ctrl(ds, workload)
    heat_arr = per_star_heat(workload, n_distinct, heat, rank)
    nz, zoff, H, heat_arr = tile_inputs(nz, H, heat_arr, reps)
    n_tot = len(nz)
    return {"stars": stars, "nz": nz, "zoff": zoff, "H": H, "mc": mc, "ec": ec, "heat": heat, "heat_arr": heat_arr,
            "T_atm": T_atm, "T_atm_arr": np.full(n_tot, T_atm), "n_distinct": n_distinct, "reps": reps, "n_star": n_tot,
            "workload": workload}


def upload_all(b, W):
    H = W["H"]
    b.upload_structure(*[H[k] for k in ARRS_STRUCT], H["Zion"], H["Aion"], H["Tc_cell"], H["Tc_face"])
    b.upload_evolve(*[H[k] for k in ARRS_GEOM], H["atm_lgTb"], H["atm_lgTeff"], H["atm_lgflux"],
                    np.arange(W["n_star"]), W["heat_arr"], W["T_atm_arr"], H["lnT"], H["eb"], H["em"])


def sharing_owners(W):
    """Owner lists for dstar_batch_set_table_sharing, as the host layer derives them (host_nscool.cpp
    group_identical): stars with byte-identical coefficient-pass inputs."""
    H, zoff, n = W["H"], W["zoff"], W["n_star"]
    nch = H["ncharged"]

    def group(keys_of):
        first, owner = {}, np.zeros(n, dtype=np.int32)
        for i in range(n):
            k = keys_of(i)
            owner[i] = first.setdefault(k, i)
        return owner

    def sl(name, i, w=1):
        return H[name][zoff[i] * w:zoff[i + 1] * w]

    def cell_key(i):
        io = sl("ionic", i, 11).reshape(-1, 11).copy(); io[:, 10] = 0.0
        parts = [sl("rho", i), sl("P", i), sl("P_bar", i), sl("dm", i), io.ravel(), sl("Yion", i, nch)]
        if H["Tc_cell"] is not None:
            parts.append(sl("Tc_cell", i, 3))
        return b"".join(np.ascontiguousarray(x).tobytes() for x in parts)
    cell = group(cell_key)

    def face_key(i):
        parts = [np.array([cell[i]], dtype=np.int64), sl("rho_bar", i)[1:], sl("ionic_bar", i, 11), sl("Yion_bar", i, nch)]
        if H["Tc_face"] is not None:
            parts.append(sl("Tc_face", i, 3))
        return b"".join(np.ascontiguousarray(x).tobytes() for x in parts)
    return cell, group(face_key)


def measure_resident(ds, ctx, W, steps, warmup, share=False, sampler=None, sync=None):
    """K steps of the hot path on inputs resident in HBM; device times from the library's CUDA events."""
    b = ds.Batch(ctx, W["zoff"], W["H"]["ncharged"])
    upload_all(b, W)
    sets = None
    if share:
        cell, face = sharing_owners(W)
        b.set_table_sharing(1, cell); b.set_table_sharing(2, face)
        sets = (int(np.sum(cell == np.arange(W["n_star"]))), int(np.sum(face == np.arange(W["n_star"]))))

    launches = [0]
    eos_ms = [0.0]

    def step():
        b.reset_state()
        b.micro_tables(W["mc"], 3)
        t_c, t_f = b.kernel_ms(0), b.kernel_ms(1)
        eos_ms[0] += b.kernel_ms(3)
        launches[0] += b.launch_count()
        b.evolve(W["ec"], 0)
        launches[0] += b.launch_count()
        return t_c, t_f, b.kernel_ms(2)
    for _ in range(warmup):
        step()
    launches[0] = 0; eos_ms[0] = 0.0
    if sync:
        sync()
    if sampler:
        sampler.start()
    tc = tf = te = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        a_, b_, c_ = step()
        tc += a_; tf += b_; te += c_
    wall = time.perf_counter() - t0
    if sync:
        sync()
    clocks = sampler.stop() if sampler else None
    res = b.download_results(full=False)
    out = {"cells_ms": tc / steps, "faces_ms": tf / steps, "evolve_ms": te / steps, "step_ms": (tc + tf + te) / steps,
           "wall_s": wall, "clocks": clocks, "res": res, "converged": bool(np.all(res["idid"] >= 0)), "table_sets": sets,
           "launches": launches[0], "eos_ms": eos_ms[0] / steps, "redo": [int(x) for x in b.redo_counts()],
           "h2d": b._h2d_structure + b._h2d_evolve}
    return out, b


def load_json(name, default):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return default


def workload_roofline(workload, W, m, fp64_peak):
    """FP64 roofline of the coefficient pass of one workload.  flops per evaluation: profiles/r02_flop_counts.json
    (FP64 operations EXECUTED per evaluation, hardware counters: dadd + dmul + 2 dfma per thread, one ncu pass per
    workload, tools/count_flops.py) next to SURVEY 8d's algorithmic estimate for the one-species / electron-only case."""
    counts = load_json("r02_flop_counts.json", {}).get(workload)
    evals_cell = int(np.sum(W["nz"])) * 128
    eos_ms = m.get("eos_ms", 0.0)
    out = {"evals_per_pass": evals_cell, "kernel_ms": {"cells": m["cells_ms"], "cells_eos": eos_ms, "cells_nu": m["cells_ms"] - eos_ms,
                                                       "faces": m["faces_ms"], "evolve": m["evolve_ms"]}}
    if counts:
        for kind, ms in (("eos", eos_ms), ("nu", m["cells_ms"] - eos_ms), ("cell", m["cells_ms"]), ("face", m["faces_ms"])):
            if kind not in counts or ms <= 0:
                continue
            fl = counts[kind]["flop_per_eval"]
            tf_ = fl * evals_cell / (ms * 1e-3) / 1e12
            out[kind] = {"executed_flop_per_eval": fl, "achieved_tflops": tf_, "frac": tf_ / fp64_peak if fp64_peak > 0 else None,
                         "fp64_inst_per_eval": counts[kind].get("inst_per_eval")}
        out["flop_source"] = counts.get("source")
    return out


def e2e_entry(evals, curves, serial_s, pair_s, h2d, d2h):
    """End-to-end entry of the bench line: the serial loop (one batch, one host thread) and, when measured, the loop with
    two batches in flight; `value` is the faster and `mode` names it."""
    best, mode = serial_s, "one batch at a time"
    if pair_s is not None and pair_s < serial_s:
        best, mode = pair_s, "two batches in flight (two contexts on the device, one host thread each)"
    out = {"value": evals / best, "unit": UNIT, "curves_per_sec": curves / best, "ms_per_step": best * 1e3, "mode": mode,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "one_batch_at_a_time": {"value": evals / serial_s, "curves_per_sec": curves / serial_s, "ms_per_step": serial_s * 1e3}}
    if pair_s is not None:
        out["two_batches_in_flight"] = {"value": evals / pair_s, "curves_per_sec": curves / pair_s, "ms_per_step": pair_s * 1e3}
        out["note"] = ("wall clock per step, every step copying its inputs host->device and its results back; with two batches "
                       "in flight one batch's copies, host work and the idle SMs of its evolve launch's last wave overlap the "
                       "other's kernels, so the figure can exceed `value`, which times one batch's kernels back to back")
    return out


def measure_two_in_flight(ds, local, args, W, bb, dist):
    """Seconds per step of the end-to-end loop with TWO batches in flight, as a sweep driver that prepares the next batch
    while the current one computes: a second context (its own stream) on the same device, one host thread per batch, the
    threads take the steps in turn.  Every step still uploads all its inputs and downloads its results; one batch's
    copies, host work and the idle SMs of its last launch's tail overlap the other's kernels.  Returns None if any rank
    failed; every rank takes part in the two barriers and the reduction whatever happens to it."""
    mc, ec = W["mc"], W["ec"]
    errs = []
    ctx2 = bb2 = None
    pair = [bb, None]

    def run_steps(which, n):
        try:
            for _ in range(n):
                upload_all(pair[which], W)
                pair[which].micro_tables(mc, 3)
                pair[which].evolve(ec, 0)
                pair[which].download_results(full=False)
        except Exception as e_:   # noqa: BLE001
            errs.append(e_)

    def run_pair(n_total):
        n0 = (n_total + 1) // 2
        ts = [threading.Thread(target=run_steps, args=(0, n0)), threading.Thread(target=run_steps, args=(1, n_total - n0))]
        for t_ in ts:
            t_.start()
        for t_ in ts:
            t_.join()
    try:
        try:
            ctx2 = ds.Context(local, gaps=("ns", "sfb03" if args.workload != "accretion" else "gc", "t72"))
            bb2 = ds.Batch(ctx2, W["zoff"], W["H"]["ncharged"])
            pair[1] = bb2
            run_pair(2)
        except Exception as e_:   # noqa: BLE001
            errs.append(e_)
        barrier(dist, local)
        t0 = time.perf_counter()
        if not errs:
            run_pair(args.steps)
        barrier(dist, local)
        s_ = allmax(dist, 1.0e30 if errs else (time.perf_counter() - t0) / args.steps, local)
        if errs:
            print("two batches in flight: not measured (%s)" % errs[0], file=sys.stderr)
        return None if s_ >= 1.0e29 else s_
    finally:
        if bb2 is not None:
            bb2.close()
        if ctx2 is not None:
            ctx2.close()


def run_b200(args):
    rank, world, local, dist = dist_setup(args)
    import dstar_b200 as ds
    ds.NScool_init(device=local)
    ctx = ds.Context(local, gaps=("ns", "sfb03" if args.workload != "accretion" else "gc", "t72"))
    if args.total_stars > 0:   # strong scaling: a fixed job dealt out in contiguous blocks, as dstar_b200/dist.py shard_range
        from dstar_b200.dist import shard_range
        if args.total_stars % world:
            raise SystemExit("--total-stars must be a multiple of the number of GPUs (the line reports rank 0's share x N)")
        lo_, hi_ = shard_range(args.total_stars, rank, world)
        args.stars = hi_ - lo_
    W = prepare_workload(ds, args.workload, args.stars, args.resolution, rank)
    stars, nz, zoff, H = W["stars"], W["nz"], W["zoff"], W["H"]
    n_star, total = W["n_star"], int(zoff[-1])
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
    mc, ec, heat, T_atm = W["mc"], W["ec"], W["heat"], W["T_atm"]
    evals_cell = int(np.sum(nz)) * 128
    evals = 2 * evals_cell

    # ---------------- resident measurement
    m, b = measure_resident(ds, ctx, W, args.steps, args.warmup, share=False, sampler=ClockSampler(local),
                            sync=lambda: barrier(dist, local))
    tc, tf, te = m["cells_ms"] * args.steps, m["faces_ms"] * args.steps, m["evolve_ms"] * args.steps
    wall_res, clocks, res, h2d = m["wall_s"], m["clocks"], m["res"], m["h2d"]
    ms_step = allmax(dist, m["step_ms"], local)   # device time of the step's kernels (CUDA events in the library)
    # ---------------- end-to-end through the C ABI with host buffers
    # (the batch handle is reused between steps, as a sweep driver would: device buffers are allocated once,
    #  every input is copied host->device again on every step)
    bb = ds.Batch(ctx, zoff, H["ncharged"])

    def step_e2e():
        upload_all(bb, W)
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
    # The same steps with TWO batches in flight, as a sweep driver that prepares the next batch while the current one
    # computes: a second context (its own stream) on the same device, one host thread per batch, the threads take the
    # steps in turn.  Every step still uploads all its inputs and downloads its results; one batch's copies and host
    # work now overlap the other's kernels.  Reported next to the serial figure; `e2e.value` is the better of the two
    # and says which.
    e2e2_s = None
    if not args.no_two_in_flight:
        e2e2_s = measure_two_in_flight(ds, local, args, W, bb, dist)
    bb.close()
    # ---------------- the same step with table sharing (stars with identical coefficient-pass inputs share one table
    # set; evals/s above stays the unshared figure, this is what a sweep driver gets in curves/s)
    shared = None
    if not args.no_sharing:
        ms_, bs_ = measure_resident(ds, ctx, W, max(3, args.steps // 2), 2, share=True)
        same = bool(np.array_equal(ms_["res"]["Teff_monitor"], res["Teff_monitor"]))
        sh_step = allmax(dist, ms_["step_ms"], local)
        shared = {"ms_per_step": sh_step, "curves_per_sec": n_star * world / (sh_step * 1e-3),
                  "kernel_ms": {"cells": ms_["cells_ms"], "faces": ms_["faces_ms"], "evolve": ms_["evolve_ms"]},
                  "cell_table_sets": ms_["table_sets"][0], "face_table_sets": ms_["table_sets"][1], "stars": n_star,
                  "light_curves_bit_identical_to_unshared": same,
                  "table_bytes": 4 * 4096 * int(np.median(nz)) * (3 * ms_["table_sets"][0] + ms_["table_sets"][1]),
                  "table_bytes_unshared": 4 * 4 * 4096 * total}
        bs_.close()
    # ---------------- roofline of the dominant kernel, FP64 pipe
    fp64_peak = ctx.fp64_peak_tflops()
    # The cell pass runs as two halves: the EOS half (eval_crust_eos -> lnCp, lnGamma; ds_block_kernel<EOS> is the
    # dominant kernel, ~46 % of the step) and the neutrino half (crust neutrino emissivity -> lnEnu).  Durations are the
    # library's CUDA-event times on the launching stream, averaged over the timed steps; the EOS time includes its
    # block sort, its (normally empty) plain-division launch and its interpolation kernel.  `roofline` is the EOS kernel; `cell_pass` is both kernels together, the figure
    # comparable with round 1's fused kernel (6.0 kflop per evaluation over the cell time).
    cell_ms = tc / args.steps
    eos_ms = m["eos_ms"] if m["eos_ms"] > 0 else cell_ms
    nu_ms = cell_ms - eos_ms
    micro_ms = (tc + tf) / args.steps
    split = m["eos_ms"] > 0
    flop_dom = FLOP_EOS if split else FLOP_CELL
    achieved_tf = flop_dom * evals_cell / (eos_ms * 1e-3) / 1e12
    cell_tf = FLOP_CELL * evals_cell / (cell_ms * 1e-3) / 1e12
    both_tf = (FLOP_CELL + FLOP_FACE) * evals_cell / (micro_ms * 1e-3) / 1e12
    alg_bytes = 64.0 * evals_cell       # ~64 B / eval (32 B of spline coefficients per knot and quantity)
    alg_bytes_dom = (2.0 / 3.0 if split else 1.0) * alg_bytes   # the EOS kernel writes two of the three cell tables
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full`
    # capture of the named configuration (profiles/r02_traffic.json names the capture); scaled by the zone count when
    # the run's batch differs from the captured one (the traffic is spill frames + tables, both proportional to the zones)
    tr = load_json("r02_traffic.json", {}).get(args.workload)
    traffic, traffic_note = None, "no ncu capture for this workload"
    if tr and split:
        traffic = tr["eos_dram_bytes_per_launch"] * (total / float(tr["zones"]))
        traffic_note = "bytes per launch from %s (%d zones), scaled to %d zones; algorithmic bytes per launch %.3g" % (
            tr["capture"], tr["zones"], total, alg_bytes_dom)
    wr = workload_roofline(args.workload, W, m, fp64_peak)
    roofline = {"bound": "fp64", "kernel": "ds_block_kernel<EOS> (cell pass, EOS half: evaluation + block sort + interpolation)" if split else "ds_micro_kernel<cell>",
                "achieved": achieved_tf, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": "DFMA-chain probe measured in this run (MEASURED_PEAKS.json has no FP64 figure); the "
                               "reference's formulas may not be contracted into FMAs, so 0.5 is the ceiling of frac",
                "flops_per_eval": {"eos": FLOP_EOS, "nu": FLOP_CELL - FLOP_EOS, "cell": FLOP_CELL, "face": FLOP_FACE,
                                   "source": "SURVEY.md 8d algorithmic estimate (cell 6.0 k, face 1.6 k); the cell figure is apportioned to the EOS and neutrino kernels by their executed-flop shares (profiles/r02_flop_counts.json)"},
                "cell_pass": {"achieved": cell_tf, "frac": cell_tf / fp64_peak if fp64_peak > 0 else None,
                              "note": "EOS + neutrino kernels together: 6.0 kflop per evaluation over the cell time"},
                "executed": {k: wr.get(k) for k in ("eos", "nu", "cell", "face", "flop_source")},
                "cell_plus_face": {"achieved": both_tf, "frac": both_tf / fp64_peak if fp64_peak > 0 else None},
                "ncu": load_json("r02_traffic.json", {}).get("ncu_" + args.workload),
                "hbm": {"achieved_gbs": alg_bytes_dom / (eos_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
                "kernel_ms": {"cells": cell_ms, "cells_eos": eos_ms, "cells_nu": nu_ms, "faces": tf / args.steps,
                              "evolve": te / args.steps},
                "redo_zones": m["redo"]}
    b.close()
    # ---------------- the other named configs at their BASELINE sizes (rank-local; N = 1 default run only)
    others = None
    if args.other_workloads and world == 1:
        others = {}
        for name, wl, n_ in (("fit_lightcurve_1024", "fit_lightcurve", 1024), ("accretion_4096", "accretion", 4096),
                             ("int16_1024_per_gpu", "int16_demo", 1024)):
            if wl == args.workload:
                continue
            cx2 = ds.Context(local, gaps=("ns", "gc" if wl == "accretion" else "sfb03", "t72"))
            W2 = prepare_workload(ds, wl, n_, args.resolution, rank)
            m2, b2 = measure_resident(ds, cx2, W2, 3, 1, share=False)
            ev2 = 2 * int(np.sum(W2["nz"])) * 128
            ent = {"stars": W2["n_star"], "distinct_stars": W2["n_distinct"], "zones": int(W2["zoff"][-1]), "nz": int(np.median(W2["nz"])),
                   "epochs": int(W2["H"]["em"].shape[0]), "species": int(W2["H"]["ncharged"]), "ms_per_step": m2["step_ms"],
                   "evals_per_sec": ev2 / (m2["step_ms"] * 1e-3), "curves_per_sec": W2["n_star"] / (m2["step_ms"] * 1e-3),
                   "all_stars_converged": m2["converged"], "roofline": workload_roofline(wl, W2, m2, fp64_peak),
                   "table_bytes": 4 * 4 * 4096 * int(W2["zoff"][-1])}
            b2.close()
            ms2, bs2 = measure_resident(ds, cx2, W2, 2, 1, share=True)
            ent["shared_tables"] = {"ms_per_step": ms2["step_ms"], "curves_per_sec": W2["n_star"] / (ms2["step_ms"] * 1e-3),
                                    "cell_table_sets": ms2["table_sets"][0], "face_table_sets": ms2["table_sets"][1],
                                    "kernel_ms": {"cells": ms2["cells_ms"], "faces": ms2["faces_ms"], "evolve": ms2["evolve_ms"]}}
            bs2.close(); cx2.close()
            for s_ in W2["stars"]:
                s_.free()
            others[name] = ent
    # ---------------- CPU baseline (oracle) on rank 0, N = 1
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # a bounded sample: 8 stars spread over the batch on ONE core (~10 s of CPU work)
        picks = sorted(set(int(x) for x in np.linspace(0, W["n_distinct"] - 1, min(W["n_distinct"], args.cpu_sample_stars or 8)).round()))
        a2 = argparse.Namespace(**vars(args)); a2.cpu_sample_stars = len(picks)
        cpu = cpu_baseline(sample_from_inputs(H, zoff, picks), a2, threads=1)
    ok = m["converged"]
    desc = "%s: %d stars/GPU x nz~%d, %d epochs, %s crust, bc09 atmosphere (device-generated table), rebuilt HELM/gap tables" % (
        args.workload, n_star, int(np.median(nz)), H["em"].shape[0], "abuntime (synthetic)" if args.workload == "int16_demo" else "HZ90")
    if W["reps"] > 1:
        desc += "; %d distinct stars repeated %d times" % (W["n_distinct"], W["reps"])
    line = {"metric": METRIC, "value": evals * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if args.total_stars > 0 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "stars_per_gpu": n_star, "zones_total_per_gpu": total, "seed": 20261019,
                       "l2_policy": "inputs+tables per step %.0f MB > 126 MB L2" % ((4 * 4096.0 * total + h2d) / 1e6)},
            "curves_per_sec": n_star * world / (ms_step * 1e-3),
            "e2e": e2e_entry(evals * world, n_star * world, e2e_s, e2e2_s, int(h2d), int(d2h)),
            # kernels launched in the timed region, counted by the library (per step: zone pre-pass, EOS kernel + its
            # plain-division pass, neutrino kernel + its plain-division pass, zone pre-pass, face kernel, evolve)
            "gpu_launches": m["launches"], "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "all_stars_converged": ok, "wall_s_resident_loop": wall_res, "shared_tables": shared, "other_workloads": others}
    if dist is not None:
        # final gather of the light curves (the only collective of the job)
        from dstar_b200.dist import gather_monitors
        full = gather_monitors(dist, res["Teff_monitor"], n_star * world, rank, world, device="cuda:%d" % local)
        if rank == 0:
            line["gathered_monitor_shape"] = list(full.shape)
    if rank == 0:
        print(json.dumps(line))
    ctx.close(); ds.NScool_shutdown()
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------- CPU arm (oracle port)
ZONE_W = {"ionic": 11, "ionic_bar": 11, "Tc_cell": 3, "Tc_face": 3}


class SampleStar:
    """One star of a flat input set, with the accessors the oracle helpers use (tests/parity_helpers.py)."""

    def __init__(self, H, zoff, i):
        self.H, self.i, self.a, self.e = H, i, int(zoff[i]), int(zoff[i + 1])

    def get_int(self, name):
        return {"nz": self.e - self.a, "ncharged": int(self.H["ncharged"])}[name]

    def array(self, name):
        H = self.H
        if name in ("Zion", "Aion"):
            return np.asarray(H[name], dtype=float)
        if name.startswith("atm_"):
            return np.asarray(H[name][self.i], dtype=float)
        if name == "epoch_boundaries":
            return np.asarray(H["eb"][:, self.i], dtype=float)
        if name == "epoch_Mdots":
            return np.asarray(H["em"][:, self.i], dtype=float)
        w = ZONE_W.get(name, int(H["ncharged"]) if name in ("Yion", "Yion_bar") else 1)
        return np.asarray(H[name][self.a * w:self.e * w], dtype=float)


def sample_from_inputs(H, zoff, picks):
    """A self-contained sample (the arrays of the picked stars only)."""
    keys = ARRS_STRUCT + ARRS_GEOM + ["lnT"] + [k for k in TAB_KEYS if H.get(k) is not None]
    nch = int(H["ncharged"])
    out = {"ncharged": nch, "Zion": np.asarray(H["Zion"]), "Aion": np.asarray(H["Aion"])}
    for k in keys:
        w = ZONE_W.get(k, nch if k in ("Yion", "Yion_bar") else 1)
        out[k] = np.concatenate([np.asarray(H[k][int(zoff[i]) * w:int(zoff[i + 1]) * w]) for i in picks])
    for k in TAB_KEYS:
        out.setdefault(k, None)
    for k in ("atm_lgTb", "atm_lgTeff", "atm_lgflux"):
        out[k] = np.stack([np.asarray(H[k][i]) for i in picks])
    out["eb"] = np.stack([np.asarray(H["eb"][:, i]) for i in picks], axis=1)
    out["em"] = np.stack([np.asarray(H["em"][:, i]) for i in picks], axis=1)
    nz = np.array([int(zoff[i + 1] - zoff[i]) for i in picks])
    out["zoff"] = np.concatenate([[0], np.cumsum(nz)]).astype(np.int32)
    return out


def sample_path(workload):
    return os.path.join(ROOT, "bench_data", "cpu_sample_%s.npz" % workload)


def load_sample(workload):
    d = np.load(sample_path(workload), allow_pickle=False)
    out = {k: d[k] for k in d.files}
    out["ncharged"] = int(out["ncharged"])
    for k in TAB_KEYS:
        if k not in out:
            out[k] = None
    return out


def oracle_star_inputs(S, i):
    st = SampleStar(S, S["zoff"], i)
    kw = {k: st.array(k) for k in ARRS_STRUCT}
    kw.update(nz=st.get_int("nz"), ncharged=st.get_int("ncharged"), Zion=st.array("Zion"), Aion=st.array("Aion"),
              Tc_cell=None if S["Tc_cell"] is None else st.array("Tc_cell"),
              Tc_face=None if S["Tc_face"] is None else st.array("Tc_face"))
    return st, kw


def workload_ctrl_plain(workload):
    """workload_ctrl without touching the product library (the reference arm must not load it)."""
    class NS:   # the few fields the CPU arm reads
        pass
    mc, ec = NS(), NS()
    mc.use_neutron_conductivity = mc.use_superfluid_phonon_conductivity = mc.turn_on_shell_Urca = 0
    ec.fix_atmosphere_temperature_when_accreting = 1; ec.turn_on_extra_heating = 0
    heat = [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 3.0]
    T_atm = 1.0e8
    if workload == "custom_Tc_Qimp":
        ec.fix_atmosphere_temperature_when_accreting = 0; ec.turn_on_extra_heating = 1
        heat[6:9] = [27.0, 28.0, 1.7]; T_atm = 2.4e8
    elif workload == "basic_run":
        ec.turn_on_extra_heating = 1; heat[6:9] = [27.5, 28.5, 0.1]; T_atm = 2.4e8
    elif workload in ("fit_lightcurve", "int16_demo"):
        ec.fix_atmosphere_temperature_when_accreting = 0; ec.turn_on_extra_heating = 1
    elif workload == "accretion":
        mc.use_neutron_conductivity = 1; mc.use_superfluid_phonon_conductivity = 1; mc.turn_on_shell_Urca = 1
        ec.turn_on_extra_heating = 1; heat[6:9] = [27.0, 28.0, 1.0]; T_atm = 5.0e8
    return mc, ec, heat, T_atm


def cpu_baseline(S, args, threads):
    """Time the CPU oracle on a bounded sample of the same workload: `threads` stars in parallel, each
    doing its full coefficient pass and light curve.  S: a sample as sample_from_inputs / load_sample give it."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from parity_helpers import load_oracle_gaps, oracle_evolve_star
    from concurrent.futures import ThreadPoolExecutor
    o = Oracle()
    load_oracle_gaps(o, "ns", "sfb03" if args.workload != "accretion" else "gc", "t72")
    mc, ec, heat, T_atm = workload_ctrl_plain(args.workload)
    cond = (mc.use_neutron_conductivity, mc.use_superfluid_phonon_conductivity, 1, 1, 10.0, 9.0)
    n_avail = len(S["zoff"]) - 1
    n = min(args.cpu_sample_stars or threads, n_avail)

    def one(i):
        st, kw = oracle_star_inputs(S, i)
        t0 = time.perf_counter()
        tabs = o.micro_tables(cond_ctrl=cond, shell_Urca=(mc.turn_on_shell_Urca, 1.0, 28.5), **kw)
        t1 = time.perf_counter()
        oracle_evolve_star(o, st, st.array("lnT"), tabs=tabs, heat=heat, T_atm=T_atm, quirks=1,
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
            "sample": "%d stars of the same workload (full coefficient pass + light curve each), oracle C++ -O2, %d thread(s)"
                      % (n, threads),
            "curves_per_sec": n / wall, "seconds": wall,
            "micro_tables_s_per_star": float(np.mean([r[0] for r in rs])),
            "evolve_s_per_star": float(np.mean([r[1] for r in rs])),
            "note": "the Fortran+MESA reference cannot be built in this image (no Fortran compiler, MESA r15140 absent)"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on all host threads (oracle port).  The sample stars come
    from a committed file (bench_data/cpu_sample_<workload>.npz, written on a GPU box by `bench.py --write-cpu-sample`):
    this arm loads neither the product library nor a GPU."""
    rank, world, local, dist = dist_setup(args)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    threads = os.cpu_count() or 1
    if not os.path.exists(sample_path(args.workload)):
        print(json.dumps({"impl": "reference", "unavailable": "no committed CPU sample for workload %s (bench_data/)" % args.workload}))
        return
    S = load_sample(args.workload)
    n_avail = len(S["zoff"]) - 1
    a2 = argparse.Namespace(**vars(args))
    n_sample = min(n_avail, args.cpu_sample_stars or threads)   # one star per host thread
    a2.cpu_sample_stars = n_sample
    nz = np.diff(S["zoff"])
    # one step = the bounded sample (<= 32 stars, at most one per host thread: ~1.3 s), so W + K steps stay
    # within a few minutes for any reasonable K
    for _ in range(args.warmup):
        cpu_baseline(S, a2, threads=threads)
    vals, secs = [], []
    for _ in range(max(1, args.steps)):
        r = cpu_baseline(S, a2, threads=threads)
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
    try:   # evidence for the driver: this arm ran without the product's native library
        line["product_library_loaded"] = "libdstar_b200" in open("/proc/self/maps").read()
    except OSError:
        line["product_library_loaded"] = None
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def write_cpu_sample(args):
    """GPU box: build the workload's stars through the product's host layer and store a spread of 32 of them."""
    import dstar_b200 as ds
    ds.NScool_init(device=0)
    W = prepare_workload(ds, args.workload, min(args.stars, MAX_DISTINCT.get(args.workload, 64)), args.resolution, 0)
    n = W["n_distinct"]
    picks = sorted(set(int(x) for x in np.linspace(0, n - 1, min(n, 32)).round()))
    S = sample_from_inputs(W["H"], W["zoff"], picks)
    os.makedirs(os.path.dirname(sample_path(args.workload)), exist_ok=True)
    np.savez_compressed(sample_path(args.workload), **{k: v for k, v in S.items() if v is not None})
    print("wrote", sample_path(args.workload), "stars", len(picks), "of", n)
    ds.NScool_shutdown()


if __name__ == "__main__":
    a = parse()
    if a.write_cpu_sample:
        write_cpu_sample(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
