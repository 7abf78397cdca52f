#!/usr/bin/env python
"""bench.py -- particle-updates/s of one full COLA particle-mesh step on B200, with its own parity check.

A "step" = MoveParticles + PtoMesh (+ in-step P(k)) + ComputeFifthForce + Forces + MtoParticles +
[SCALEDEPENDENT: the displacement fields of the step] + Kick + Drift (main.c:474-592 of the reference) on synthetic
Gaussian 2LPT initial conditions drawn from the reference's bundled CAMB table (tests/golden/input_power_spectrum.npz),
stepped along the reference's COLA schedule (z_init = 9 -> 0, 30 steps linear in a).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--nmesh 512] [--model fofr|dgp|lcdm] [--impl ours|reference]

Workload: f(R) with screening, SCALEDEPENDENT growth, Npart = Nmesh^3 = 512^3, double grids, at every N (strong
scaling; 512^3 is the largest configuration of BASELINE.json that fits one GPU).  At N = 8 the north-star target
(1024^3 on 8 GPUs) is measured as well and reported under "target".

One JSON line on stdout (rank 0):
  value         device-timed, particles resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e           the same step driven through the C ABI with HOST particle buffers (pinned): every rank uploads Pos / Vel / ID
                of its particles and downloads them again inside the timed region
  roofline      the dominant hand-written kernel (live CUDA-event time of its phase) and the whole step against HBM + NVLink
  parity        P(k) of the CUDA library against the UNMODIFIED reference (oracle/_ref) run on the same parameter file and
                seed at 128^3: every in-step P(k) file, bin by bin for k < k_Nyquist / 2 (compute_pofk.c:229-236); on
                several ranks additionally P(k) of the last timed step against a 1-GPU run of the same mesh
  cpu_baseline  the unmodified reference on the host's cores (multi-process MPI stand-in), a bounded sample
--impl reference: the reference's own CPU path timed the same way, same `config`.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed when NCCL_DEBUG is set) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

OMEGA, SIGMA8, Z_INIT, NSTEPS_RUN = 0.267, 0.8, 9.0, 30
FOFR0, NFOFR, RCH0, RSMOOTH = 1e-5, 1.0, 1.0, 1.0
DEFAULT_NMESH, TARGET_NMESH, PARITY_NMESH = 512, 1024, 128
METRIC = "particle-updates/sec per COLA PM step"
# compulsory HBM bytes per particle-step of a maximally fused step
ALGO_BYTES = {"lcdm": lambda g: 120 + 32 * g, "fofr": lambda g: 120 + 51 * g, "dgp": lambda g: 120 + 51 * g}
# SCALEDEPENDENT add-on (section 8(d)): merged two-order fields 52 g, reference-structured (four fields) 48 + 100 g
ALGO_BYTES_SD = {"merged": lambda g: 52 * g, "ref": lambda g: 48 + 100 * g}
NVLINK_GBS = 770.0       # measured peer copy per direction (B200_PROFILING.md)


def box_for(nmesh):
    return 200.0 * nmesh / 256.0 if nmesh > 256 else 200.0     # keeps >= 0.78 Mpc/h cells like the example runs


def use_sd_of(args):
    return args.scale_dependent if args.scale_dependent >= 0 else int(args.model in ("fofr", "dgp"))


def config_of(args, nmesh):
    """The workload, named the same way by both arms (the reference arm times a bounded sample of it)."""
    use_sd = use_sd_of(args)
    g = args.grid_bytes
    return {"workload": "%s%s COLA step, Npart=Nmesh=%d^3, Box=%g Mpc/h, P(k) every step, z=9->0 in %d steps, %s"
                        % (args.model, " with screening" if args.model != "lcdm" else "", nmesh, box_for(nmesh), NSTEPS_RUN,
                           ("SCALEDEPENDENT growth (reference build MODEL=%s, use_lcdm_growth_factors=0)"
                            % {"fofr": "FOFR", "dgp": "DGP -DSCALEDEPENDENT"}.get(args.model, args.model)) if use_sd
                           else "scale-independent growth (reference build MODEL=FOFR_LCDM / DGP)"),
            "nmesh": nmesh, "npart": nmesh ** 3, "grid_bytes": g, "scale_dependent": use_sd,
            "l2": "inputs larger than L2 (particles %.1f GB, grids %.1f GB each)" % (nmesh ** 3 * 56 / 1e9, nmesh ** 3 * g / 1e9)}


# ------------------------------------------------------------------ synthetic initial conditions

def power_table():
    t = np.load(os.path.join(ROOT, "tests", "golden", "input_power_spectrum.npz"))
    k, P = t["k"], t["P"]
    # sigma8 normalisation (power.c:481-516): top-hat R = 8 Mpc/h
    kk = np.exp(np.linspace(np.log(k[0]), np.log(k[-1]), 20000))
    Pk = np.exp(np.interp(np.log(kk), np.log(k), np.log(P)))
    kr = 8.0 * kk
    w = 3.0 * (np.sin(kr) / kr ** 3 - np.cos(kr) / kr ** 2)
    sig2 = np.trapezoid(kk ** 3 * w * w * Pk, np.log(kk)) / (2.0 * np.pi ** 2)
    return np.log10(k), np.log10(P * SIGMA8 ** 2 / sig2)


def amplitude_table(nmesh, box):
    """sqrt-free P(k) at every integer |d|^2 = m (what the C adapter fills by calling PowerSpec)."""
    lk, lP = power_table()
    h = nmesh // 2
    m = np.arange(3 * h * h + 1, dtype=np.float64)
    kmag = 2.0 * np.pi / box * np.sqrt(m)
    with np.errstate(divide="ignore"):
        lkm = np.log10(kmag)
    P = 10.0 ** np.interp(lkm, lk, lP, left=-np.inf, right=-np.inf)    # PowerSpec_Tabulated: 0 outside the table
    P[0] = 0.0
    return P


# ------------------------------------------------------------------ clocks

class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML polled every 50 ms (rare enough not to disturb the
    launches of a step, which has a handful of host synchronisations); `nvidia-smi` every 200 ms when NVML cannot be
    initialised."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device=0, period=0.2):
        super().__init__(daemon=True)
        self.device, self.period, self.stop_flag = device, period, False
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.masks = [(pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                          (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
            self.nvml = pynvml
            self.period = 0.05
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        for mask, nm in self.masks:
            if bits & int(mask):
                self.reasons.add(nm)

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown," \
            "clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.sm.append(float(out[0]))
        self.sm_max = float(out[1])
        for nm, v in zip(self.NAMES, out[2:]):
            if "Active" in v and "Not" not in v:
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self.nvml is not None:        # NVML query failed: fall back to nvidia-smi for the rest of the run
                    self.nvml, self.period = None, 0.2
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml, 50 ms period" if self.nvml is not None else "nvidia-smi, 200 ms period"}


# ------------------------------------------------------------------ our arm

class Stepper:
    """Drives the library exactly as main.c's loop does (main.c:394-611)."""

    def __init__(self, pm, cos, model, box, sd=None, sd_mode="merged"):
        from mgpicola_b200 import cosmology
        self.pm, self.cos, self.model, self.box = pm, cos, model, box
        self.sd, self.sd_mode = sd, sd_mode
        self.sched = cosmology.schedule(Z_INIT, [(0.0, NSTEPS_RUN)])
        self.cosmology = cosmology
        self.i = 0
        # the per-step host scalars (what main.c / cosmo.c compute) for every step of the schedule
        self.pre = []
        for s in self.sched[:NSTEPS_RUN]:
            A, AI, AF, AFF = s["A"], s["AI"], s["AF"], s["AFF"]
            Di, Di2 = cos.growth_D(A), cos.growth_D2(A)
            # scale-dependent growth: the tables assign_displacment_field_to_particles needs (main.c:496-501), in call order
            tabs = [sd.table(ft, o, A, AFF) for ft in (3, 2) for o in (1, 2)] if sd is not None else None
            self.pre.append((self.scalars(A), A, cos.Sphi(AI, AF, A), cos.growth_ddDddy(A), cos.growth_ddD2ddy(A),
                             cos.Sq(A, AFF, AF), cos.growth_D(AFF) - Di, cos.growth_D2(AFF) - Di2, tabs))

    def scalars(self, A):
        if self.model == "fofr":
            return self.pm.scalars(compute_pofk=1, **self.cosmology.fofr_step_scalars(A, OMEGA, self.box, FOFR0, NFOFR))
        if self.model == "dgp":
            return self.pm.scalars(compute_pofk=1, **self.cosmology.dgp_step_scalars(A, OMEGA, RCH0, RSMOOTH))
        return self.pm.scalars(a=A, compute_pofk=1)

    def step(self):
        sc, A, dda, ddD, ddD2, dyyy, dD, dD2, tabs = self.pre[self.i % NSTEPS_RUN]   # stay inside the regular (non-output) steps
        self.i += 1
        pm = self.pm
        pm.GetDisplacements(sc)
        if tabs is not None:
            if self.sd_mode == "merged":
                pm.assign_displacement_fields_merged(3, tabs[0], tabs[1])
                pm.assign_displacement_fields_merged(2, tabs[2], tabs[3])
            else:
                for i, (ft, o) in enumerate(((3, 1), (3, 2), (2, 1), (2, 2))):
                    pm.assign_displacment_field_to_particles(ft, o, tabs[i])
        pm.Kick(A, dda, ddD, ddD2)
        pm.Drift(dyyy, dD, dD2)


class Dist:
    """torch.distributed plumbing of the run (one process per GPU)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        assert self.world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def nccl_id(self):
        import mgpicola_b200 as mgp
        if self.world == 1:
            return None
        ids = [mgp.nccl_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxf(self, vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sumf(self, vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def make_context(args, N, rank, world, local, nccl_id, env=None):
    """A library context + the stepper of the workload at mesh N, initial conditions generated on the GPU(s)."""
    import mgpicola_b200 as mgp
    from mgpicola_b200 import cosmology
    saved = {}
    for k, v in (env or {}).items():            # engine knobs are read when the context is created
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        g, box, model = args.grid_bytes, box_for(N), args.model
        cos = cosmology.LCDM(OMEGA, Z_INIT)
        model_id = {"lcdm": mgp.MODEL_NONE, "fofr": mgp.MODEL_FOFR, "dgp": mgp.MODEL_DGP}[model]
        use_sd = use_sd_of(args)
        sd = cosmology.ScaleDependentGrowth(cos, box, N, model, fofr0=FOFR0, nfofr=NFOFR, rcH0=RCH0) if use_sd else None
        pm = mgp.PM(N, N, box, omega=OMEGA, model=model_id, include_screening=1, grid_bytes=g, rank=rank, nranks=world,
                    device=local, nccl_id=nccl_id, deposit_mode=args.deposit_mode, sort_particles=args.sort_interval,
                    scale_dependent=use_sd)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    pm.set_pofk(64, 1, 1, 0.03, 2.0)          # paramfiles/additions_compute_pofk.txt
    t0 = time.time()
    A0 = 1.0 / (1.0 + Z_INIT)
    power = amplitude_table(N, box)
    if sd is not None:
        power = power * sd.pofk_ratio_by_k2()                    # input_pofk_is_for_lcdm = 1 (2LPT.c:396-397)
    pm.ic_generate(power, seed=5001)                             # displacement_fields() on the GPU(s)
    if sd is not None:
        for o in (1, 2):                                         # main.c:246-247 (UseCOLA = 1: Vel = 0, dDdy not needed)
            pm.assign_displacment_field_to_particles(0, o, sd.table(0, o, A0))
    pm.init_particles(cos.growth_D(A0), cos.growth_D2(A0))
    st = Stepper(pm, cos, model, box, sd, args.sd_mode)
    return pm, st, time.time() - t0


def pofk_compare(a, b, nmesh, box, what):
    """max over bins with k < k_Nyquist / 2 of |P_a - P_b| / P_b; a, b = (P, k, nmodes) of compute_power_spectrum."""
    pa, ka, na = a
    pb, kb, nb = b
    knyq = np.pi * nmesh / box
    sel = (nb > 0) & (kb < 0.5 * knyq) & (np.abs(pb) > 0)
    if not sel.any():
        return {"max_rel_err": None, "bins": 0, "against": what}
    return {"max_rel_err": float(np.max(np.abs(pa[sel] - pb[sel]) / np.abs(pb[sel]))), "bins": int(sel.sum()),
            "modes_equal": bool(np.array_equal(na, nb)), "k_max": float(kb[sel].max()), "against": what}


def measure(args, D, N, steps, warmup, want_e2e, engine_env=None):
    """The timed region at mesh N on all ranks.  Returns (record, P(k) of the last timed step)."""
    import torch
    pm, st, t_ic = make_context(args, N, D.rank, D.world, D.local, D.nccl_id(), env=engine_env)
    world, rank = D.world, D.rank
    use_sd = use_sd_of(args)
    npart_total = N ** 3
    stream = torch.cuda.ExternalStream(pm.stream)
    for _ in range(warmup):
        st.step()
    pm.launch_count(reset=True)
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        st.step()
    e1.record(stream)
    D.barrier()
    ms = e0.elapsed_time(e1)
    pk_last = pm.step_power_spectrum()
    launches = pm.launch_count()
    clocks = sampler.result() if rank == 0 else None
    ms = D.maxf([ms])[0]
    ms_per_step = ms / steps
    rec = {"nmesh": N, "ms_per_step": ms_per_step, "value": npart_total / (ms_per_step * 1e-3), "gpu_launches": int(launches),
           "clocks": clocks, "ic_seconds_gpu": round(t_ic, 2), "steps_taken": warmup + steps}

    # ---- per-phase CUDA-event timing (separate, synchronising pass) -> dominant own kernel
    pm.set_phase_timing(True)
    pm.phase_times(reset=True)
    nph = 3
    for _ in range(nph):
        st.step()
    rec["phases_ms"] = {k: round(v[0] / nph, 4) for k, v in pm.phase_times().items() if v[1]}
    pm.set_phase_timing(False)

    # ---- e2e: host particle buffers through the C ABI every step; every rank round-trips ITS particles through pinned
    # host buffers sized to its capacity (the count per rank changes with the migration); wall clock between barriers
    if want_e2e:
        try:
            cap = int(np.ceil(pm.local_np * N * N * (1.5 if world > 1 else 1.0))) + 64
            got = pm.download_particles()
            n0 = len(got["id"])
            keys = ("pos", "vel") if use_sd else ("pos", "vel", "D", "D2")   # SD: the fields are rebuilt on the device
            hp = {k: torch.empty((cap, 3), dtype=torch.float32).pin_memory() for k in keys}
            hid = torch.empty((cap,), dtype=torch.int64).pin_memory()
            for k in keys:
                hp[k][:n0] = torch.from_numpy(got[k])
            hid[:n0] = torch.from_numpy(got["id"].astype(np.int64))
            del got
            ne2e = max(2, min(steps, 5))
            h2d = d2h = 0
            t_up = t_st = t_dn = 0.0
            D.barrier()
            t0 = time.perf_counter()
            for _ in range(ne2e):
                n = pm.numpart
                h2d += n * (12 * len(keys) + 8)
                ta = time.perf_counter()
                pm.upload_raw(hp["pos"].data_ptr(), hp["vel"].data_ptr(), hp["D"].data_ptr() if "D" in hp else 0,
                              hp["D2"].data_ptr() if "D2" in hp else 0, hid.data_ptr(), n)
                tb = time.perf_counter()
                st.step()
                tc = time.perf_counter()
                pm.download_raw(hp["pos"].data_ptr(), hp["vel"].data_ptr(), 0, 0, hid.data_ptr())
                td = time.perf_counter()
                d2h += pm.numpart * 32
                t_up += tb - ta
                t_st += tc - tb
                t_dn += td - tc
            D.barrier()
            dt = (time.perf_counter() - t0) / ne2e
            dt, up, stp, dn = D.maxf([dt, t_up / ne2e, t_st / ne2e, t_dn / ne2e])
            bi, bo = D.sumf([h2d / ne2e, d2h / ne2e])
            rec["e2e"] = {"value": npart_total / dt, "unit": "particle-updates/s", "h2d_bytes_per_step": int(bi),
                          "d2h_bytes_per_step": int(bo), "ms_per_step": dt * 1e3, "steps": ne2e,
                          "upload_ms": up * 1e3, "step_ms": stp * 1e3, "download_ms": dn * 1e3,
                          "note": "every rank uploads Pos / Vel / ID of its own particles from pinned host memory and downloads "
                                  "them again (all ranks, summed bytes); wall clock between barriers, max over ranks"}
        except Exception as exc:       # the device-timed numbers above stand; say why there is no end-to-end one
            rec["e2e"] = None
            sys.stderr.write("e2e on %d ranks failed: %r\n" % (world, exc))
    pm.close()
    return rec, pk_last


def single_gpu_pofk(args, D, N, nsteps):
    """P(k) of step `nsteps` of the same workload on ONE GPU (rank 0's), for the full-size parity of a multi-rank run."""
    pm, st, _ = make_context(args, N, 0, 1, D.local, None)
    for _ in range(nsteps):
        st.step()
    pk = pm.step_power_spectrum()
    pm.close()
    return pk


def roofline_of(args, rec, world):
    N, g, model = rec["nmesh"], args.grid_bytes, args.model
    use_sd = use_sd_of(args)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))     # fallback of B200_PROFILING.md when the driver-written file is absent
    A_min = ALGO_BYTES[model](g) + (ALGO_BYTES_SD[args.sd_mode](g) if use_sd else 0)
    n_loc = N ** 3 / world
    # dominant hand-written kernels: deposit (PtoMesh phase) and gather (MtoParticles phase)
    kern_bytes = {"PtoMesh": (16 + g) * n_loc,                       # Pos(+id) 16 B read, grid g write per cell
                  "MtoParticles": (16 + 3 * g + 12) * n_loc}         # Pos 16 R, 3 grids R, Disp 12 W
    # whole-step roofline per GPU: HBM time of the algorithmic bytes + NVLink time of the slab transposes
    # (g * n/P * (P-1)/P bytes per distributed FFT and direction)
    nfft = {"lcdm": 4, "fofr": 6, "dgp": 6}[model] + ((6 if args.sd_mode == "merged" else 12) if use_sd else 0)
    t_hbm = A_min * n_loc / (peak * 1e9)
    nvl_bytes = nfft * g * n_loc * (world - 1) / world
    t_nvl = nvl_bytes / (NVLINK_GBS * 1e9)
    t_step = rec["ms_per_step"] * 1e-3
    step_roof = {"algorithmic_bytes_per_particle": A_min, "ffts_per_step": nfft, "achieved": A_min * n_loc / t_step / 1e9,
                 "hbm_ms": t_hbm * 1e3, "nvlink_bytes_per_gpu": nvl_bytes, "nvlink_ms": t_nvl * 1e3,
                 "nvlink_peak": "%g GB/s per direction (B200_PROFILING.md, measured peer copy)" % NVLINK_GBS,
                 "frac": (t_hbm + t_nvl) / t_step}
    phases = rec["phases_ms"]
    dom = max(("PtoMesh", "MtoParticles"), key=lambda k: phases.get(k, 0.0))
    ach = kern_bytes[dom] / (phases[dom] * 1e-3) / 1e9 if phases.get(dom) else None
    rows = args.deposit_mode == 3 and g == 8
    kname = {"PtoMesh": "k_deposit_tiles" if rows else ["k_deposit_atomic", "k_deposit_tile", "k_deposit_rowseg", "k_deposit_atomic"][args.deposit_mode],
             "MtoParticles": "k_gather_tiles" if rows else "k_gather"}[dom]
    traffic = None          # DRAM bytes per launch of that kernel from the committed ncu --set full capture of the same mesh
    try:
        if world == 1:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            traffic = tr.get("%s@%d" % (kname, N))
    except Exception:
        pass
    return {"bound": "hbm", "kernel": kname + {"PtoMesh": " (CIC deposit; the phase includes the -1 fill)", "MtoParticles": " (trilinear gather)"}[dom],
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None, "traffic": traffic,
            "note": "algorithmic bytes: deposit 16 + g per particle, gather 16 + 3 g + 12 (DESIGN.md section 4); the f64 deposit is bound "
                    "by the L2 reduction units (~3.5e11 f64 adds / s chip-wide, DESIGN.md section 5), not by HBM",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)",
            "step": step_roof, "phases_ms": phases}


def run_ours(args):
    D = Dist(args)
    rank, world = D.rank, D.world
    N = args.nmesh if args.nmesh else DEFAULT_NMESH
    rec, pk_last = measure(args, D, N, args.steps, args.warmup, not args.no_e2e)

    # ---- parity at the benched size on several ranks: P(k) of the last timed step against one GPU
    parity_full = None
    if world > 1 and not args.no_parity:
        pk1 = [None]
        if rank == 0:
            try:
                pk1[0] = single_gpu_pofk(args, D, N, args.warmup + args.steps)
            except Exception as exc:
                sys.stderr.write("1-GPU parity run failed: %r\n" % (exc,))
        D.barrier()
        if rank == 0 and pk1[0] is not None:
            parity_full = pofk_compare(pk_last, pk1[0], N, box_for(N),
                                       "the same workload (same seed, same %d steps) on ONE GPU (3-D cuFFT plans, no slabs), Npart=Nmesh=%d^3"
                                       % (args.warmup + args.steps, N))

    # ---- the north-star target on 8 GPUs: 1024^3
    target = None
    TARGET_NMESH = args.target_nmesh
    if world == args.target_gpus and not args.no_target and N != TARGET_NMESH:
        try:
            tsteps, twarm = min(args.steps, 10), min(args.warmup, 3)
            trec, tpk = measure(args, D, TARGET_NMESH, tsteps, twarm, False)
            # full-size parity: the same run with independent engines (NCCL all-to-all transposes around cuFFT 1-D plans
            # instead of the fused x-transform over peer memory); one GPU cannot hold 1024^3
            tpar = None
            if not args.no_parity:
                trec2, tpk2 = measure(args, D, TARGET_NMESH, tsteps, twarm, False, engine_env={"MGP_P2P": "0", "MGP_XFFT": "0"})
                if rank == 0:
                    tpar = pofk_compare(tpk, tpk2, TARGET_NMESH, box_for(TARGET_NMESH),
                                        "the same 8-rank run with independent transform engines (pack / NCCL all-to-all / unpack + "
                                        "cuFFT 1-D plans): one GPU cannot hold 1024^3 and the CPU reference cannot run it here"
                                        if TARGET_NMESH >= 1024 else "the same run with independent transform engines (pack / NCCL "
                                        "all-to-all / unpack + cuFFT 1-D plans)")
                    tpar["ms_per_step_other_engine"] = trec2["ms_per_step"]
            if rank == 0:
                roof = roofline_of(args, trec, world)
                target = {"config": config_of(args, TARGET_NMESH), "ms_per_step": trec["ms_per_step"], "value": trec["value"],
                          "steps": tsteps, "warmup": twarm, "roofline_frac_hbm_nvlink": roof["step"]["frac"], "roofline": roof,
                          "clocks": trec["clocks"], "parity_full_size": tpar,
                          "goal": "1024^3 f(R) + screening step on 8 GPUs at >= 0.60 of the HBM + NVLink roofline (BASELINE.json north_star)",
                          "n_gpus": world}
        except Exception as exc:
            sys.stderr.write("target run failed: %r\n" % (exc,))

    if rank != 0:
        return
    use_sd = use_sd_of(args)
    g = args.grid_bytes
    cfg = config_of(args, N)
    cfg.update({"sd_mode": args.sd_mode if use_sd else None, "deposit_mode": args.deposit_mode, "sort_interval": args.sort_interval,
                "sd_fields": ("merged per field type: D+D2 in one pass, 6 extra c2r/step" if args.sd_mode == "merged"
                              else "reference-structured: 4 fields, 12 extra c2r/step") if use_sd else None,
                "slab_transform": (None if world == 1 else
                                   ("2-D cuFFT + x-transform kernel fused with the exchange (stores / loads on peer memory over NVLink)"
                                    if (N & (N - 1)) == 0 and os.environ.get("MGP_XFFT", "1") != "0" and os.environ.get("MGP_P2P", "1") != "0"
                                    else "2-D cuFFT + transpose kernel storing into peer memory + 1-D cuFFT")),
                "ic_seconds_gpu": rec["ic_seconds_gpu"]})
    line = {"metric": METRIC, "value": rec["value"], "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 particles, f%d grids/FFTs, f64 weights" % (8 * g), "data": "synthetic",
            "config": cfg, "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"], "roofline": roofline_of(args, rec, world),
            "e2e": rec.get("e2e")}
    if not args.no_parity:
        try:
            par = parity_vs_reference(args)
        except Exception as exc:
            par = {"max_rel_err": None, "error": repr(exc)}
        if parity_full is not None:
            par["full_size"] = parity_full
        line["parity"] = par
    if target is not None:
        line["target"] = target
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(args, nmesh_full=N, steps=3, warmup=1)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ parity against the unmodified reference

def parity_vs_reference(args):
    """The reference's own C driver bound to the CUDA library (adapter/_build/MG_PICOLA_CUDA_<v>) and the unmodified CPU
    reference (oracle/_ref/MG_PICOLA_<v>[_mp]) on the SAME parameter file: every in-step P(k) file compared bin by bin for
    k < k_Nyquist / 2.  The files carry %10.5f, i.e. 1e-5 (Mpc/h)^3 absolute, which is below 1e-6 relative for every
    bin of this spectrum."""
    from oracle import mprun
    use_sd, variant = _ref_variant(args)
    gpu_exe = os.path.join(ROOT, "adapter", "_build", "MG_PICOLA_CUDA_%s" % variant)
    cpu_exe_mp, cpu_exe = mprun.exe_path(variant), os.path.join(ROOT, "oracle", "_ref", "MG_PICOLA_%s" % variant)
    if not os.path.exists(gpu_exe) or not (os.path.exists(cpu_exe_mp) or os.path.exists(cpu_exe)):
        return {"max_rel_err": None, "bins": 0, "against": "drop-in driver / oracle/_ref not built (need /root/reference at build time)"}
    N, nsteps = PARITY_NMESH, 10
    box = box_for(256) * N / 256.0
    outs = {}
    t0 = time.time()
    K = 1
    for kind in ("gpu", "cpu"):
        wd = tempfile.mkdtemp(prefix="mgp_parity_%s_" % kind)
        pf = write_paramfile(wd, N, box, args.model, nsteps, lcdm_growth=0 if use_sd else 1)
        env = dict(os.environ, MGP_SD_MERGED="1" if args.sd_mode == "merged" else "0", MGP_DEPOSIT_MODE=str(args.deposit_mode))
        if kind == "gpu":
            r = subprocess.run([gpu_exe, pf], capture_output=True, text=True, cwd=wd, timeout=600, env=env)
            if r.returncode != 0:
                raise RuntimeError("drop-in driver failed: " + (r.stdout[-500:] + r.stderr[-500:]))
        else:
            K = ref_ranks(N) if os.path.exists(cpu_exe_mp) else 1
            if K > 1:
                rc, so, errs = mprun.run([cpu_exe_mp, pf], K, scratch_mb=mprun.scratch_mb_for(N), timeout=600, cwd=wd)
                if rc != 0:
                    raise RuntimeError("reference on %d ranks failed (rc %s)" % (K, rc))
            else:
                r = subprocess.run([cpu_exe, pf], capture_output=True, text=True, cwd=wd, timeout=900)
                if r.returncode != 0:
                    raise RuntimeError("reference failed: " + r.stdout[-500:])
        outs[kind] = os.path.join(wd, "output")
    files = sorted(f for f in os.listdir(outs["cpu"]) if f.startswith("pofk_") and f.endswith("_CDM.txt"))
    knyq = np.pi * N / box
    worst, bins, nfiles = 0.0, 0, 0
    for f in files:
        a = np.loadtxt(os.path.join(outs["gpu"], f), comments="#").reshape(-1, 4)
        b = np.loadtxt(os.path.join(outs["cpu"], f), comments="#").reshape(-1, 4)
        if a.shape != b.shape or not np.array_equal(a[:, 0], b[:, 0]):
            raise RuntimeError("P(k) file %s: different bins" % f)
        sel = (b[:, 0] < 0.5 * knyq) & (np.abs(b[:, 1]) > 1e-3)
        if sel.any():
            worst = max(worst, float(np.max(np.abs(a[sel, 1] - b[sel, 1]) / np.abs(b[sel, 1]))))
            bins += int(sel.sum())
            nfiles += 1
    return {"max_rel_err": worst, "bins": bins, "pofk_files": nfiles, "nmesh": N, "steps": nsteps, "seconds": round(time.time() - t0, 1),
            "against": "the UNMODIFIED reference (oracle/_ref/MG_PICOLA_%s%s, %d rank%s of the host) on the same parameter file as the "
                       "reference's C driver bound to the CUDA library: %s%s, Npart=Nmesh=%d^3, Box=%g, Seed 5001, %d steps z=9->0, "
                       "every in-step P(k) file, bins with k < k_Nyquist/2"
                       % (variant, "_mp" if K > 1 else "", K, "s" if K > 1 else "", args.model,
                          " SCALEDEPENDENT (%s fields)" % args.sd_mode if use_sd else "", N, box, nsteps),
            "bound": "1e-4 (BASELINE.json north_star); merged two-order fields differ from the reference's two separately rounded "
                     "floats by one float32 ulp per particle and step"}


# ------------------------------------------------------------------ reference arm (CPU)

def write_paramfile(workdir, nmesh, box, model, nsteps, lcdm_growth=1, extra="", z_init=Z_INIT):
    """Parameter file for the reference build in oracle/_ref (tags: read_param.c:107-445,
    user_defined_functions.h:227-406)."""
    os.makedirs(os.path.join(workdir, "output"), exist_ok=True)
    t = np.load(os.path.join(ROOT, "tests", "golden", "input_power_spectrum.npz"))
    with open(os.path.join(workdir, "pk.dat"), "w") as f:
        for k, P in zip(t["k"], t["P"]):
            f.write("%16.8E %16.8E\n" % (k, P))
    with open(os.path.join(workdir, "out.dat"), "w") as f:
        f.write("0, %d\n" % nsteps)
    mg = {"fofr": "modified_gravity_active 1\nfofr0 %g\nnfofr %g\ninclude_screening 1\n" % (FOFR0, NFOFR),
          "lcdm": "modified_gravity_active 0\nfofr0 %g\nnfofr %g\ninclude_screening 0\n" % (FOFR0, NFOFR),
          "dgp": "modified_gravity_active 1\nrcH0_DGP %g\nRsmooth %g\ninclude_screening 1\n" % (RCH0, RSMOOTH)}[model]
    txt = mg + extra + """use_lcdm_growth_factors %d
input_pofk_is_for_lcdm 1
input_sigma8_is_for_lcdm 1
inverted_initial_condition 0
amplitude_fixed_initial_condition 0
OutputDir %s/output
FileBase bench
OutputRedshiftFile %s/out.dat
NumFilesWrittenInParallel 1
UseCOLA 1
Buffer 1.5
Nmesh %d
Nsample %d
Box %g
Init_Redshift %g
Seed 5001
SphereMode 0
WhichSpectrum 1
WhichTransfer 0
FileWithInputSpectrum %s/pk.dat
FileWithInputTransfer none
Omega %g
OmegaBaryon 0.049
HubbleParam 0.71
Sigma8 %g
PrimordialIndex 0.966
UnitLength_in_cm 3.085678e24
UnitMass_in_g 1.989e43
UnitVelocity_in_cm_per_s 1e5
InputSpectrum_UnitLength_in_cm 3.085678e24
pofk_compute_every_step 1
pofk_compute_rsd_pofk 0
pofk_nbins 64
pofk_bintype 1
pofk_subtract_shotnoise 1
pofk_kmin 0.03
pofk_kmax 2.0
""" % (lcdm_growth, workdir, workdir, nmesh, nmesh, box, z_init, workdir, OMEGA, SIGMA8)
    p = os.path.join(workdir, "param.txt")
    open(p, "w").write(txt)
    return p


def usable_cpus():
    """Logical CPUs this process may actually use: the affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def ref_ranks(nmesh):
    """Ranks of the multi-rank reference: a power of two, at most half the usable logical CPUs (SMT) and 16, and slabs of
    at least 16 planes (the reference aborts when more than Buffer - 1 of a rank's particles leave it in one step,
    auxPM.c:178-199: thin slabs do that)."""
    logical = usable_cpus()
    K = 1
    while K * 2 <= min(max(logical // 2, 2), nmesh // 16, 16):
        K *= 2
    return K if logical >= 2 else 1


def _ref_variant(args):
    use_sd = use_sd_of(args)
    if use_sd:
        return use_sd, {"fofr": "fofr", "dgp": "dgp_sd", "lcdm": "fofr"}[args.model]     # MODEL=FOFR / DGP -DSCALEDEPENDENT
    return use_sd, ("dgp" if args.model == "dgp" else "lcdm")  # 'lcdm' = -DFOFRGRAVITY without SCALEDEPENDENT (MODEL=FOFR_LCDM)


def cpu_reference(args, nmesh_full, steps, warmup, budget_s=150.0):
    """The reference's CPU path on the host's cores, as the multi-rank program it is: its unmodified driver (main.c) on K
    ranks = K processes of the multi-process MPI stand-in (shim_mpi_mp.c via oracle/mprun.py; bit-identical to the
    one-rank run, tests/test_ref_multirank.py), the workload's parameter file with the full 30-step schedule at a bounded
    mesh (min(Nmesh, --ref-nmesh), same cell size).  Iterations `warmup + 1 .. warmup + steps` are timed by the time stamps
    of the driver's "Iteration = i" lines (main.c:457; the driver flushes stdout in every phase of a step); the run is
    stopped after the last timed iteration or when the time budget is used up -- `steps` in the record is what was timed."""
    from oracle import mprun
    use_sd, variant = _ref_variant(args)
    N = min(nmesh_full, args.ref_nmesh)
    box = box_for(nmesh_full) * N / nmesh_full
    base = {"unit": "particle-updates/s", "kind": "reference", "nmesh_sample": N}
    if not mprun.available(variant):
        return dict(base, value=None, cores=0, sample="oracle/_ref not built (needs /root/reference at build time)")
    K = ref_ranks(N)
    want = min(warmup + steps, NSTEPS_RUN - 1)
    stamps = {}

    def on_line(line, t):
        m = re.match(r"Iteration = (\d+)", line)
        if m:
            stamps[int(m.group(1))] = t
            if int(m.group(1)) > want:
                return True
            if len(stamps) >= 3 and t > budget_s:                 # bounded: stop with what has been timed
                return True
        return False

    wd = tempfile.mkdtemp(prefix="mgp_refarm_")
    pf = write_paramfile(wd, N, box, args.model, NSTEPS_RUN, lcdm_growth=0 if use_sd else 1)
    rc, out, errs = mprun.run([mprun.exe_path(variant), pf], K, scratch_mb=mprun.scratch_mb_for(N), timeout=budget_s + 240,
                              line_cb=on_line)
    its = sorted(stamps)
    if rc != 0 or len(its) < 3:
        sys.stderr.write("reference run failed (rc %s): %s\n" % (rc, " | ".join(e[-200:] for e in errs if e.strip())))
        return dict(base, value=None, cores=K, sample="reference run failed (rc %s)" % rc)
    first = min(warmup + 1, its[-2])
    last = its[-1]                                                # the start of iteration `last` ends iteration last - 1
    timed = last - first
    dt = (stamps[last] - stamps[first]) / timed
    return dict(base, value=N ** 3 / dt, cores=K, ms_per_step=dt * 1e3, steps=timed, warmup=first - 1,
                sample="the unmodified reference driver (variant %s%s) on %d ranks of a multi-process MPI stand-in (the box has no MPI / "
                       "FFTW / GSL: shared-memory MPI, CPU FFT and mini-GSL stand-ins), the %s workload at Npart=Nmesh=%d^3 (Box=%g, same "
                       "cell size as the %d^3 workload), 30-step schedule; iterations %d..%d timed from the driver's own per-iteration "
                       "output; %d usable logical CPUs on the host"
                       % (variant, ", SCALEDEPENDENT: 12 extra c2r + 4 field assignments per step" if use_sd else "", K, args.model, N, box,
                          nmesh_full, first, last - 1, usable_cpus()))


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    nm = args.nmesh if args.nmesh else DEFAULT_NMESH
    cb = cpu_reference(args, nmesh_full=nm, steps=args.steps, warmup=args.warmup, budget_s=240.0)
    same = cb.get("nmesh_sample") == nm
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "particle-updates/s", "n_gpus": args.gpus,
            "steps": cb.get("steps", 0), "warmup": cb.get("warmup", 0), "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": cb.get("ms_per_step"), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 particles, f64 grids/FFTs", "data": "synthetic", "config": config_of(args, nm),
            "same_config": same, "sample_nmesh": cb.get("nmesh_sample"),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nmesh", type=int, default=0, help="0: 512 (strong scaling with --gpus)")
    ap.add_argument("--ref-nmesh", type=int, default=256, help="largest mesh the CPU reference is sampled at")
    ap.add_argument("--model", default="fofr", choices=["fofr", "dgp", "lcdm"])
    ap.add_argument("--grid-bytes", type=int, default=8, choices=[4, 8])
    ap.add_argument("--deposit-mode", type=int, default=0, help="0 atomic (warp-aggregated), 1 shared-memory tile, 2 deterministic, 3 bins + tiles (TMA)")
    ap.add_argument("--sort-interval", type=int, default=4, help="re-sort particles by cell every k-th step (0 never)")
    ap.add_argument("--scale-dependent", type=int, default=-1, help="-1: as the reference build of the model (fofr, dgp: 1; lcdm: 0)")
    ap.add_argument("--sd-mode", default="merged", choices=["merged", "ref"],
                    help="merged: D+D2 per field type in one pass; ref: the reference's four separate fields")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-target", action="store_true")
    ap.add_argument("--target-gpus", type=int, default=8, help="rank count at which the north-star target is measured as well")
    ap.add_argument("--target-nmesh", type=int, default=TARGET_NMESH)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
