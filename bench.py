#!/usr/bin/env python
"""bench.py -- particle-updates/s of one full COLA particle-mesh step on B200.

A "step" = MoveParticles + PtoMesh (+ in-step P(k)) + ComputeFifthForce + Forces + MtoParticles +
Kick + Drift (main.c:474-592 of the reference) on synthetic Gaussian 2LPT initial conditions drawn
from the reference's bundled CAMB table (tests/golden/input_power_spectrum.npz), stepped along the
reference's COLA schedule (z_init = 9, linear in a).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--nmesh 256] [--model fofr|dgp|lcdm]
                  [--grid-bytes 8|4] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` = device-timed, particles resident in HBM.  `e2e` = the
same step driven through the C ABI with HOST particle buffers (pinned): upload of Pos/Vel/D/D2/ID
and download of Pos/Vel inside the timed region (on several ranks every rank round-trips its own slab's particles).  `roofline` = the dominant hand-written kernel
of the step, timed live with CUDA events on the library's stream.  `cpu_baseline` / `--impl
reference` = the unmodified reference compiled against the oracle stand-ins (oracle/_ref), one core.
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
# stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed when NCCL_DEBUG is set) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

OMEGA, SIGMA8, Z_INIT, NSTEPS_RUN = 0.267, 0.8, 9.0, 30
FOFR0, NFOFR, RCH0, RSMOOTH = 1e-5, 1.0, 1.0, 1.0
# compulsory HBM bytes per particle-step of a maximally fused step
WEAK_NMESH = {1: 256, 2: 320, 4: 400, 8: 512}
ALGO_BYTES = {"lcdm": lambda g: 120 + 32 * g, "fofr": lambda g: 120 + 51 * g, "dgp": lambda g: 120 + 51 * g}
# SCALEDEPENDENT add-on (section 8(d)): merged two-order fields 52 g, reference-structured (four fields) 48 + 100 g
ALGO_BYTES_SD = {"merged": lambda g: 52 * g, "ref": lambda g: 48 + 100 * g}


def box_for(nmesh):
    return 200.0 * nmesh / 256.0 if nmesh > 256 else 200.0     # keeps >= 0.78 Mpc/h cells like the example runs


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
    launches of a 4 ms step, which has two host synchronisations); `nvidia-smi` every 200 ms when NVML cannot be
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


def run_ours(args):
    import torch
    import mgpicola_b200 as mgp
    from mgpicola_b200 import cosmology

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [mgp.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]

    # weak scaling: the mesh side grows so that every GPU keeps ~256^3 = 16.8 M particles and cells
    # (1: 256, 2: 320, 4: 400, 8: 512; the slab decomposition needs a cubic mesh divisible by the rank count)
    N = args.nmesh if args.nmesh else WEAK_NMESH.get(world, 256)
    g = args.grid_bytes
    box = box_for(N)
    model = args.model
    cos = cosmology.LCDM(OMEGA, Z_INIT)
    model_id = {"lcdm": mgp.MODEL_NONE, "fofr": mgp.MODEL_FOFR, "dgp": mgp.MODEL_DGP}[model]
    use_sd = args.scale_dependent if args.scale_dependent >= 0 else int(model in ("fofr", "dgp"))
    sd = cosmology.ScaleDependentGrowth(cos, box, N, model, fofr0=FOFR0, nfofr=NFOFR, rcH0=RCH0) if use_sd else None
    pm = mgp.PM(N, N, box, omega=OMEGA, model=model_id, include_screening=1, grid_bytes=g, rank=rank, nranks=world,
                device=local, nccl_id=nccl_id, deposit_mode=args.deposit_mode, sort_particles=args.sort_interval,
                scale_dependent=use_sd)
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
    t_ic = time.time() - t0
    npart_total = N ** 3
    st = Stepper(pm, cos, model, box, sd, args.sd_mode)
    stream = torch.cuda.ExternalStream(pm.stream)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        st.step()
    pm.launch_count(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    dbg = []
    for _ in range(args.steps):
        t_dbg = time.perf_counter()
        st.step()
        dbg.append(time.perf_counter() - t_dbg)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if os.environ.get("MGP_BENCH_DEBUG") and rank == 0:
        sys.stderr.write("host ms per timed step: " + " ".join("%.2f" % (v * 1e3) for v in dbg) + " | device total %.2f\n" % ms)
    launches = pm.launch_count()
    clocks = sampler.result() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = npart_total / (ms_per_step * 1e-3)

    # ---- per-phase CUDA-event timing (separate, synchronising pass) -> dominant own kernel
    pm.set_phase_timing(True)
    pm.phase_times(reset=True)
    nph = 3
    for _ in range(nph):
        st.step()
    phases = {k: v[0] / nph for k, v in pm.phase_times().items() if v[1]}
    pm.set_phase_timing(False)

    # ---- e2e: host particle buffers through the C ABI every step
    e2e = None
    if world == 1 and not args.no_e2e:
        got = pm.download_particles()
        hp = {k: torch.from_numpy(got[k].copy()).pin_memory() for k in ("pos", "vel", "D", "D2")}
        hid = torch.from_numpy(got["id"].astype(np.int64)).pin_memory()
        if use_sd:      # the per-particle displacement fields are rebuilt on the device every step: only Pos, Vel, ID travel
            hp = {k: hp[k] for k in ("pos", "vel")}
        h2d = sum(t.numel() * t.element_size() for t in hp.values()) + hid.numel() * 8
        d2h = hp["pos"].numel() * 4 * 2
        ne2e = max(2, min(args.steps, 5))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ne2e):
            pm.upload_raw(hp["pos"].data_ptr(), hp["vel"].data_ptr(), hp["D"].data_ptr() if "D" in hp else 0,
                          hp["D2"].data_ptr() if "D2" in hp else 0, hid.data_ptr(), hid.numel())
            st.step()
            pm.download_raw(hp["pos"].data_ptr(), hp["vel"].data_ptr(), 0, 0, hid.data_ptr())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ne2e
        e2e = {"value": npart_total / dt, "unit": "particle-updates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h + hid.numel() * 8), "ms_per_step": dt * 1e3, "steps": ne2e}

    if world > 1 and not args.no_e2e:
        # every rank round-trips ITS particles through pinned host buffers sized to the rank's capacity (the count per
        # rank changes with the migration); wall clock between barriers, max over ranks
        try:
            import torch.distributed as dist
            cap = int(np.ceil(pm.local_np * N * N * 1.5)) + 64
            got = pm.download_particles()
            n0 = len(got["id"])
            keys = ("pos", "vel") if use_sd else ("pos", "vel", "D", "D2")
            hp = {k: torch.empty((cap, 3), dtype=torch.float32).pin_memory() for k in keys}
            hid = torch.empty((cap,), dtype=torch.int64).pin_memory()
            for k in keys:
                hp[k][:n0] = torch.from_numpy(got[k])
            hid[:n0] = torch.from_numpy(got["id"].astype(np.int64))
            del got
            ne2e = max(2, min(args.steps, 5))
            nbytes = 0
            barrier()
            t0 = time.perf_counter()
            for _ in range(ne2e):
                n = pm.numpart
                nbytes += n * (12 * len(keys) + 8)
                pm.upload_raw(hp["pos"].data_ptr(), hp["vel"].data_ptr(), hp["D"].data_ptr() if "D" in hp else 0,
                              hp["D2"].data_ptr() if "D2" in hp else 0, hid.data_ptr(), n)
                st.step()
                pm.download_raw(hp["pos"].data_ptr(), hp["vel"].data_ptr(), 0, 0, hid.data_ptr())
            barrier()
            dt = (time.perf_counter() - t0) / ne2e
            t = torch.tensor([dt, float(nbytes) / ne2e, float(pm.numpart) * 32.0], device="cuda", dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt = float(tmax[0].item())
            e2e = {"value": npart_total / dt, "unit": "particle-updates/s", "h2d_bytes_per_step": int(t[1].item()),
                   "d2h_bytes_per_step": int(t[2].item()), "ms_per_step": dt * 1e3, "steps": ne2e,
                   "note": "every rank uploads / downloads its own slab's particles (all ranks, summed bytes); max over ranks"}
        except Exception as exc:       # the device-timed numbers above stand; say why there is no end-to-end one
            e2e = None
            sys.stderr.write("e2e on %d ranks failed: %r\n" % (world, exc))

    if rank != 0:
        pm.close()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))     # fallback of B200_PROFILING.md when the driver-written file is absent
    A_min = ALGO_BYTES[model](g) + (ALGO_BYTES_SD[args.sd_mode](g) if use_sd else 0)
    # dominant hand-written kernels: deposit (PtoMesh phase) and gather (MtoParticles phase)
    kern_bytes = {"PtoMesh": (16 + g) * (N ** 3) / max(world, 1),            # Pos(+id) 16 B read, grid g write per cell
                  "MtoParticles": (16 + 3 * g + 12) * (N ** 3) / max(world, 1)}   # Pos 16 R, 3 grids R, Disp 12 W
    # whole-step roofline per GPU: HBM time of the algorithmic bytes + NVLink time of the slab transposes
    # (g * n/P * (P-1)/P bytes per distributed FFT and direction; 770 GB/s measured peer copy)
    nfft = {"lcdm": 4, "fofr": 6, "dgp": 6}[model] + ((6 if args.sd_mode == "merged" else 12) if use_sd else 0)
    n_loc = npart_total / world
    t_hbm = A_min * n_loc / (peak * 1e9)
    nvl_bytes = nfft * g * n_loc * (world - 1) / world
    t_nvl = nvl_bytes / 770e9
    step_roof = {"algorithmic_bytes_per_particle": A_min, "ffts_per_step": nfft,
                 "achieved": A_min * n_loc / (ms_per_step * 1e-3) / 1e9, "hbm_ms": t_hbm * 1e3,
                 "nvlink_bytes_per_gpu": nvl_bytes, "nvlink_ms": t_nvl * 1e3, "nvlink_peak": "770 GB/s per direction (B200_PROFILING.md, measured peer copy)",
                 "frac": (t_hbm + t_nvl) / (ms_per_step * 1e-3)}
    own = {k: phases.get(k, 0.0) for k in ("PtoMesh", "MtoParticles", "Forces", "ComputeFifthForce", "Kick", "Drift", "Sort", "Pofk", "SDField", "SDAssign")}
    dom = max(("PtoMesh", "MtoParticles"), key=lambda k: own.get(k, 0.0))
    ach = kern_bytes[dom] / (own[dom] * 1e-3) / 1e9 if own.get(dom) else None
    kname = {"PtoMesh": ["k_deposit_atomic", "k_deposit_tile", "k_deposit_rowseg", "k_deposit_scatter"][args.deposit_mode],
             "MtoParticles": "k_gather_rows" if args.deposit_mode == 3 else "k_gather"}[dom]
    traffic = None          # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same command, 256^3)
    try:
        if N == 256 and world == 1:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01b_traffic.json")))
            traffic = next((v for k, v in tr.items() if k.startswith(kname)), None)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": kname + {"PtoMesh": " (CIC deposit, incl. the fill kernel in the timed phase)", "MtoParticles": " (trilinear gather)"}[dom],
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None, "traffic": traffic,
            "note": "the deposit is bound by the L2 reduction unit, not by HBM: 8 RED.ADD.F64 per particle at ~0.72 cycles per lane and SM "
                    "= 0.29 ms of its 0.32 ms at 256^3 (DESIGN.md section 5); the whole step runs at roofline.step.frac of the copy bandwidth",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)",
            "step": step_roof,
            "phases_ms": {k: round(v, 4) for k, v in phases.items()}}
    line = {"metric": "particle-updates/sec per COLA PM step", "value": value, "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 particles, f%d grids/FFTs, f64 weights" % (8 * g), "data": "synthetic",
            "config": {"workload": "%s%s COLA step, Npart=Nmesh=%d^3, Box=%g Mpc/h, P(k) every step, z=9->0 in %d steps, %s"
                                   % (model, " with screening" if model != "lcdm" else "", N, box, NSTEPS_RUN,
                                      ("SCALEDEPENDENT growth (reference build MODEL=%s, use_lcdm_growth_factors=0), displacement fields %s"
                                       % ({"fofr": "FOFR", "dgp": "DGP -DSCALEDEPENDENT"}.get(model, model),
                                          "merged per field type: D+D2 in one pass, 6 extra c2r/step" if args.sd_mode == "merged"
                                          else "reference-structured: 4 fields, 12 extra c2r/step")) if use_sd
                                      else "scale-independent growth (reference build MODEL=FOFR_LCDM / DGP)"),
                       "nmesh": N, "npart": npart_total, "grid_bytes": g, "scale_dependent": use_sd, "sd_mode": args.sd_mode if use_sd else None, "deposit_mode": args.deposit_mode, "sort_interval": args.sort_interval,
                       "slab_transform": (None if world == 1 else
                                          ("2-D cuFFT + x-transform kernel fused with the exchange (stores / loads on peer memory over NVLink)"
                                           if (N & (N - 1)) == 0 and os.environ.get("MGP_XFFT", "1") != "0" and os.environ.get("MGP_P2P", "1") != "0"
                                           else "2-D cuFFT + transpose kernel storing into peer memory + 1-D cuFFT (Nmesh not a power of two)")),
                       "l2": "inputs larger than L2 (particles %.1f GB, grids %.1f GB each)" % (N ** 3 * 56 / 1e9, N ** 3 * g / 1e9),
                       "ic_seconds_gpu": round(t_ic, 2)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "e2e": e2e}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference(args, sample_nmesh=min(N, 128), steps=2, warmup=1)
    print(json.dumps(line), flush=True)
    pm.close()


# ------------------------------------------------------------------ reference arm (CPU)

def write_paramfile(workdir, nmesh, box, model, nsteps, lcdm_growth=1, extra=""):
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
""" % (lcdm_growth, workdir, workdir, nmesh, nmesh, box, Z_INIT, workdir, OMEGA, SIGMA8)
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


def _ref_variant(args):
    use_sd = args.scale_dependent if args.scale_dependent >= 0 else int(args.model in ("fofr", "dgp"))
    if use_sd:
        return use_sd, {"fofr": "fofr", "dgp": "dgp_sd", "lcdm": "fofr"}[args.model]     # MODEL=FOFR / DGP -DSCALEDEPENDENT
    return use_sd, ("dgp" if args.model == "dgp" else "lcdm")  # 'lcdm' = -DFOFRGRAVITY without SCALEDEPENDENT (MODEL=FOFR_LCDM)


def cpu_reference_ranks(args, sample_nmesh):
    """The reference as the multi-rank program it is: its unmodified driver (main.c) on K ranks = K processes of the
    multi-process MPI stand-in (oracle/shim/shim_mpi_mp.c, started by oracle/mprun.py; bit-identical to the one-rank
    run in double precision, tests/test_ref_multirank.py).  Time per step = difference of the reference's own
    "TimeStepping" timer (timer.c) between a 6-step and a 2-step run, which cancels the set-up and the extra force
    evaluation of an output interval.  None when it cannot run here."""
    import re
    import tempfile
    from oracle import mprun
    use_sd, variant = _ref_variant(args)
    if not mprun.available(variant):
        return None
    N = sample_nmesh
    logical = usable_cpus()
    K = 1
    while K * 2 <= min(max(logical // 2, 2), N // 4, 32):          # half the usable logical CPUs (SMT), a power of two dividing Nmesh
        K *= 2
    if K < 2 or logical < 2:
        return None
    nm_full = args.nmesh if args.nmesh else WEAK_NMESH.get(args.gpus, 256)
    box = box_for(nm_full) * N / nm_full
    tt = {}
    for nsteps in (2, 6):
        wd = tempfile.mkdtemp(prefix="mgp_refmp_")
        pf = write_paramfile(wd, N, box, args.model, nsteps, lcdm_growth=0 if use_sd else 1)
        # bounded: a 6-step run at 128^3 takes ~20 s on 4 ranks; a box that cannot give the ranks their cores falls back
        rc, out, errs = mprun.run([mprun.exe_path(variant), pf], K, scratch_mb=mprun.scratch_mb_for(N), timeout=90)
        m = re.search(r"TimeStepping\s+([0-9.]+)", out or "")
        if rc != 0 or not m:
            sys.stderr.write("multi-rank reference run failed (rc %s): %s\n" % (rc, " | ".join(e[-200:] for e in errs if e.strip())))
            return None
        tt[nsteps] = float(m.group(1))
    dt = (tt[6] - tt[2]) / 4.0
    if not dt > 0:
        return None
    return {"value": N ** 3 / dt, "unit": "particle-updates/s", "cores": K, "kind": "reference", "ms_per_step": dt * 1e3,
            "sample": "the unmodified reference driver (variant %s%s) on %d ranks of a multi-process MPI stand-in (no MPI / FFTW / GSL on "
                      "the box: shared-memory MPI, CPU FFT and mini-GSL stand-ins), %s workload at Npart=Nmesh=%d^3 (Box=%g); per-step time = "
                      "(TimeStepping of a 6-step run - TimeStepping of a 2-step run) / 4 from the reference's own timer; %d logical CPUs on the host"
                      % (variant, ", SCALEDEPENDENT: 12 extra c2r + 4 field assignments per step" if use_sd else "", K, args.model, N, box, logical)}


def cpu_reference(args, sample_nmesh, steps, warmup):
    """The reference's CPU path on the box's host cores: on several ranks when the multi-process MPI stand-in can run
    (cpu_reference_ranks), plus -- always -- the one-rank run stepped through its own GetDisplacements / Kick / Drift."""
    one = cpu_reference_one(args, sample_nmesh, steps, warmup)
    try:
        many = None if os.environ.get("MGP_BENCH_REF_RANKS", "1") == "0" else cpu_reference_ranks(args, sample_nmesh)
    except Exception as exc:
        sys.stderr.write("multi-rank reference run failed: %r\n" % (exc,))
        many = None
    if many is None or one.get("value") is None or many["value"] <= one["value"]:
        return one                                   # no ranks to be had (or no faster): the one-core figure stands
    many["one_core"] = {"value": one["value"], "ms_per_step": one.get("ms_per_step")}
    return many


def cpu_reference_one(args, sample_nmesh, steps, warmup):
    """Times the unmodified reference (oracle/_ref, built by oracle/Makefile) stepping the same
    workload on one host core: its own GetDisplacements / Kick / Drift, wall clock per step."""
    import tempfile
    from oracle import ref_lib
    use_sd, variant = _ref_variant(args)
    if not ref_lib.available(variant):
        return {"value": None, "unit": "particle-updates/s", "cores": 1, "kind": "reference",
                "sample": "oracle/_ref not built (needs /root/reference at build time)"}
    N = sample_nmesh
    nm_full = args.nmesh if args.nmesh else WEAK_NMESH.get(args.gpus, 256)
    box = box_for(nm_full) * N / nm_full
    wd = tempfile.mkdtemp(prefix="mgp_ref_")
    pf = write_paramfile(wd, N, box, args.model, NSTEPS_RUN, lcdm_growth=0 if use_sd else 1)
    drv = ref_lib.RefRun(variant, pf, quiet=True)
    for _ in range(warmup):
        drv.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        drv.step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": N ** 3 / dt, "unit": "particle-updates/s", "cores": 1, "kind": "reference",
            "ms_per_step": dt * 1e3,
            "sample": "%d timed + %d warm-up steps of the same %s workload at Npart=Nmesh=%d^3 (Box=%g), unmodified reference "
                      "sources (variant %s%s) on serial-MPI / CPU-FFT / mini-GSL stand-ins (no FFTW/MPI/GSL on the box), 1 core"
                      % (steps, warmup, args.model, N, box, variant, ", SCALEDEPENDENT: 12 extra c2r + 4 field assignments per step" if use_sd else "")}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    nm = args.nmesh if args.nmesh else WEAK_NMESH.get(args.gpus, 256)
    N = min(nm, args.ref_nmesh)
    cb = cpu_reference(args, sample_nmesh=N, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "particle-updates/sec per COLA PM step", "value": cb["value"],
            "unit": "particle-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb.get("ms_per_step"), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 particles, f64 grids/FFTs", "data": "synthetic",
            "config": {"workload": "%s COLA step, reference CPU path, bounded sample Npart=Nmesh=%d^3 of the %d^3 workload; %s"
                                   % (args.model, N, nm, cb["sample"]), "nmesh": N},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nmesh", type=int, default=0, help="0: 256 on one GPU, weak-scaled with --gpus")
    ap.add_argument("--ref-nmesh", type=int, default=128)
    ap.add_argument("--model", default="fofr", choices=["fofr", "dgp", "lcdm"])
    ap.add_argument("--grid-bytes", type=int, default=8, choices=[4, 8])
    ap.add_argument("--deposit-mode", type=int, default=0, help="0 atomic (warp-aggregated), 1 shared-memory tile, 2 deterministic")
    ap.add_argument("--sort-interval", type=int, default=4, help="re-sort particles by cell every k-th step (0 never)")
    ap.add_argument("--scale-dependent", type=int, default=-1, help="-1: as the reference build of the model (fofr, dgp: 1; lcdm: 0)")
    ap.add_argument("--sd-mode", default="merged", choices=["merged", "ref"],
                    help="merged: D+D2 per field type in one pass; ref: the reference's four separate fields")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
