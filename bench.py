#!/usr/bin/env python
"""Benchmark of the hot path: BASELINE.json metric "1D-EM forward evals/sec (= soundings x iters / s)".

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement on all host cores

Workload (config.workload): BASELINE.json configs[1] - 4096 synthetic RESOLVE FDEM soundings per GPU,
<= 30 layers, n_markov_chains = 10000, every chain run to the reference's own termination rule
(Inference1D.infer: N iterations, or N + burn-in + 1 once burned in).  One "step" = one pass of the fused
rjMCMC kernel over that whole batch.  unit: one accept_reject()+update() pair of one sounding.
Prints ONE JSON line (rank 0).

    python bench.py --workload skytem ...   # BASELINE configs[3] family: SkyTEM dual-moment time-domain soundings
                                            # (45 windows, skytem_options); not the default bench line
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

METRIC = "1D-EM forward evals/sec (=soundings x iters/s)"
UNIT = "evals/s"
SOUNDINGS_PER_GPU = 4096
N_MARKOV_CHAINS = 10000
SEED = 20261017


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--soundings", type=int, default=SOUNDINGS_PER_GPU, help="soundings per GPU")
    ap.add_argument("--chains", type=int, default=N_MARKOV_CHAINS, help="n_markov_chains")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--workload", default="resolve", choices=["resolve", "skytem", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class Workload:
    """What differs between the frequency-domain (RESOLVE) and time-domain (SkyTEM) benches."""

    def __init__(self, name):
        self.name = name
        self.tdem = name == "skytem"
        self.C = 45 if self.tdem else 12
        self.synth = dict(max_depth=400.0, n_channels=45) if self.tdem else {}

    def product(self, chains):
        from geobipy_b200 import ops
        if self.tdem:
            return ops.skytem_survey_struct(), ops.make_options(n_markov_chains=chains, **ops.SKYTEM_OPTIONS)
        return ops.resolve_system_struct(), ops.make_options(n_markov_chains=chains)

    def noise_std(self, clean, xp):
        """sqrt((5 % d)^2 + additive^2): RESOLVE 5 ppm; SkyTEM 2e-14 / 2e-13 V/Am^4 at 1 ms scaled by t^-1/2."""
        if not self.tdem:
            return xp.sqrt((0.05 * clean) ** 2 + 25.0)
        from geobipy_b200 import ops
        _, _, _, t = ops.tdem_window_operator(ops.skytem_survey_struct())
        add = np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / t)
        if xp is not np:
            add = xp.tensor(add, device=clean.device)
        return xp.sqrt((0.05 * clean) ** 2 + add ** 2)

    def oracle(self, chains):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        if self.tdem:
            return O, O.make_tdem_system(), O.skytem_options(n_markov_chains=chains), O.tdem_forward
        return O, O.make_system(), O.resolve_options(n_markov_chains=chains), O.fdem_forward


def workload_config(args, world):
    if args.workload == "mixed":
        return {"workload": "BASELINE configs[4] family: mixed flight line, %d soundings per GPU, even = RESOLVE FDEM, odd = SkyTEM TDEM, "
                            "n_markov_chains=%d" % (args.soundings, args.chains), "soundings_per_gpu": args.soundings,
                "soundings_total": args.soundings * world, "n_markov_chains": args.chains,
                "options": "resolve_options / skytem_options", "parallelism": "shard%d" % world}
    if args.workload == "skytem":
        nd = 1209
        return {
            "workload": "BASELINE configs[3] family: %d synthetic SkyTEM dual-moment TDEM soundings per GPU (26 + 19 windows), "
                        "<=30 layers, n_markov_chains=%d, chains run to the reference's termination rule" % (args.soundings, args.chains),
            "soundings_per_gpu": args.soundings, "soundings_total": args.soundings * world,
            "n_markov_chains": args.chains, "options": "skytem_options", "parallelism": "shard%d" % world,
            "forward_precision": "fp%d" % args.precision,
            "l2": "no flush needed: each step rewrites %.1f GB of posterior arrays per GPU (> 126 MB L2)"
                  % (args.soundings * (250 * nd * 4 + 2 * args.chains * 9) / 1e9),
        }
    return {
        "workload": "BASELINE configs[1]: %d synthetic RESOLVE FDEM soundings per GPU (6 freq, 12 channels), "
                    "<=30 layers, n_markov_chains=%d, chains run to the reference's termination rule" % (args.soundings, args.chains),
        "soundings_per_gpu": args.soundings, "soundings_total": args.soundings * world,
        "n_markov_chains": args.chains, "options": "resolve_options", "parallelism": "shard%d" % world,
        "forward_precision": "fp%d" % args.precision,
        "l2": "no flush needed: each step rewrites %.1f GB of posterior arrays per GPU (> 126 MB L2)"
              % (args.soundings * (250 * 440 * 4 + 2 * args.chains * 9) / 1e9),
    }


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_CACHE = {}


def _cpu_chain(job):
    idx, data, alt, chains, max_it, wname = job
    key = (wname, chains)
    if key not in _CPU_CACHE:  # per worker process: the time-domain tables take ~0.5 s to build
        _CPU_CACHE[key] = Workload(wname).oracle(chains)
    O, osys, oopt, _ = _CPU_CACHE[key]
    r = O.run_chain(osys, oopt, data, alt, SEED, idx, max_iterations=max_it)
    return float(r["scalars"][O.S_TOTAL_ITER])


def _observed_cpu(n, wname="resolve"):
    """Synthetic observed data of soundings 0..n-1 computed with the oracle forward (CPU arm only)."""
    from geobipy_b200.synthetic import synthetic_batch
    wl = Workload(wname)
    O, osys, _, fwd = wl.oracle(1)
    O.build()
    b = synthetic_batch(0, n, **wl.synth)
    data = np.zeros((n, wl.C))
    for i in range(n):
        L = int(b["nlayers"][i])
        clean = fwd(osys, b["height"][i], b["sigma"][i, :L], b["thickness"][i, :L])
        data[i] = clean + b["noise"][i] * wl.noise_std(clean, np)
    return data, b["height"]


def cpu_sample(chains, n_chains, max_it, pool, wname="resolve"):
    """Run n_chains chains of the workload on the host cores (one process per core); returns (iterations, seconds)."""
    data, alt = _observed_cpu(n_chains, wname)
    jobs = [(i, data[i], float(alt[i]), chains, max_it, wname) for i in range(n_chains)]
    t0 = time.perf_counter()
    its = pool.map(_cpu_chain, jobs)
    return float(sum(its)), time.perf_counter() - t0


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on the host CPU.  The reference itself is pure Python
    (NumPy/Numba) and cannot travel to the GPU box, so this arm times its plain-C restatement (oracle/,
    kind "port") with one process per host core - an upper bound on the reference's own speed (the
    Python reference measures ~115 evals/s/core in the build container)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_chains = cores
    max_it = 4000  # bounded sample: first 4000 iterations of `cores` chains per step (~2-4 s per step)
    kinds = ("resolve", "skytem") if args.workload == "mixed" else (args.workload,)
    for k in kinds:
        _observed_cpu(1, k)
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(args.warmup):
            for k in kinds:
                cpu_sample(args.chains, n_chains, 500, pool, k)
        its, secs = 0.0, 0.0
        for _ in range(args.steps):
            for k in kinds:  # a mixed line: the same number of soundings of each kind
                i, s = cpu_sample(args.chains, n_chains, max_it, pool, k)
                its += i
                secs += s
    value = its / secs
    sample = "%d chains (soundings 0..%d of the workload) x first %d iterations per step, one process per core" % (n_chains, n_chains - 1, max_it)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from geobipy_b200 import _lib, ops
    from geobipy_b200.parallel import gather_to_rank0, summarise_hitmap
    from geobipy_b200.synthetic import synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = Workload(args.workload)
    system, opt = wl.product(args.chains)
    B = args.soundings
    first = rank * B
    # synthetic observed data: true models -> fp64 forward on the GPU + N(0, (5% d)^2 + additive^2) noise
    sb = synthetic_batch(first, B, **wl.synth)
    t_sig = torch.tensor(sb["sigma"], device=dev)
    t_thk = torch.tensor(sb["thickness"], device=dev)
    t_nl = torch.tensor(sb["nlayers"], device=dev)
    t_alt = torch.tensor(sb["height"], device=dev)
    clean = ops.forward(system, t_nl, t_sig, t_thk, t_alt, precision=64)
    noise = torch.tensor(sb["noise"], device=dev)
    d_data = (clean + noise * wl.noise_std(clean, torch)).contiguous()
    torch.cuda.synchronize()

    outputs = ops.DEFAULT_OUTPUTS
    shapes = ops.chain_buffer_shapes(opt, B)
    tdt = {np.int32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}
    buffers = {n: torch.zeros(shapes[n][0], dtype=tdt[shapes[n][1]], device=dev) for n in outputs}
    iters_dev = torch.zeros((), dtype=torch.float64, device=dev)
    kernel_ms = []

    def step(i, count=True):
        r = ops.rjmcmc_run(system, opt, d_data, t_alt, seed=SEED + i, first_index=first, precision=args.precision,
                           outputs=outputs, buffers=buffers)
        if count:
            iters_dev.add_(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(1000 + i, count=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        res = step(i)
    ev1.record()
    barrier()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    iters = iters_dev.clone().reshape(1)
    last_kernel_ms = ops.last_kernel_ms()
    last_iters = float(res["scalars"][:, _lib.S_TOTAL_ITER].sum().item())
    n_spec = float(ops.debug_counters()[8]) / max(args.steps + args.warmup, 1)  # per launch (all launches are alike)
    n_fwd = float(res["scalars"][:, _lib.S_N_FORWARD].sum().item())
    n_sens = float(res["scalars"][:, _lib.S_N_SENS].sum().item())
    mean_k = float((res["ncells_hist"].sum(dim=0).double() * torch.arange(opt.max_layers + 1, device=dev)).sum().item()
                   / max(1.0, float(res["ncells_hist"].sum().item())))
    burned = float(res["scalars"][:, _lib.S_BURNED_IN].sum().item())
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(iters, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    total_iters = float(iters.item())
    value = total_iters / (total_ms * 1e-3)

    # end-of-run collation (BASELINE configs[2]): one gather of posterior summaries to rank 0 over NCCL
    gather_ms = None
    if world > 1:
        grids = ops.posterior_grids(opt, 1.0)
        ln_edges = torch.tensor(np.log(grids["sigma_edges"]), device=dev)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        summ = summarise_hitmap(res["hitmap"], ln_edges)
        summ = {k: (v + torch.log(res["scalars"][:, _lib.S_HALFSPACE]).unsqueeze(1)) for k, v in summ.items()}
        summ["edges_hist"] = res["edges_hist"]
        summ["scalars"] = res["scalars"]
        gather_to_rank0(summ, world * B)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)

    # e2e: the public host-buffer API, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h_data = d_data.cpu().numpy()
        h_alt = t_alt.cpu().numpy()
        hb = {n: torch.zeros(shapes[n][0], dtype=tdt[shapes[n][1]]).pin_memory().numpy() for n in outputs}
        h2d = h_data.nbytes + h_alt.nbytes
        d2h = sum(v.nbytes for v in hb.values())
        for b in buffers.values():
            b.resize_(0)  # free the device-resident result buffers: the host path allocates its own
        del buffers, res
        torch.cuda.empty_cache()
        n_e2e = max(1, min(args.steps, 2))
        barrier()
        t0 = time.perf_counter()
        e_iters = 0.0
        for i in range(n_e2e):
            r = ops.rjmcmc_run(system, opt, h_data, h_alt, seed=SEED + i, first_index=first, precision=args.precision,
                               device=local_rank, outputs=outputs, buffers=hb)
            e_iters += float(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        ei = torch.tensor([e_iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            dist.all_reduce(ei, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ei.item()) / float(el.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "api": "geobipy_b200.ops.rjmcmc_run(numpy) -> gbp_rjmcmc_run_host (pinned host buffers)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        nz = ops.n_depth(opt)
        # algorithmic HBM bytes per iteration = 8 N_z + 8 (L-1) + 8 (1 + 2 S) + 9, S = 1 system
        n_sys = 2 if wl.tdem else 1
        bytes_per_iter = 8.0 * nz + 8.0 * max(mean_k - 1.0, 0.0) + 8.0 * (1.0 + 2.0 * n_sys) + 9.0
        achieved = bytes_per_iter * last_iters / (last_kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                "dram_bytes_per_launch" if not wl.tdem else "dram_bytes_per_launch_skytem")
        except Exception:
            pass
        # secondary (the binding one): scalar fp32 issue.  flops per unit from gbp_flops_per_forward at the mean
        # layer count; a Jacobian pass counted as 3 forwards.
        fpf = ops.flops_per_forward(system, max(1, int(round(mean_k))))
        fwd_equiv = (n_fwd - n_sens) + 3.0 * n_sens
        flops = fpf * fwd_equiv
        # measured denominators for the two pipes that bound the path; a Jacobian pass ~1.3 forwards of MUFU
        fp32_peak, mufu_peak = ops.measure_peaks()
        mufu = ops.mufu_per_forward(system, max(1, int(round(mean_k)))) * ((n_fwd - n_sens) + 1.3 * n_sens)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "kernel": "gbp::rjmcmc_kernel<%s,%d,%s>" % ("float" if args.precision == 32 else "double", 48 if wl.tdem else 12, "TDEM" if wl.tdem else "FDEM"),
                         "kernel_ms": last_kernel_ms, "units_per_launch": last_iters, "bytes_per_unit": bytes_per_iter,
                         "peak_source": peak_src,
                         "note": "BASELINE.json asks for the HBM fraction; this path is bound by scalar FP/SFU issue and latency, not by HBM (SURVEY.md 8(d))"},
            "roofline_compute": {"bound": "fp32-issue", "achieved": flops / (last_kernel_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                 "flops_per_forward": fpf, "forward_equivalents_per_launch": fwd_equiv,
                                 "peak": fp32_peak, "peak_source": "measured on this device: 8 independent FFMA chains per thread (gbp_measure_peaks); nominal 148 SM x 128 FMA/clk x 1965 MHz = 74.4",
                                 "mufu_achieved_gops": mufu / (last_kernel_ms * 1e-3) / 1e9, "mufu_peak_gops": mufu_peak,
                                 "mufu_frac": mufu / (last_kernel_ms * 1e-3) / 1e9 / mufu_peak},
            "speculation": {"helpers_per_chain_max": int(os.environ.get("GBP_SPEC_HELPERS", "12")),
                            "iterations_committed_from_speculation": n_spec / last_iters,
                            "note": "idle warps evaluate future iterations of running chains; results bit-identical with it off (tests)"},
            "chain_stats": {"iterations_per_chain": last_iters / B, "mean_layers": mean_k, "forwards_per_iteration": n_fwd / last_iters,
                            "jacobians_per_iteration": n_sens / last_iters, "burned_in_fraction": burned / B},
        }
        line["roofline_compute"]["frac"] = line["roofline_compute"]["achieved"] / line["roofline_compute"]["peak"]
        if gather_ms is not None:
            line["gather_ms"] = gather_ms
        if world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            cores = os.cpu_count() or 1
            with mp.get_context("spawn").Pool(cores) as pool:
                cpu_sample(args.chains, cores, 200, pool, args.workload)
                its, secs = cpu_sample(args.chains, cores, 0 if not wl.tdem else 4000, pool, args.workload)
            line["cpu_baseline"] = {"value": its / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d chains (soundings 0..%d of the workload) %s, one process per core, C restatement of the reference (oracle/)" % (cores, cores - 1, "run to termination" if not wl.tdem else "x first 4000 iterations")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_mixed_arm(args):
    """BASELINE configs[4] family: a mixed flight line - even soundings RESOLVE FDEM, odd soundings SkyTEM TDEM -
    with posterior hitmaps and interface probabilities (edges histogram / sum, Inference2D.py:959-961) accumulated
    per sounding.  The two sampler kernels are both persistent (one CTA per SM), so a step runs them back to back on
    one stream.  Not the default bench line; prints the same JSON shape without the roofline block."""
    import torch
    import torch.distributed as dist
    from geobipy_b200 import _lib, ops
    from geobipy_b200.synthetic import synthetic_batch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.soundings
    tdt = {np.int32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}
    outputs = ("hitmap", "edges_hist", "ncells_hist", "scalars")
    parts = []
    for name, parity in (("resolve", 0), ("skytem", 1)):
        wl = Workload(name)
        system, opt = wl.product(args.chains)
        idx = np.arange(rank * B + parity, (rank + 1) * B, 2)       # global sounding indices of this kind on this rank
        sb = [synthetic_batch(int(i), 1, **wl.synth) for i in idx]
        cat = {k: np.concatenate([b[k] for b in sb]) for k in sb[0]}
        t = {k: torch.tensor(v, device=dev) for k, v in cat.items()}
        clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
        data = (clean + t["noise"] * wl.noise_std(clean, torch)).contiguous()
        shapes = ops.chain_buffer_shapes(opt, len(idx))
        bufs = {n: torch.zeros(shapes[n][0], dtype=tdt[shapes[n][1]], device=dev) for n in outputs}
        parts.append(dict(wl=wl, system=system, opt=opt, data=data, alt=t["height"], bufs=bufs, n=len(idx), first=int(idx[0])))
    iters_dev = torch.zeros((), dtype=torch.float64, device=dev)

    def step(i, count=True):
        out = []
        for p in parts:
            # sounding index of the random stream: position within this kind's list (streams stay sharding independent)
            r = ops.rjmcmc_run(p["system"], p["opt"], p["data"], p["alt"], seed=SEED + i, first_index=p["first"] // 2,
                               precision=args.precision, outputs=outputs, buffers=p["bufs"])
            if count:
                iters_dev.add_(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
            # interface probability per sounding (Inference2D.interface_probability)
            e = r["edges_hist"].to(torch.float64)
            out.append((r, e / e.sum(dim=1, keepdim=True).clamp_min(1.0)))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(1000 + i, count=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        res = step(i)
    ev1.record()
    barrier()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    iters = iters_dev.clone().reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(iters, op=dist.ReduceOp.SUM)
    # e2e through the host-pointer API (both kinds), one step
    h = [dict(data=p["data"].cpu().numpy(), alt=p["alt"].cpu().numpy()) for p in parts]
    h2d = sum(x["data"].nbytes + x["alt"].nbytes for x in h)
    for p in parts:
        for b in p["bufs"].values():
            b.resize_(0)
    del res
    torch.cuda.empty_cache()
    barrier()
    t0 = time.perf_counter()
    e_iters, d2h = 0.0, 0
    for p, x in zip(parts, h):
        r = ops.rjmcmc_run(p["system"], p["opt"], x["data"], x["alt"], seed=SEED, first_index=p["first"] // 2,
                           precision=args.precision, device=local_rank, outputs=outputs)
        e_iters += float(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
        d2h += sum(v.nbytes for v in r.values())
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    ei = torch.tensor([e_iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(ei, op=dist.ReduceOp.SUM)
    if rank == 0:
        total_ms = float(ms.item())
        line = {
            "metric": METRIC, "value": float(iters.item()) / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4] family: mixed flight line, %d soundings per GPU, even = RESOLVE FDEM, odd = SkyTEM TDEM, "
                                   "n_markov_chains=%d, posterior hitmaps + interface probabilities" % (B, args.chains),
                       "soundings_per_gpu": B, "soundings_total": B * world, "n_markov_chains": args.chains,
                       "options": "resolve_options / skytem_options", "parallelism": "shard%d" % world,
                       "l2": "no flush needed: each step rewrites GBs of posterior arrays per GPU (> 126 MB L2)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": float(ei.item()) / float(el.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": 1, "api": "geobipy_b200.ops.rjmcmc_run(numpy), one call per datapoint kind"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "mixed":
        run_mixed_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
