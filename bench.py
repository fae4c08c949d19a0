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
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--soundings", type=int, default=SOUNDINGS_PER_GPU, help="soundings per GPU")
    ap.add_argument("--chains", type=int, default=N_MARKOV_CHAINS, help="n_markov_chains")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--workload", default="resolve", choices=["resolve", "skytem", "mixed", "tempest"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="streams consecutive steps alternate over (1 = one batch at a time)")
    ap.add_argument("--no-forward-only", action="store_true", help="skip the forward-only operator throughput block")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the C restatement even when baseline/_ref is importable")
    return ap.parse_args()


class Workload:
    """What differs between the frequency-domain (RESOLVE) and time-domain (SkyTEM) benches."""

    def __init__(self, name):
        self.name = name
        self.tempest = name == "tempest"
        self.tdem = name in ("skytem", "tempest")
        self.C = 30 if self.tempest else (45 if self.tdem else 12)
        self.synth = dict(max_depth=400.0, n_channels=self.C) if self.tdem else {}

    def altitude(self, h, xp=np):
        """Sensor / transmitter height of the synthetic soundings: the generator's 25-45 m, a fixed-wing Tempest at 120 m."""
        return h if not self.tempest else (xp.full_like(h, 120.0))

    def observe(self, clean, noise, xp):
        """Observed data from the clean response: + noise; a Tempest datapoint's data are secondary + primary field
        (Tempest_datapoint.py:107-118), noise = tempest_options' error model (0.1 %, the per-channel additive levels)."""
        if not self.tempest:
            return clean + noise * self.noise_std(clean, xp)
        from geobipy_b200 import ops
        prim = np.repeat(ops.tdem_primary_field(ops.tempest_survey_struct()), 15)
        add = np.asarray(ops.TEMPEST_ADDITIVE)
        if xp is not np:
            prim, add = xp.tensor(prim, device=clean.device), xp.tensor(add, device=clean.device)
        total = clean + prim
        return total + noise * xp.sqrt((0.001 * total) ** 2 + add ** 2)

    def product(self, chains):
        from geobipy_b200 import ops
        if self.tempest:
            return (ops.tempest_survey_struct(additive_level=ops.TEMPEST_ADDITIVE),
                    ops.make_options(**dict(ops.TEMPEST_OPTIONS, n_markov_chains=chains)))
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
        if self.tempest:
            return O, O.make_tempest_system(), O.tempest_options(n_markov_chains=chains), O.tdem_forward
        if self.tdem:
            return O, O.make_tdem_system(), O.skytem_options(n_markov_chains=chains), O.tdem_forward
        return O, O.make_system(), O.resolve_options(n_markov_chains=chains), O.fdem_forward


def workload_config(args, world):
    if args.workload == "mixed":
        return {"workload": "BASELINE configs[4] family: mixed flight line, %d soundings per GPU, even = RESOLVE FDEM, odd = SkyTEM TDEM, "
                            "n_markov_chains=%d" % (args.soundings, args.chains), "soundings_per_gpu": args.soundings,
                "soundings_total": args.soundings * world, "n_markov_chains": args.chains,
                "options": "resolve_options / skytem_options", "parallelism": "shard%d" % world}
    if args.workload == "tempest":
        return {
            "workload": "Tempest fixed-wing soundings (tempest_options; not in BASELINE.json): %d synthetic soundings per GPU, 15 X + 15 Z "
                        "windows of B field + primary field, <=30 layers, n_markov_chains=%d, chains run to the reference's termination rule"
                        % (args.soundings, args.chains),
            "soundings_per_gpu": args.soundings, "soundings_total": args.soundings * world, "n_markov_chains": args.chains,
            "options": "tempest_options", "parallelism": "shard%d" % world, "forward_precision": "fp%d" % args.precision,
            "l2": "no flush needed: each step rewrites GBs of posterior arrays per GPU (> 126 MB L2)",
            "pipelining": ("consecutive steps alternate over %d CUDA streams" % args.streams) if args.streams > 1 else "one batch at a time"}
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
            "pipelining": ("consecutive steps (independent batches) alternate over %d CUDA streams with separate result buffers: "
                           "a batch's tail of long chains overlaps the start of the next batch" % args.streams) if args.streams > 1
                          else "one batch at a time",
        }
    return {
        "workload": "BASELINE configs[1]: %d synthetic RESOLVE FDEM soundings per GPU (6 freq, 12 channels), "
                    "<=30 layers, n_markov_chains=%d, chains run to the reference's termination rule" % (args.soundings, args.chains),
        "soundings_per_gpu": args.soundings, "soundings_total": args.soundings * world,
        "n_markov_chains": args.chains, "options": "resolve_options", "parallelism": "shard%d" % world,
        "forward_precision": "fp%d" % args.precision,
        "l2": "no flush needed: each step rewrites %.1f GB of posterior arrays per GPU (> 126 MB L2)"
              % (args.soundings * (250 * 440 * 4 + 2 * args.chains * 9) / 1e9),
        "pipelining": ("consecutive steps (independent batches) alternate over %d CUDA streams with separate result buffers: "
                       "a batch's tail of long chains overlaps the start of the next batch" % args.streams) if args.streams > 1
                      else "one batch at a time",
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
        h = float(wl.altitude(b["height"][i:i + 1])[0])
        if wl.tempest:   # the forward tables only (the sampler system adds the primary field itself)
            clean = fwd(O.make_tdem_system([O.tempest_definition()], rx_offset=(-107.0, 0.0, -45.0)), h, b["sigma"][i, :L], b["thickness"][i, :L])
        else:
            clean = fwd(osys, h, b["sigma"][i, :L], b["thickness"][i, :L])
        data[i] = wl.observe(clean, b["noise"][i], np)
    return data, wl.altitude(b["height"])


def cpu_sample(chains, n_chains, max_it, pool, wname="resolve"):
    """Run n_chains chains of the workload on the host cores (one process per core); returns (iterations, seconds)."""
    data, alt = _observed_cpu(n_chains, wname)
    jobs = [(i, data[i], float(alt[i]), chains, max_it, wname) for i in range(n_chains)]
    t0 = time.perf_counter()
    its = pool.map(_cpu_chain, jobs)
    return float(sum(its)), time.perf_counter() - t0


def _ref_live():
    """oracle/ref_live.py when an importable copy of the reference exists on this machine (baseline/_ref), else None."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_live
    except Exception:
        return None
    return ref_live if ref_live.reference_root() else None


def _ref_chain(job):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_live
    return ref_live.time_chain(job)


def reference_sample(pool, n_chains, iters, seed=1):
    """The UNMODIFIED reference (baseline/_ref: Inference1D.accept_reject + update, Inference1D.py:537, :705, its own
    numba forward kernels) on soundings 0..n_chains-1 of the RESOLVE workload, `iters` iterations each, one process per
    chain.  Returns (iterations, wall seconds of the pool.map, max of the per-chain loop times)."""
    data, alt = _observed_cpu(n_chains, "resolve")
    jobs = [(i, data[i], float(alt[i]), N_MARKOV_CHAINS, iters, seed) for i in range(n_chains)]
    t0 = time.perf_counter()
    out = pool.map(_ref_chain, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return float(sum(o[0] for o in out)), wall, max(o[1] for o in out)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.

    RESOLVE workload with baseline/_ref present (the offline `pip install --no-deps --target baseline/_ref` of the
    reference; numba is in the image): the UNMODIFIED Python reference, one process per core, each step a bounded
    sample of `cores` soundings x 400 iterations (kind "reference"; import + numba compilation happen in the warm-up).
    Otherwise (time-domain workloads: the reference's gatdaem1d is absent; or no baseline/_ref): the plain-C
    restatement (oracle/, kind "port") - ~18x faster per core than the Python reference, i.e. an upper bound."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[k] = "1"   # one process per core already
    cores = os.cpu_count() or 1
    n_chains = cores
    kinds = ("resolve", "skytem") if args.workload == "mixed" else (args.workload,)
    for k in kinds:
        _observed_cpu(1, k)
    live = _ref_live() if args.workload == "resolve" and not args.port else None
    port = None
    with mp.get_context("spawn").Pool(cores) as pool:
        if live is not None:
            ref_iters = 400
            try:
                for w in range(max(args.warmup, 1)):   # the first one imports the reference and compiles its kernels
                    reference_sample(pool, n_chains, 50)
                its, secs = 0.0, 0.0
                for _ in range(args.steps):
                    i, wall, _ = reference_sample(pool, n_chains, ref_iters)
                    its += i
                    secs += wall
                kind = "reference"
                sample = ("%d soundings (0..%d of the workload) x first %d iterations per step, one process per core: the unmodified "
                          "reference from baseline/_ref (Inference1D.accept_reject + update, numba forward kernels)" % (n_chains, n_chains - 1, ref_iters))
            except Exception as e:   # the reference failed to import / run on this box: fall back to the port, say so
                sys.stderr.write("reference arm: live reference failed (%r); timing the C port instead\n" % (e,))
                live = None
        # the C restatement: the arm itself when the reference cannot run, a side figure otherwise
        max_it = 4000  # bounded sample: first 4000 iterations of `cores` chains per step (~2-4 s per step)
        p_its, p_secs = 0.0, 0.0
        for _ in range(args.warmup if live is None else 1):
            for k in kinds:
                cpu_sample(args.chains, n_chains, 500, pool, k)
        for _ in range(args.steps if live is None else 1):
            for k in kinds:  # a mixed line: the same number of soundings of each kind
                i, s = cpu_sample(args.chains, n_chains, max_it, pool, k)
                p_its += i
                p_secs += s
        port = {"value": p_its / p_secs, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d chains (soundings 0..%d of the workload) x first %d iterations per step, one process per core, C restatement (oracle/)" % (n_chains, n_chains - 1, max_it)}
        if live is None:
            its, secs, kind, sample = p_its, p_secs, "port", port["sample"]
    value = its / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if kind == "reference":
        line["cpu_port"] = port
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines, self.first = gpu_index, None, [], 0

    def mark(self):
        """The timed region starts here: only samples taken from now on count.  (nvidia-smi is started BEFORE the warm-up:
        its start-up holds a driver lock for a second or two, which stalled the first timed launch when it ran inside.)"""
        self.first = len(self.lines)

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
        for ln in self.lines[self.first:]:
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
def forward_only_block(ops, system, dev, precision, B=262144):
    """Throughput of the standalone forward / forward+Jacobian operators (gbp_fdem_forward / gbp_fdem_sensitivity, or the
    time-domain twins) at L = 3, 10, 30 layers: B soundings per launch resident in HBM, kernel time from the library's
    CUDA events on the launching stream (best of 3 after one warm-up), with the SURVEY 8(d) flop and special-function
    counts against the peaks measured on this device.  "10-layer RESOLVE" of the north-star target is the L = 10 row."""
    import torch
    rng = np.random.default_rng(99)
    fp32_peak, mufu_peak = ops.measure_peaks()
    out = {"soundings_per_launch": B, "precision": "fp%d" % precision, "rows": []}
    sig = torch.tensor(10.0 ** rng.uniform(-3.0, 0.0, (B, 30)), device=dev)
    thk = torch.tensor(np.exp(rng.uniform(np.log(1.0), np.log(20.0), (B, 30))), device=dev)
    alt = torch.tensor(rng.uniform(25.0, 45.0, B), device=dev)
    for L in (3, 10, 30):
        nl = torch.full((B,), L, dtype=torch.int32, device=dev)
        row = {"layers": L}
        for sens in (False, True):
            best = None
            for rep in range(4):
                ops.forward(system, nl, sig, thk, alt, precision=precision, sensitivity=sens)
                torch.cuda.synchronize(dev)
                if rep:
                    best = ops.last_kernel_ms() if best is None else min(best, ops.last_kernel_ms())
            key = "forward_jacobian" if sens else "forward"
            per_s = B / (best * 1e-3)
            row[key + "_per_s"] = per_s
            row[key + "_ms"] = best
            if not sens:
                row["fp32_frac"] = ops.flops_per_forward(system, L) * per_s / 1e12 / fp32_peak
                row["mufu_frac"] = ops.mufu_per_forward(system, L) * per_s / 1e9 / mufu_peak
        out["rows"].append(row)
    out["note"] = ("flops per forward = filter points x (75 L + 39) (SURVEY 8(d)); the sampler's chains hold 2-3 layers on these data "
                   "(the reference's own chains: mean 2.1), so its evals/s sit between the L = 3 forward and forward+Jacobian rows")
    return out


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from geobipy_b200 import _lib, ops
    from geobipy_b200.parallel import Collator
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
    t_alt = wl.altitude(torch.tensor(sb["height"], device=dev), torch)
    clean = ops.forward(system, t_nl, t_sig, t_thk, t_alt, precision=64)
    noise = torch.tensor(sb["noise"], device=dev)
    d_data = wl.observe(clean, noise, torch).contiguous()
    torch.cuda.synchronize()

    outputs = ops.DEFAULT_OUTPUTS
    shapes = ops.chain_buffer_shapes(opt, B)
    tdt = {np.int32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}
    # Consecutive steps (independent batches, as the flight lines of a survey are) go to alternating streams, each with
    # its own result buffers: the persistent kernel of step i+1 gets the SMs one by one as the CTAs of step i run out of
    # chains, so the tail of a batch - a few long chains on a few SMs - overlaps the bulk of the next.  --streams 1 =
    # strictly one batch at a time.
    ns = max(1, args.streams)
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(ns)] if ns > 1 else [main]
    bufsets = [{n: torch.zeros(shapes[n][0], dtype=tdt[shapes[n][1]], device=dev) for n in outputs} for _ in range(ns)]
    iters_part = [torch.zeros((), dtype=torch.float64, device=dev) for _ in range(ns)]
    grids = ops.posterior_grids(opt, 1.0)
    ln_edges = np.log(grids["sigma_edges"])
    sig_lo = torch.full((B,), float(ln_edges[0]), dtype=torch.float64, device=dev)   # bins relative to each sounding's half-space
    sig_dx = float(ln_edges[-1] - ln_edges[0]) / (ln_edges.size - 1)
    k_ev, c_ev = [], []   # CUDA events on the launching stream: sampler kernel / collation of every timed step
    collators = [None] * ns

    def collate(r, j):
        """End-of-run collation of a step (BASELINE configs[2]): posterior summaries on the device
        (gbp_summarise_hitmap), then ONE gather of the packed per-sounding rows to rank 0 over NCCL."""
        # (bin edges resident on the device: nothing here copies from the host or waits for the stream)
        mean, pct = ops.summarise_hitmap(r["hitmap"], sig_lo, sig_dx, (5.0, 50.0, 95.0))
        ln_ref = torch.log(r["scalars"][:, _lib.S_HALFSPACE]).unsqueeze(1)
        summ = {"mean": mean + ln_ref, "p5": pct[0] + ln_ref, "p50": pct[1] + ln_ref, "p95": pct[2] + ln_ref}
        summ["edges_hist"] = r["edges_hist"]
        summ["scalars"] = r["scalars"]
        if world > 1:
            if collators[j] is None:   # preallocated send / receive buffers; warm_up() creates the NCCL p2p channels
                collators[j] = Collator(summ, world * B)
                collators[j].warm_up()
            return collators[j].gather(summ)
        return summ

    def step(i, count=True, j=0):
        with torch.cuda.stream(streams[j]):
            if count:
                k_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                c_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                k_ev[-1][0].record()
            r = ops.rjmcmc_run(system, opt, d_data, t_alt, seed=SEED + i, first_index=first, precision=args.precision,
                               outputs=outputs, buffers=bufsets[j])
            if count:
                k_ev[-1][1].record()
                iters_part[j].add_(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
                c_ev[-1][0].record()
            collate(r, j)
            if count:
                c_ev[-1][1].record()
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # warm-up: one batch at a time (also builds the Collators and warms their channels); its launches, timed alone with
    # events on their stream, give the duration of ONE launch that does not share the machine (roofline.kernel_ms_alone)
    for i in range(max(args.warmup, ns)):
        torch.cuda.synchronize()
        # (exactly the code path of a timed step: a kernel used for the first time inside the timed region is loaded
        # lazily, and loading waits for the running persistent kernel to finish - that serialised the first two launches)
        r = step(1000 + i, count=True, j=i % ns)
    torch.cuda.synchronize()
    alone_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev[-2:]]))
    # the collation on its own: every rank's results ready (barrier), buffers and channels warm.  Inside the timed region
    # a rank's gather also waits for the slowest rank's batch to end - that skew is in ms_per_step, not in the collective.
    jl = (max(args.warmup, ns) - 1) % ns
    ce = []
    for _ in range(3):
        barrier()
        with torch.cuda.stream(streams[jl]):
            ce.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            ce[-1][0].record()
            collate(r, jl)
            ce[-1][1].record()
    torch.cuda.synchronize()
    collate_alone = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ce[1:]]))], dtype=torch.float64, device=dev)
    k_ev.clear()
    c_ev.clear()
    for x in iters_part:
        x.zero_()
    barrier()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    ev0.record(main)
    for st in streams:
        st.wait_event(ev0)
    host_t = [time.perf_counter()]
    for i in range(args.steps):
        res = step(i, j=i % ns)
        host_t.append(time.perf_counter())
    for st in streams:
        main.wait_stream(st)
    ev1.record(main)
    barrier()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    iters_dev = torch.stack(iters_part).sum()
    iters = iters_dev.clone().reshape(1)
    # the dominant kernel: timed region / launches (launches overlap when ns > 1) and mean units over the timed launches
    mean_kernel_ms = float(ms.item()) / max(args.steps, 1) if ns > 1 else float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    if os.environ.get("GBP_BENCH_TRACE"):   # where the step's time goes, on the streams' own clocks
        tl = [("k%d" % i, ev0.elapsed_time(a), ev0.elapsed_time(b)) for i, (a, b) in enumerate(k_ev)]
        tl += [("c%d" % i, ev0.elapsed_time(a), ev0.elapsed_time(b)) for i, (a, b) in enumerate(c_ev)]
        sys.stderr.write("trace rank %d: %s total %.1f host enqueue done at %s ms\n" % (
            rank, sorted(tl, key=lambda x: x[1]), ev0.elapsed_time(ev1), [round(1e3 * (x - host_t[0]), 1) for x in host_t[1:]]))
    mean_iters = float(iters_dev.item()) / max(args.steps, 1)
    collate_ms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in c_ev]))], dtype=torch.float64, device=dev)
    last_iters = float(res["scalars"][:, _lib.S_TOTAL_ITER].sum().item())
    n_spec = float(ops.debug_counters()[8]) / max(args.steps + max(args.warmup, ns), 1)  # per launch (all launches are alike)
    n_fwd = float(res["scalars"][:, _lib.S_N_FORWARD].sum().item())
    n_sens = float(res["scalars"][:, _lib.S_N_SENS].sum().item())
    mean_k = float((res["ncells_hist"].sum(dim=0).double() * torch.arange(opt.max_layers + 1, device=dev)).sum().item()
                   / max(1.0, float(res["ncells_hist"].sum().item())))
    burned = float(res["scalars"][:, _lib.S_BURNED_IN].sum().item())
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(iters, op=dist.ReduceOp.SUM)
        dist.all_reduce(collate_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(collate_alone, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    total_iters = float(iters.item())
    value = total_iters / (total_ms * 1e-3)
    gather_ms = float(collate_alone.item())          # the collation's own duration (summaries + pack + one gather), warm
    gather_span_ms = float(collate_ms.item())       # first to last event of it inside the timed region: with steps overlapping,
                                                    # its small kernels wait for SMs the next batch's persistent kernel holds
    gather_bytes = collators[0].bytes_per_rank if collators[0] is not None else 0

    # forward-only operators ("report also raw kernel forwards/s for the forward-only op")
    forward_only = None
    if rank == 0 and not args.no_forward_only:
        forward_only = forward_only_block(ops, system, dev, args.precision)

    # e2e: the public host-buffer API, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h_data = d_data.cpu().numpy()
        h_alt = t_alt.cpu().numpy()
        ne = max(1, min(ns, 2))   # calls in flight: the library keeps two sets of device buffers, each with its own stream
        hbs = [{n: torch.zeros(shapes[n][0], dtype=tdt[shapes[n][1]]).pin_memory().numpy() for n in outputs} for _ in range(ne)]
        h2d = h_data.nbytes + h_alt.nbytes
        d2h = sum(v.nbytes for v in hbs[0].values())
        for bs in bufsets:
            for b in bs.values():
                b.resize_(0)  # free the device-resident result buffers: the host path allocates its own
        del bufsets, res, r
        torch.cuda.empty_cache()
        # two calls in flight: at least 4 steps, so that the pipeline is not all ramp (as many as the device-timed region, up to 8)
        n_e2e = max(ne, min(args.steps, 4)) if ne == 1 else max(4, min(args.steps, 8))

        def host_step(i):
            rr = ops.rjmcmc_run(system, opt, h_data, h_alt, seed=SEED + i, first_index=first, precision=args.precision,
                                device=local_rank, outputs=outputs, buffers=hbs[i % ne])
            return float(rr["scalars"][:, _lib.S_TOTAL_ITER].sum())
        from concurrent.futures import ThreadPoolExecutor
        barrier()
        t0 = time.perf_counter()
        if ne > 1:   # two host threads, as a survey driver feeding consecutive flight lines would: ctypes drops the GIL
            with ThreadPoolExecutor(ne) as ex:
                # step i reuses the host buffers of step i - ne: submit it only when that one is done
                futs = []
                for i in range(n_e2e):
                    if i >= ne:
                        futs[i - ne].result()
                    futs.append(ex.submit(host_step, i))
                e_iters = sum(f.result() for f in futs)
        else:
            e_iters = sum(host_step(i) for i in range(n_e2e))
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        ei = torch.tensor([e_iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            dist.all_reduce(ei, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ei.item()) / float(el.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e, "calls_in_flight": ne,
               "api": "geobipy_b200.ops.rjmcmc_run(numpy) -> gbp_rjmcmc_run_host (pinned host buffers)"
                      + (", consecutive steps issued from %d host threads" % ne if ne > 1 else "")}

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
        achieved = bytes_per_iter * mean_iters / (mean_kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                "dram_bytes_per_launch" if not wl.tdem else ("dram_bytes_per_launch_skytem" if not wl.tempest else "dram_bytes_per_launch_tempest"))
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
                         "traffic": traffic, "kernel": "gbp::rjmcmc_kernel<%s,%d,%s>" % ("float" if args.precision == 32 else "double", 48 if wl.tdem else 12, "TEMPEST" if wl.tempest else ("TDEM" if wl.tdem else "FDEM")),
                         "kernel_ms": mean_kernel_ms, "units_per_launch": mean_iters, "bytes_per_unit": bytes_per_iter,
                         "kernel_ms_is": ("timed region / %d launches (CUDA events; consecutive launches overlap on %d streams)" % (args.steps, ns)) if ns > 1
                                         else "mean over the %d timed launches (CUDA events on the launching stream)" % args.steps,
                         "kernel_ms_alone": alone_ms,
                         "peak_source": peak_src,
                         "note": "BASELINE.json asks for the HBM fraction; this path is bound by scalar FP/SFU issue and latency, not by HBM (SURVEY.md 8(d))"},
            "roofline_compute": {"bound": "fp32-issue", "achieved": flops * (mean_iters / last_iters) / (mean_kernel_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                 "flops_per_forward": fpf, "forward_equivalents_per_launch": fwd_equiv,
                                 "peak": fp32_peak, "peak_source": "measured on this device: 8 independent FFMA chains per thread (gbp_measure_peaks); nominal 148 SM x 128 FMA/clk x 1965 MHz = 74.4",
                                 "mufu_achieved_gops": mufu * (mean_iters / last_iters) / (mean_kernel_ms * 1e-3) / 1e9, "mufu_peak_gops": mufu_peak,
                                 "mufu_frac": mufu * (mean_iters / last_iters) / (mean_kernel_ms * 1e-3) / 1e9 / mufu_peak},
            "speculation": {"helpers_per_chain_max": int(os.environ.get("GBP_SPEC_HELPERS", "12")),
                            "iterations_committed_from_speculation": n_spec / last_iters,
                            "note": "idle warps evaluate future iterations of running chains; results bit-identical with it off (tests)"},
            "chain_stats": {"iterations_per_chain": last_iters / B, "mean_layers": mean_k, "forwards_per_iteration": n_fwd / last_iters,
                            "jacobians_per_iteration": n_sens / last_iters, "burned_in_fraction": burned / B},
        }
        line["roofline_compute"]["frac"] = line["roofline_compute"]["achieved"] / line["roofline_compute"]["peak"]
        line["collation"] = {"ms_per_step": gather_ms, "ms_is": "duration of one collation with every rank ready (barrier before it; max over ranks, warm channels)",
                             "span_in_timed_region_ms": gather_span_ms, "inside_timed_region": True, "bytes_per_rank": int(gather_bytes),
                             "what": "gbp_summarise_hitmap (mean, p5/p50/p95 per depth cell) + edges histogram + scalars"
                                     + (", packed and gathered to rank 0 with one NCCL collective (warm channels)" if world > 1 else "")}
        if world > 1:
            line["gather_ms"] = gather_ms
        if forward_only is not None:
            line["forward_only"] = forward_only
        if world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
                os.environ[k] = "1"
            cores = os.cpu_count() or 1
            with mp.get_context("spawn").Pool(cores) as pool:
                cpu_sample(args.chains, cores, 200, pool, args.workload)
                its, secs = cpu_sample(args.chains, cores, 0 if not wl.tdem else 4000, pool, args.workload)
                line["cpu_baseline"] = {"value": its / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": "%d chains (soundings 0..%d of the workload) %s, one process per core, C restatement of the reference (oracle/)" % (cores, cores - 1, "run to termination" if not wl.tdem else "x first 4000 iterations")}
                # the reference's OWN NumPy/Numba path on this box's host cores, in the same run
                if _ref_live() is not None and not wl.tdem:
                    try:
                        reference_sample(pool, cores, 30)                 # import + numba compilation in every worker
                        i1, _, t1 = reference_sample(pool, 1, 1500)       # one core
                        ia, wa, _ = reference_sample(pool, cores, 400)    # all cores
                        line["cpu_baseline_reference"] = {
                            "kind": "reference", "unit": UNIT, "one_core": i1 / t1, "all_cores": ia / wa, "cores": cores,
                            "sample": "unmodified reference from baseline/_ref (Inference1D.accept_reject + update, numba kernels): "
                                      "sounding 0 x 1500 iterations on one core; %d soundings x 400 iterations, one process per core" % cores}
                    except Exception as e:
                        line["cpu_baseline_reference"] = {"unavailable": repr(e)[:200]}
                else:
                    line["cpu_baseline_reference"] = {"unavailable": "time-domain forward of the reference (gatdaem1d) is absent" if wl.tdem
                                                      else "baseline/_ref is not importable on this machine"}
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
    # the two datapoint kinds are independent launches of two persistent kernels: each on its own stream (--streams 1: one
    # after the other), so that the tail of one kind's chains runs under the other kind's bulk (DESIGN.md "Batches overlap")
    ns = 1 if args.streams == 1 else 2
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(ns)] if ns > 1 else [main, main]
    iters_part = [torch.zeros((), dtype=torch.float64, device=dev) for _ in parts]

    def step(i, count=True):
        out = []
        for j, p in enumerate(parts):
            with torch.cuda.stream(streams[j % len(streams)]):
                # sounding index of the random stream: position within this kind's list (streams stay sharding independent)
                r = ops.rjmcmc_run(p["system"], p["opt"], p["data"], p["alt"], seed=SEED + i, first_index=p["first"] // 2,
                                   precision=args.precision, outputs=outputs, buffers=p["bufs"])
                if count:
                    iters_part[j].add_(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
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
    ev0.record(main)
    if ns > 1:
        for st in streams:
            st.wait_event(ev0)
    for i in range(args.steps):
        res = step(i)
    if ns > 1:
        for st in streams:
            main.wait_stream(st)
    ev1.record(main)
    barrier()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    iters = torch.stack(iters_part).sum().reshape(1)
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

    n_e2e = 2   # host calls per kind, back to back on that kind's thread

    def host_call(px, n=n_e2e):
        p, x = px
        its, nbytes = 0.0, 0
        for i in range(n):
            r = ops.rjmcmc_run(p["system"], p["opt"], x["data"], x["alt"], seed=SEED + i, first_index=p["first"] // 2,
                               precision=args.precision, device=local_rank, outputs=outputs)
            its += float(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
            nbytes = sum(v.nbytes for v in r.values())
        return its, nbytes
    for px in zip(parts, h):   # untimed: the library's host-path buffers are allocated on first use
        host_call(px, 1)
    barrier()
    t0 = time.perf_counter()
    if ns > 1:   # one host thread per kind (the library keeps two sets of device buffers with their own streams; ctypes drops the GIL)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(2) as ex:
            got = list(ex.map(host_call, zip(parts, h)))
    else:
        got = [host_call(px) for px in zip(parts, h)]
    e_iters, d2h = sum(g[0] for g in got), sum(g[1] for g in got)
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
                       "l2": "no flush needed: each step rewrites GBs of posterior arrays per GPU (> 126 MB L2)",
                       "pipelining": ("the two kinds run on two CUDA streams (two persistent kernels, separate result buffers): the tail of one "
                                      "kind's chains overlaps the other kind's bulk" if ns > 1 else "one kind after the other on one stream")},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": float(ei.item()) / float(el.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": n_e2e, "calls_in_flight": 2 if ns > 1 else 1,
                    "api": "geobipy_b200.ops.rjmcmc_run(numpy), one call per datapoint kind" + (", one host thread per kind" if ns > 1 else "")},
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
