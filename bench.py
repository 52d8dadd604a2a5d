#!/usr/bin/env python
"""bench.py -- the measurement contract of this repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg3] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one MCMC step of every chain (one pass of the hot path over the batch).

Headline workload (BASELINE.json `metric` "chain-steps/s ..., HMC logistic regression", configs[3]):
  cfg4  HMCDA logistic regression, synthetic N = 1e6, d = 100 (X = [1, N(0,1)], beta0 ~ N(0,1)/sqrt(d), seed 4),
        10 000 chains per GPU (chain-sharded: no collective; weak scaling), sampling phase of HMCDA: the
        dual-averaged step size found by the pilot adaptation (tools/pilot_adapt.py, profiles/pilot_cfg4.json)
        is restored with mcmcgpu_run_set_state, len = 0.02 => nLeaps = round(len/eps) leapfrogs per step.
Other workloads: cfg2 (HMC(0.75), 3-D Normal, 65 536 chains, fused per-chain kernel), cfg3 (MALA probit
N = 1e5, d = 20, 16 384 chains).

`value`      device-timed (CUDA events on the launching stream, barrier + synchronize on both sides, max over
             ranks), inputs resident in HBM.
`e2e`        the same K steps through the C-ABI run call with HOST buffers: H2D of the chain state, execute,
             D2H of the kept draws / gradients / accept flags / log-targets into pinned host memory.
`roofline`   the likelihood kernel K1 against the FP64 tensor (DMMA) roofline: algorithmic 4*N*d flop per chain
             per evaluation, peak = cuBLAS FP64 GEMM measured live (MEASURED_PEAKS.json has no FP64 entry).
`cpu_baseline` the oracle (CPU restatement of the reference, the reference itself being Julia 0.2 source that
             cannot run here) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dual-averaged HMCDA step size after burn-in on cfg4 (median over 512 pilot chains; profiles/pilot_cfg4.json)
CFG4_EPS = 1.97e-3
CFG4_LEN = 0.02
CFG2_STEPS_PER_STEP = 200     # cfg2: MCMC steps per bench "step" (a single MCMC step of 65 536 3-D chains is ~2 us of work)


def synth_logistic(N, d, seed):
    """X = [1, N(0,1)^(d-1)], beta0 ~ N(0,1)/sqrt(d), y ~ Bernoulli(sigmoid(X beta0)). Column-major X."""
    r = np.random.default_rng(seed)
    Xt = np.empty((d, N))
    Xt[0] = 1.0
    for j in range(1, d):
        Xt[j] = r.standard_normal(N)
    b0 = r.standard_normal(d) / np.sqrt(d)
    eta = b0 @ Xt
    y = (r.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return Xt.T, y, b0          # Xt.T is an F-ordered N x d view


def synth_probit(N, d, seed):
    from scipy.special import ndtr
    r = np.random.default_rng(seed)
    Xt = np.empty((d, N))
    Xt[0] = 1.0
    for j in range(1, d):
        Xt[j] = r.standard_normal(N)
    b0 = 0.5 * r.standard_normal(d)
    y = (r.random(N) < ndtr(b0 @ Xt)).astype(np.float64)
    return Xt.T, y, b0


WORKLOADS = {
    "cfg4": dict(desc="HMCDA logistic regression N=1e6 d=100, 10000 chains/GPU (BASELINE configs[3])", family="logistic",
                 N=1000000, d=100, chains=10000, sampler="HMCDA", seed=4),
    "cfg3": dict(desc="MALA probit regression N=1e5 d=20, 16384 chains/GPU (BASELINE configs[2])", family="probit",
                 N=100000, d=20, chains=16384, sampler="MALA", seed=3),
    "cfg5": dict(desc="HMC logistic regression, tall data d=200, rows sharded over the GPUs (6.25e6 rows per GPU; 5e7 at 8 GPUs), "
                      "256 replicated chains, NCCL all-reduce of (d+2) x C partials per leapfrog (BASELINE configs[4])",
                 family="logistic", N=6250000, d=200, chains=256, sampler="HMC", seed=5),
    "cfg2": dict(desc="HMC(0.75) 3-D Normal -dot(v,v), 65536 chains/GPU (BASELINE configs[1])", family="normal_fn",
                 N=0, d=3, chains=65536, sampler="HMC", seed=1),
}


class ClockSampler:
    """samples SM clock and throttle reasons during the timed region (NVML)"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def dgemm_peak_tflops(torch, n=8192, reps=5):
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e9


def make_problem(wl):
    if wl["family"] == "logistic":
        X, y, b0 = synth_logistic(wl["N"], wl["d"], wl["seed"])
        return X, y, (1.0, -1.0), b0
    if wl["family"] == "probit":
        X, y, b0 = synth_probit(wl["N"], wl["d"], wl["seed"])
        return X, y, (10.0,), b0
    return None, None, (), np.ones(wl["d"])


def sampler_for(wl, capi_or_oracle, is_oracle=False, force_eps=None):
    mk = capi_or_oracle.sampler if is_oracle else capi_or_oracle.sampler_cfg
    if wl["sampler"] == "HMCDA":
        kw = dict(len=CFG4_LEN, max_leaps=64)
        if is_oracle:
            kw["force_eps"] = force_eps
        return mk("HMCDA", **kw)
    if wl["sampler"] == "MALA":
        return mk("MALA", scale=wl.get("drift", 2.4 ** 2 * wl["d"] ** (-1.0 / 3.0) / wl["N"]))
    if wl["family"] == "logistic":          # cfg5: step size scaled from the cfg4 pilot by sqrt(N) and d^(1/4)
        ntot = wl["N"] * int(os.environ.get("WORLD_SIZE", "1"))
        return mk("HMC", scale=0.8 * CFG4_EPS * (1e6 / ntot) ** 0.5 * (100.0 / wl["d"]) ** 0.25, nleaps=10)
    return mk("HMC", scale=0.75, nleaps=10)


# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline(wl, steps, warmup, cores, problem=None):
    """oracle ("port") on `cores` host threads, one independent chain per thread (the prun analogue,
    runners.jl:35-42); ctypes releases the GIL so the threads run in parallel."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    X, y, hy, b0 = problem if problem is not None else make_problem(wl)
    d = wl["d"]
    om = O.Model(wl["family"], d, X, y, hy)
    last = warmup + steps
    results = [None] * cores

    def chain_steps(k, nsteps, state):
        rng = np.random.default_rng(1000 + k)
        zn = rng.standard_normal((nsteps + 1, d)); un = rng.random(nsteps + 1)
        fe = np.full(nsteps + 1, CFG4_EPS) if wl["sampler"] == "HMCDA" else None
        res = O.run_chain(om, sampler_for(wl, O, True, fe), (1, 1, nsteps), state, None, zn, un)
        return res["samples"][-1], res["n_grad_evals"]

    init = b0 if wl["family"] != "normal_fn" else np.ones(d)
    states = [init.copy() for _ in range(cores)]

    def work(k, nsteps, out):
        out[k] = chain_steps(k, nsteps, states[k])

    def run_all(nsteps):
        out = [None] * cores
        ts = [threading.Thread(target=work, args=(k, nsteps, out)) for k in range(cores)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.perf_counter() - t0
        for k in range(cores):
            states[k] = out[k][0]
        return dt, sum(o[1] for o in out)

    if warmup > 0:
        run_all(warmup)
    dt, nev = run_all(steps)
    return dict(value=cores * steps / dt, unit="chain-steps/s", cores=cores, kind="port",
                sample=f"{cores} chains x {steps} steps (1 chain per host thread) of the {wl['chains']}-chain workload, full N",
                grad_evals_per_s=nev / dt, seconds=dt)


def run_reference_arm(args, wl, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: steps sized so the run ends within a few minutes on the host cores
    steps, warmup = args.steps, args.warmup
    if wl_name == "cfg2":
        steps, warmup = max(steps, 1) * 2000, max(warmup, 1) * 200
    cb = cpu_baseline(wl, steps, warmup, cores)
    line = dict(metric="chain-steps/s", value=cb["value"], unit="chain-steps/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * cb["seconds"] / steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=wl_name, description=wl["desc"], N=wl["N"], d=wl["d"], sampler=wl["sampler"],
                            note="reference = CPU oracle (C restatement of MCMC.jl; the Julia 0.2 reference cannot run here), "
                                 "all host threads, one chain per thread"),
                cpu_baseline=dict(value=cb["value"], unit="chain-steps/s", cores=cores, kind="port", sample=cb["sample"]),
                e2e=dict(value=cb["value"], unit="chain-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                grad_evals_per_s=cb["grad_evals_per_s"], gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))   # the WORKLOADS dict is defined above
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--N", type=int, default=0, help="override observation count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ess", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.chains:
        wl["chains"] = args.chains
    if args.N:
        wl["N"] = args.N
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3                      # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference_arm(args, wl, args.workload)

    # keep stdout clean for the one JSON line: libraries (NCCL prints its version banner to stdout) write to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import mcmc_jl_b200  # noqa: F401
    from mcmc_jl_b200 import _capi as capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)       # library launches on torch's stream so torch events bracket them
    ctx.set_option("time_eval", 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W, C, d = args.steps, args.warmup, wl["chains"], wl["d"]
    row_sharded = (args.workload == "cfg5")
    if row_sharded:
        # tall data: every rank generates ITS rows on the device (seeded per rank), chains are replicated
        if world > 1:
            uid = mcmc_jl_b200.broadcast_unique_id(dist, capi.Context.comm_unique_id, rank)
            ctx.comm_init(rank, world, uid)
        g = torch.Generator(device="cuda"); g.manual_seed(wl["seed"] * 1000 + rank)
        g0 = torch.Generator(device="cuda"); g0.manual_seed(wl["seed"])
        b0_t = torch.randn(d, generator=g0, device="cuda", dtype=torch.float64) / d ** 0.5
        Xt = torch.randn(d, wl["N"], generator=g, device="cuda", dtype=torch.float64)   # (d, N) row-major == N x d column-major
        Xt[0] = 1.0
        y_t = (torch.rand(wl["N"], generator=g, device="cuda", dtype=torch.float64) < torch.sigmoid(b0_t @ Xt)).double()
        torch.cuda.synchronize()
        dm = capi.DeviceModel.from_device(ctx, "logistic", wl["N"], d, Xt.data_ptr(), y_t.data_ptr(), (1.0, -1.0), row_sharded=world > 1)
        b0, hy, problem = b0_t.cpu().numpy(), (1.0, -1.0), None
        del Xt, y_t
        torch.cuda.empty_cache()
        offset = 0                            # replicated chains: identical Philox keys on every rank
    else:
        problem = make_problem(wl)
        X, y, hy, b0 = problem
        dm = capi.DeviceModel(ctx, wl["family"], d, X, y, hy)
        offset = rank * C                     # global chain ids: Philox streams differ across ranks
    scfg = sampler_for(wl, capi)
    rng = np.random.default_rng(100 + rank)
    peak = dgemm_peak_tflops(torch) if wl["family"] != "normal_fn" else None

    pin = lambda shape, dt=np.float64: torch.empty(shape, dtype={np.float64: torch.float64, np.uint8: torch.uint8}[dt]).pin_memory().numpy()

    if args.workload == "cfg2":
        # fused engine: the whole chain is one launch, so warm-up and timed region are separate runs of W and K steps
        def fresh(nsteps):
            return capi.DeviceRun(dm, scfg, (1, 1, nsteps), C, np.ones(d), seed=wl["seed"], chain_offset=offset, engine="fused")
        K, W = K * CFG2_STEPS_PER_STEP, W * CFG2_STEPS_PER_STEP     # one launch covers the whole chain: time K*200 MCMC steps
        r = fresh(W); r.execute(); r.close()
        r = fresh(K)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record(stream); info = r.execute(); e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
        launches = info["n_launches"]
        eval_ms, n_eval = ms, 1
        r.close()
        init_state = np.ones(d)
        set_state = None
    else:
        # wave engine: one run, W warm-up steps then K timed steps of the same chains
        if wl["sampler"] == "HMCDA":
            init_state = b0[None, :] + 2e-3 * rng.standard_normal((C, d))
            step0 = 1000                      # sampling phase: past any burn-in, step size frozen at the adapted value
            set_state = (step0, np.full(C, CFG4_EPS), np.full(C, CFG4_EPS), np.zeros(C))
        else:
            init_state = b0[None, :] + 1e-3 * np.random.default_rng(100).standard_normal((C, d))   # same on every rank
            step0, set_state = 0, None
        r = capi.DeviceRun(dm, scfg, (step0 + 1, 1, step0 + W + K), C, init_state, seed=wl["seed"], chain_offset=offset, engine="wave")
        if set_state:
            r.set_state(*set_state)
        r.execute_steps(W)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record(stream); info = r.execute_steps(K); e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
        launches = info["n_launches"]
        eval_ms, n_eval = info["eval_ms"], info["n_waves"]
        acc_rate = float(r.fetch(samples=False, grads=False, logtarget=False)["accept"][:, W:].mean())
        r.close()

    # ---- e2e: same K steps through the C-ABI call with host buffers (pinned), H2D + execute + D2H in the timed region
    S = K
    bufs = dict(samples=pin((C, S, d)), grads=pin((C, S, d)), accept=pin((C, S), np.uint8), logtarget=pin((C, S)))
    barrier()
    t0 = time.perf_counter()
    if args.workload == "cfg2":
        r = capi.DeviceRun(dm, scfg, (1, 1, K), C, init_state, seed=wl["seed"], chain_offset=offset, engine="fused")
        r.execute()
        h2d = d * 8
    else:
        r = capi.DeviceRun(dm, scfg, (step0 + 1, 1, step0 + K), C, init_state, seed=wl["seed"] + 1, chain_offset=offset, engine="wave")
        if set_state:
            r.set_state(*set_state)
        r.execute()
        h2d = C * d * 8 + (3 * C * 8 if set_state else 0)
    r.fetch(out=bufs)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    r.close()
    d2h = C * S * d * 8 * 2 + C * S + C * S * 8

    # ---- reduce over ranks: max time
    times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    evals = torch.tensor([float(info["n_grad_evals"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        if not row_sharded:                   # replicated chains: every rank counts the same evaluations
            dist.all_reduce(evals, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = times.tolist()
    total_chains = C if row_sharded else C * world
    value = total_chains * K / (ms_max / 1e3)
    e2e_value = total_chains * K / (e2e_ms_max / 1e3)

    if rank == 0:
        line = dict(metric="chain-steps/s", value=value, unit="chain-steps/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=args.workload, description=wl["desc"], N=wl["N"], d=d, chains_per_gpu=C,
                                sampler=wl["sampler"],
                                parallelism=(f"rows sharded over {world} GPU(s) ({wl['N']} rows each), chains replicated, ncclAllReduce per leapfrog"
                                             if row_sharded else f"chains sharded over {world} GPU(s), no collective"),
                                l2="inputs larger than L2 (packed X = %.0f MB)" % (wl["N"] * (8 * ((d + 7) // 8) + 4) * 8 / 1e6)
                                if wl["N"] else "no input data; kept draws written once",
                                seed=wl["seed"]),
                    grad_evals_per_s=evals.item() / (ms_max / 1e3), gpu_launches=int(launches),
                    e2e=dict(value=e2e_value, unit="chain-steps/s", h2d_bytes_per_step=h2d / K, d2h_bytes_per_step=d2h / K),
                    clocks=clk.summary())
        if args.workload != "cfg2":
            per_launch_ms = eval_ms / max(n_eval, 1)
            flop = 4.0 * wl["N"] * d * C             # per GPU (its rows x all chains for cfg5; all rows x its chains otherwise)
            ach = flop / per_launch_ms / 1e9
            traffic = None
            prof = os.path.join(ROOT, "profiles", "k1_full_r01_summary.json")
            if args.workload == "cfg4" and wl["N"] == 1000000 and C == 10000 and os.path.exists(prof):
                m = json.load(open(prof))["metrics"]     # one ncu --set full capture of this kernel on this workload
                traffic = float(m["dram__bytes_read.sum"]["values"][0]) * 1e9 + float(m["dram__bytes_write.sum"]["values"][0]) * 1e6
            line["roofline"] = dict(bound="tensor", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, traffic=traffic,
                                    kernel="k1_kernel (FP64 DMMA m8n8k4)", ms_per_launch=per_launch_ms,
                                    share_of_step=eval_ms / ms,
                                    peak_source="cuBLAS FP64 GEMM 8192^3 measured live in this run (MEASURED_PEAKS.json has no FP64 entry)")
            line["config"].update(eps=CFG4_EPS if wl["sampler"] == "HMCDA" else None, len=CFG4_LEN if wl["sampler"] == "HMCDA" else None,
                                  accept_rate=acc_rate, grad_evals_per_chain_step=info["n_grad_evals"] / (C * K))
        else:
            bytes_per = 8 * d * 2 + 8 + 1       # sample + gradient + log-target + accept flag per kept chain-step
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
            ach = bytes_per * C * K / (ms / 1e3) / 1e9
            line["roofline"] = dict(bound="hbm", achieved=ach, peak=hbm, unit="GB/s", frac=ach / hbm, traffic=None,
                                    kernel="fused_chain_kernel", note="store bandwidth of kept draws; the kernel is FP64-ALU/latency bound, see DESIGN.md")
        if not args.no_ess and not row_sharded:
            # min-ESS/s (BASELINE metric): ESS per chain-step of this sampler on this target, measured with the
            # device stats pass (Geyer IMSE, ess.jl:6-10) on a bounded pilot of the same chains, times the measured chain-steps/s
            if args.workload == "cfg2":
                ce, se = C, 2000
                rr = capi.DeviceRun(dm, scfg, (201, 1, se), ce, np.ones(d), seed=wl["seed"], chain_offset=offset, engine="fused",
                                    store_grad=False, store_logtarget=False)
                rr.execute()
            else:
                ce, se, skip = 256, 300, 50
                # `skip` steps from the synthetic start are run first and discarded through a state hand-over, so that the kept
                # range starts right after step0' = step0 + skip (a kept range starting later would turn HMCDA's burn-in
                # adaptation back on: it runs while i < first - 1, HMCDA.jl:133)
                r0 = capi.DeviceRun(dm, scfg, (step0 + 1, 1, step0 + skip), ce, init_state[:ce], seed=wl["seed"] + 7, engine="wave",
                                    store_grad=False, store_logtarget=False)
                if set_state:
                    r0.set_state(set_state[0], *(a[:ce] for a in set_state[1:]))
                r0.execute()
                st0 = r0.get_state()
                r0.close()
                rr = capi.DeviceRun(dm, scfg, (step0 + skip + 1, 1, step0 + se), ce, st0["pars"], seed=wl["seed"] + 7, engine="wave",
                                    store_grad=False, store_logtarget=False)
                if set_state:
                    rr.set_state(step0 + skip, st0["leapstep"], st0["dual_leapstep"], st0["dualH"])
                else:
                    rr.set_state(step0 + skip)
                rr.execute()
            st = rr.stats("imse")
            kept = rr.S
            # ess.jl:9 is n*var_iid/var_imse; for antithetic chains (HMC overshooting: lag-1 autocorrelation < -0.5) Geyer's
            # estimate is <= 0 or tiny and the ratio is negative or above n: such parameters are counted as ESS = n
            # (at least as good as independent draws), i.e. ESS is clipped to (0, n]
            ess = np.where((st["ess"] > 0) & (st["ess"] < kept), st["ess"], float(kept))
            ess_per_step = float(np.median(ess.min(axis=1)) / kept)
            rr.close()
            line["min_ess_per_s"] = dict(value=ess_per_step * value, unit="min-ESS/s (min over parameters, summed over chains)",
                                         ess_per_chain_step=ess_per_step, sample=f"{ce} chains x {kept} kept steps, Geyer IMSE on the device, ESS clipped to (0, n]")
        if not args.no_cpu_baseline and not row_sharded:
            cores = os.cpu_count() or 1
            cs = args.cpu_steps if args.workload != "cfg2" else 2000
            cb = cpu_baseline(wl, cs, 1 if args.workload != "cfg2" else 200, cores, problem)
            line["cpu_baseline"] = dict(value=cb["value"], unit="chain-steps/s", cores=cores, kind="port", sample=cb["sample"],
                                        grad_evals_per_s=cb["grad_evals_per_s"])
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
