#!/usr/bin/env python
"""bench.py -- PSVI hot-path benchmark on B200 (contract: see the task description; metric from BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference (baseline/_ref) on the host CPU

Workload (BASELINE.json configs[1], "cfg2" in SURVEY.md section 8): halfmoon-shaped data (N=800, D=2, C=2),
`fn` BNN with one hidden layer of 100 units, psvi_learn_v, coreset M=50, S=10 MC samples, T=100 unrolled inner Adam
steps, minibatch B=128.  One STEP = one PSVI outer step = PSVILearnV.nested_step: T inner steps + outer psvi_elbo +
hypergradient on (u, v) + the Adam updates of u and v.

  value : outer steps/s with the minibatch already resident in HBM (device-timed, CUDA events per step, L2 flushed
          between steps); for N > 1: N * K / (max over ranks of the summed step times) -- one all-reduce(MAX)
  e2e   : outer steps/s through the public API with HOST (pinned) minibatches: H2D copy of x/y and D2H copy of every
          step's loss inside the timed region
  N > 1 : one independent PSVI chain (trial) per GPU -- the reference's own multi-GPU mode
          (flow-psvi-parallel.py:455-479): cfg2's data term is 128 rows, so this path is "replicas only" (DESIGN.md 6).
          The part of the hot path that DOES shard -- the full-data pass at BASELINE configs[4] ("cfg5": 10 M rows x 256,
          fn H=1024, S=64) -- is timed as a STRONG-scaling run (fixed 10 M rows split over the ranks, one all-reduce)
          and reported in the first-class field `sharded_fulldata`.
  --impl reference : the reference's own PSVILearnV.nested_step (pip-installed into baseline/_ref, unmodified) on the host
          CPU with all cores; also, when a GPU is visible, the same unmodified class eagerly on cuda:0
          (`extra.eager_b200`, SURVEY.md section 8d "the real on-box comparator").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200"))

CFG = dict(workload="cfg2: halfmoon N=800 D=2 C=2, fn H=100 (1 hidden layer), psvi_learn_v, M=50, S=10, T=100, B=128",
           D=2, H=100, C=2, M=50, S=10, T=100, B=128, N=800, init_sd=1e-3)
METRIC, UNIT = "psvi_outer_steps_per_sec", "outer_steps/s"


def make_config(world):
    """The `config` object -- identical in both arms (`--impl ours` / `--impl reference`) for the same --gpus."""
    return dict(CFG, parallelism="single chain" if world == 1 else
                f"{world} independent chains (trials), one per GPU, no data-path collective",
                l2="GPU arm: flushed between timed steps (256 MiB write); CPU arm: n/a",
                noise="GPU arm: in-kernel Philox4x32-10; reference arm: torch generator")


def flops_outer_step(c):
    """Algorithmic FLOPs of one outer step: F_fwd(R) = 2 S R (D H + H C);
    inner step = 3 F_fwd(M); reverse step = 6 F_fwd(M); outer fwd+bwd = 3 F_fwd(M+B)."""
    f = lambda R: 2.0 * c["S"] * R * (c["D"] * c["H"] + c["H"] * c["C"])  # noqa: E731
    return c["T"] * (3 + 6) * f(c["M"]) + 3 * f(c["M"] + c["B"])


def make_data(seed=42):
    import numpy as np
    import torch
    from sklearn.datasets import make_moons
    X, Y = make_moons(n_samples=1000, noise=0.1, random_state=seed)
    X, Y = torch.from_numpy(X.astype(np.float32)), torch.from_numpy(Y.astype(np.float32))
    return X[:800], Y[:800], X[800:], Y[800:]


# ------------------------------------------------------------------------------------------------ clocks sampler
class Clocks:
    """SM clock and throttle reasons of one GPU, sampled WHILE the timed regions run.  In-process NVML (pynvml, one query
    takes microseconds) on a thread every 10 ms; `nvidia-smi -lms` as the fallback.  start() is called before the warm-up so
    that NVML / nvidia-smi initialisation (seconds on an 8-GPU box) is over when mark() opens the timed window; stop() keeps
    the samples taken inside the window."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t0, self.nvml, self.stop_flag, self.mx = [], None, index, 0.0, None, False, None

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.h = self._handle()
            self.mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def mark(self):
        self.t0 = time.monotonic()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.monotonic(), mhz, self.mx, {name for bit, name in self.REASONS if bits & bit}))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        names = [name for _, name in self.REASONS]
        for line in self.proc.stdout:
            r = [t.strip() for t in line.split(",")]
            try:
                self.rows.append((time.monotonic(), float(r[0]), float(r[1]),
                                  {n for n, v in zip(names, r[3:7]) if v.lower().startswith("active")}))
            except Exception:
                continue

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi on this box"], "samples": 0}
        t_end = time.monotonic()
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        rows = [r for r in self.rows if self.t0 <= r[0] <= t_end]
        note = []
        if not rows and self.rows:      # timed window shorter than the sampling period: the nearest samples
            rows = sorted(self.rows, key=lambda r: min(abs(r[0] - self.t0), abs(r[0] - t_end)))[:3]
            note = ["no sample fell inside the timed window; nearest samples reported"]
        sm = sorted(r[1] for r in rows)
        reasons = set().union(*[r[3] for r in rows]) if rows else set()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": rows[-1][2] if rows else self.mx,
               "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}
        if note:
            out["note"] = note[0]
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_fn(c):
    """One outer step of the oracle port (numpy restatement of the reference path, fp32 like the reference)."""
    import numpy as np
    from oracle import psvi_oracle as po
    rng = np.random.default_rng(0)
    dims = [c["D"], c["H"], c["C"]]
    P = po.p_theta(dims)
    x, y, _, _ = make_data()
    x, y = x.numpy(), y.numpy()
    f32 = np.float32
    mu = (rng.standard_normal(P) * 0.3).astype(f32)
    rho = np.full(P, po.inverse_softplus(c["init_sd"]), f32)
    u = x[: c["M"]].astype(f32).copy()
    z = y[: c["M"]].astype(f32)
    v = np.zeros(c["M"], f32)
    st = dict(mu=mu, rho=rho, u=u, v=v)

    def step():
        eps = rng.standard_normal((c["T"] + 1, c["S"], P)).astype(f32)
        idx = rng.permutation(c["N"])[: c["B"]]
        r = po.nested_step(st["mu"], st["rho"], eps[: c["T"]], eps[c["T"]], st["u"], z, st["v"], x[idx].astype(f32),
                           y[idx], f32(c["N"]), dims, f32(1e-3), vmode=1)
        st["mu"], st["rho"] = r["mu_T"].astype(f32), r["rho_T"].astype(f32)
        st["u"] = (st["u"] - 1e-4 * np.sign(r["u_grad"])).astype(f32)   # first-step Adam == sign step
        st["v"] = (st["v"] - 1e-3 * np.sign(r["v_grad"])).astype(f32)
        return float(r["loss"])
    return step


def time_cpu(c, steps, warmup):
    step = cpu_step_fn(c)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps


def reference_sample(steps, warmup, device="cpu", anomaly=False, budget_s=None):
    """steps/s of the UNMODIFIED reference's outer step (oracle/ref_runner.py drives baseline/_ref); None if the reference
    tree is not reachable."""
    try:
        from oracle.ref_runner import time_reference
        return time_reference(CFG, steps, warmup, device=device, anomaly=anomaly, threads=os.cpu_count(), budget_s=budget_s)
    except ImportError as e:
        return {"error": repr(e)[:200]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    r = reference_sample(K, W, "cpu", False, budget_s=240.0)
    extra = {}
    if "error" in r:
        # reference tree not on this box: the oracle port (numpy restatement) stands in, and says so
        sps, sec = time_cpu(CFG, max(1, min(K, 40)), max(1, W))
        kind, cores, steps_done = "port", 1, max(1, min(K, 40))
        sample = (f"{steps_done} full cfg2 outer steps of oracle/psvi_oracle.py (numpy fp32, single thread): the reference "
                  f"tree was not importable here ({r['error']})")
    else:
        sps, sec, kind, cores, steps_done = r["steps_per_s"], r["s_per_step"], "reference", r["threads"], r["steps"]
        sample = (f"{steps_done} full cfg2 outer steps of the unmodified reference PSVILearnV.nested_step ({r['root']}, "
                  f"torch CPU, {r['threads']} threads on {r['cores']} cores, set_detect_anomaly off; next(iter(train_loader)) "
                  "+ nested_step + loss.item() per step)")
        a = reference_sample(2, 1, "cpu", True, budget_s=60.0)
        extra["anomaly_on"] = ({"steps_per_s": a["steps_per_s"], "steps": a["steps"],
                                "what": "same, with torch.autograd.set_detect_anomaly(True) as flow_psvi.py:50 ships"}
                               if "error" not in a else a)
        try:
            import torch
            if torch.cuda.is_available():
                g = reference_sample(min(K, 30), max(1, min(W, 3)), "cuda", False, budget_s=90.0)
                extra["eager_b200"] = ({"steps_per_s": g["steps_per_s"], "ms_per_step": g["s_per_step"] * 1e3, "steps": g["steps"],
                                        "what": "the same unmodified reference class running eagerly on cuda:0 (PyTorch ATen "
                                                "kernels + autograd double backward); SURVEY.md section 8d on-box comparator"}
                                       if "error" not in g else g)
        except Exception as e:  # noqa: BLE001
            extra["eager_b200"] = {"error": repr(e)[:200]}
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
            "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": make_config(args.gpus),
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": extra}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def build_chain(c, seed, device):
    import torch
    from psvi.experiments.experiments_utils import SynthDataset
    from psvi.inference.psvi_classes import PSVILearnV
    x, y, xt, yt = make_data()
    tr, te = SynthDataset(x, y), SynthDataset(xt, yt)
    kw = dict(mc_samples=c["S"], num_epochs=0, data_minibatch=c["B"], D=c["D"], N=c["N"], inner_it=c["T"],
              trainer="nested", log_every=150, lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample",
              init_sd=c["init_sd"], num_pseudo=c["M"], seed=seed, architecture="fn", n_hidden=c["H"], n_layers=1,
              logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=c["C"],
              compute_weights_entropy=False, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    return obj, (x, y, xt, yt)


def bench_replicas(c, dev, n_chains, K, W):
    """The reference's own multi-trial mode (flow-psvi-parallel.py:455-479: independent trials in parallel) on ONE GPU: n_chains
    independent cfg2 chains, one CUDA stream each.  A chain's step is one 10-CTA cluster launch (10 of the SMs), so several
    chains run concurrently; the aggregate is reported as an extra, never as the headline."""
    import torch
    import psvi.inference.psvi_classes as pc
    real = pc._dist_info
    pc._dist_info = lambda: (None, 0, 1)
    try:
        chains = [build_chain(c, seed=100 + i, device=dev) for i in range(n_chains)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_chains)]
        g = torch.Generator().manual_seed(99)
        batches = []
        for obj, (x, y, _, _) in chains:
            xd, yd = x.to(dev), y.to(dev)
            idx = [torch.randperm(c["N"], generator=g)[: c["B"]].to(dev) for _ in range(K + W)]
            batches.append([(xd[i].contiguous(), yd[i].contiguous()) for i in idx])
        torch.cuda.synchronize()

        def run(lo, hi):
            for k in range(lo, hi):
                for (obj, _), st, bt in zip(chains, streams, batches):
                    with torch.cuda.stream(st):
                        obj.nested_step(*bt[k])
        run(0, W)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in streams]
        for (a, _), st in zip(evs, streams):
            a.record(st)
        t0 = time.perf_counter()
        run(W, W + K)
        for (_, b), st in zip(evs, streams):
            b.record(st)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = max(a.elapsed_time(b) for a, b in evs)
        span = max(dev_ms * 1e-3, wall)
        return {"what": f"{n_chains} independent cfg2 chains (trials) on one GPU, one stream each, {K} outer steps per chain; "
                        "aggregate = chains * steps / max(device span, wall clock)",
                "chains": n_chains, "aggregate_outer_steps_per_s": n_chains * K / span, "device_span_ms": dev_ms,
                "wall_ms": wall * 1e3}
    except Exception as e:  # noqa: BLE001 -- an extra must not take the headline down
        return {"error": repr(e)[:300]}
    finally:
        pc._dist_info = real


_STDOUT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _STDOUT_FD is None:
        os.write(1, data)
    else:
        os.write(_STDOUT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner, library
    # chatter) is sent to stderr for the whole run; the line goes to the saved descriptor at the end
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from psvi import _native
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _native.require_cuda()
    c = CFG
    K, W = args.steps, max(args.warmup, 3)

    # -------- chain (one per rank; independent trials when world > 1) ----------------------------------------------
    import torch.distributed as _d
    # independent replicas: the sharded code path of nested_step must stay off for the headline number
    obj, (x, y, xt, yt) = build_chain(c, seed=rank, device=dev)
    import psvi.inference.psvi_classes as pc
    real_dist_info = pc._dist_info
    pc._dist_info = lambda: (None, 0, 1)
    xd, yd = x.to(dev), y.to(dev)
    g = torch.Generator().manual_seed(1234 + rank)
    idxs = [torch.randperm(c["N"], generator=g)[: c["B"]] for _ in range(K + W)]
    dev_batches = [(xd[i.to(dev)].contiguous(), yd[i.to(dev)].contiguous()) for i in idxs]
    host_batches = [(x[i].contiguous().pin_memory(), y[i].contiguous().pin_memory()) for i in idxs]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -------- value: device-resident inputs, per-step CUDA events, L2 flush between steps ----------------------------
    launches0 = _native.launch_count() if hasattr(_native, "launch_count") else 0
    clocks = Clocks(local)
    clocks.start()
    for i in range(W):
        obj.nested_step(*dev_batches[i])
    barrier()
    clocks.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches1 = _native.launch_count()
    for i in range(K):
        flush.zero_()
        ev[i][0].record(stream)
        obj.nested_step(*dev_batches[W + i])
        ev[i][1].record(stream)
    barrier()
    launches = _native.launch_count() - launches1
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(ms)

    # -------- kernel-only time of the dominant kernel (same launches, events right around the native call) ----------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(K, 50))]
    for i, (a, b) in enumerate(kev):
        flush.zero_()
        _native.EVENT_HOOK = (a, b, stream)
        obj.nested_step(*dev_batches[W + i])
    _native.EVENT_HOOK = None
    torch.cuda.synchronize()
    kms = sorted(a.elapsed_time(b) for a, b in kev)
    kernel_ms = sum(kms) / len(kms)

    # -------- e2e: host (pinned) minibatches, H2D + D2H of the loss inside the timed region -------------------------
    for i in range(W):
        xb, yb = host_batches[i]
        obj.nested_step(xb.to(dev, non_blocking=True), yb.to(dev, non_blocking=True)).item()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss_host = torch.empty(K, dtype=torch.float32).pin_memory()
    e0.record(stream)
    for i in range(K):
        xb, yb = host_batches[W + i]
        loss = obj.nested_step(xb.to(dev, non_blocking=True), yb.to(dev, non_blocking=True))
        loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)   # every step's loss goes back to the host
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    assert bool(torch.isfinite(loss_host).all()), "non-finite loss in the e2e run"
    clk = clocks.stop()

    # -------- secondary metric: MC log-lik evals/s (fn, M=50) -- both readings of SURVEY.md section 8d --------------
    model, desc, S = obj._model_desc()
    mu, rho = model.flat()
    u, v = obj._uv()
    am, av = torch.zeros(2 * mu.numel(), device=dev), torch.zeros(2 * mu.numel(), device=dev)
    mu2, rho2 = mu.clone(), rho.clone()
    n_ev = 200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        a.record(stream)
        _native.unroll(desc, obj._noise(n_ev), mu2, rho2, am, av, 0, u, obj._z32(), None, v, float(c["N"]), 1, 0.0, n_ev,
                       1e-3, _native.ADAM_ROBUST_HIGHER, None)
        b.record(stream)
        torch.cuda.synchronize()
    inner_evals = n_ev / (a.elapsed_time(b) * 1e-3)
    for rep in range(3):
        a.record(stream)
        for _ in range(20):
            obj.evaluate()
        b.record(stream)
        torch.cuda.synchronize()
    pred_evals = 20 / (a.elapsed_time(b) * 1e-3)

    # -------- sharded full-data predictive pass (the part of the path that shards; one all-reduce) -------------------
    pc._dist_info = real_dist_info
    sharded = None
    try:
        from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
        n_big = 2_000_000
        xt_big, yt_big = make_synthetic_rows(n_big, c["D"], c["C"], seed=0)
        obj.test_dataset = SynthDataset(xt_big, yt_big.float())
        obj.data_minibatch = 8192
        obj._dev_data.pop("test", None)
        obj.evaluate()
        barrier()
        for _ in range(3):
            obj.evaluate()
        barrier()
        reps = 20
        per = []
        for _ in range(reps):
            a.record(stream)
            obj.evaluate()
            b.record(stream)
            torch.cuda.synchronize()
            per.append(a.elapsed_time(b))
        barrier()
        t = torch.tensor([sorted(per)[len(per) // 2]], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = t.item() * 1e-3
        sharded = {"what": f"PSVI.evaluate over {n_big} synthetic rows (D=2, fn H=100, S=10, M=50, batch 8192), rows "
                           f"sharded over {world} rank(s), one all-reduce of 8 floats; median of {reps} passes, max over ranks; "
                           "fp32 CUDA-core predictive kernels (register-form rows kernel of the cfg2 family, DESIGN.md 4.2)",
                   "passes_per_s": 1.0 / sec, "row_samples_per_s": n_big * c["S"] / sec,
                   "ms_min_med_max_rank0": [min(per), sorted(per)[len(per) // 2], max(per)]}
    except Exception as e:  # never lose the headline number to the secondary section
        sharded = {"error": repr(e)[:200]}

    # -------- HBM-bound member of the path: tensor-core full-data predictive pass, logistic regression D=256 ---------
    # (the ">= 80 % of HBM" target is demonstrated on this pass); rows sharded over ranks (weak:
    # 8 M rows per rank), ONE all-reduce of the 8-float result.
    fulldata = None
    try:
        Dl, Cl, Ml, rows = 256, 10, 50, 8_000_000
        gg = torch.Generator(device=dev).manual_seed(7 + rank)
        xb16 = torch.randn(rows, Dl, device=dev, generator=gg, dtype=torch.bfloat16)
        yl = torch.randint(0, Cl, (rows,), device=dev, dtype=torch.int32, generator=gg)
        lu = torch.randn(Ml, Dl, device=dev, generator=gg)
        lz = torch.randint(0, Cl, (Ml,), device=dev, dtype=torch.int32, generator=gg)
        lv = torch.zeros(Ml, device=dev)
        lout = torch.zeros(8, device=dev)
        lnoise = _native.make_noise(None, seed=11, domain=1)
        try:
            HBM_PEAK_GBPS = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            HBM_PEAK_GBPS = 6556.2        # the value the driver measured on this pool (fallback when the file did not travel)

        def lr_case(Sl):
            lmodel = _native.make_model([Dl, Cl], Sl)
            Pl = _native.num_theta(lmodel)
            lmu = 0.1 * torch.randn(Pl, device=dev, generator=gg)
            lrho = torch.full((Pl,), -2.97, device=dev)
            lscr = torch.zeros(_native.lr_predictive_tc_scratch_floats(lmodel), device=dev)

            def lr_pass():
                _native.lr_predictive_tc(lmodel, lnoise, lmu, lrho, lu, lz, lv, xb16, yl, 0, 1.0e4, 1, 0.0, 0, lout, lscr)
                if world > 1:
                    dist.all_reduce(lout)
            for _ in range(3):
                lr_pass()
            barrier()
            reps = 10
            a.record(stream)
            for _ in range(reps):
                lr_pass()
            b.record(stream)
            barrier()
            tt = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = tt.item() * 1e-3 / reps
            byts = world * rows * (Dl * 2 + 4)
            return {"S": Sl, "ms_per_pass": sec * 1e3, "rows_per_s": world * rows / sec, "row_samples_per_s": world * rows * Sl / sec,
                    "algorithmic_bytes": byts, "GBps": byts / sec / 1e9,
                    "frac_of_measured_hbm_peak_per_gpu": byts / sec / 1e9 / world / HBM_PEAK_GBPS}
        fulldata = lr_case(10)
        fulldata["what"] = (f"psvi_lr_predictive_tc: logistic regression D=256 C=10 S=10 M=50, {rows} bf16 rows per rank x "
                            f"{world} rank(s), whole call (log-weights + weight prep + TMA/tcgen05 kernel + reduce)"
                            + (" + all-reduce" if world > 1 else ""))
        fulldata["S16"] = lr_case(16)
    except Exception as e:
        fulldata = {"error": repr(e)[:300]}

    # -------- tensor-bound member of the path (cfg5 shapes): sampled-GEMM forward of fn D=256 H=1024 C=10 with S=64 MC
    # samples over the full data (psvi_fn_predictive_tc: importance weights from M=1000 pseudo-points + mixture over rows);
    # rows sharded over ranks (weak: 524 288 rows per rank), ONE all-reduce of the 8-float result.
    fn_tc = None
    fx = fy = fscr = None
    try:
        Df, Hf, Cf, Sf, Mf, rows_f = 256, 1024, 10, 64, 1000, 524_288
        fmodel = _native.make_model([Df, Hf, Cf], Sf)
        Pf = _native.num_theta(fmodel)
        gg = torch.Generator(device=dev).manual_seed(17 + rank)
        fmu = torch.cat([torch.randn(Hf * Df, device=dev, generator=gg) / Df ** 0.5, torch.zeros(Hf, device=dev),
                         torch.randn(Cf * Hf, device=dev, generator=gg) / Hf ** 0.5, torch.zeros(Cf, device=dev)]).contiguous()
        frho = torch.full((Pf,), -6.9, device=dev)
        fu = torch.randn(Mf, Df, device=dev, generator=gg)
        fz = torch.randint(0, Cf, (Mf,), device=dev, dtype=torch.int32, generator=gg)
        fv = torch.zeros(Mf, device=dev)
        fx = torch.randn(rows_f, Df, device=dev, generator=gg, dtype=torch.bfloat16)
        fy = torch.randint(0, Cf, (rows_f,), device=dev, dtype=torch.int32, generator=gg)
        fout = torch.zeros(8, device=dev)
        fscr = torch.zeros(_native.fn_tc_scratch_floats(fmodel, rows_f, Mf), device=dev)
        fnoise = _native.make_noise(None, seed=13, domain=1)

        def fn_pass():
            _native.fn_predictive_tc(fmodel, fnoise, fmu, frho, fu, fz, fv, fx, fy, 0, 1.0e4, 1, 0.0, 0, fout, fscr)
            if world > 1:
                dist.all_reduce(fout)
        for _ in range(3):
            fn_pass()
        barrier()
        reps = 8
        a.record(stream)
        for _ in range(reps):
            fn_pass()
        b.record(stream)
        barrier()
        tt = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = tt.item() * 1e-3 / reps
        fl_f = 2.0 * Sf * (rows_f + Mf) * (Df * Hf + Hf * Cf)       # F_fwd(N) of SURVEY.md section 8d, per rank
        fn_tc = {"what": f"psvi_fn_predictive_tc: fn D=256 H=1024 C=10, S=64, M=1000, {rows_f} bf16 rows per rank x {world} "
                         "rank(s); whole call (Philox sampling of 64 x 273k weights, pseudo-data forward for the importance "
                         "weights, TMA/tcgen05 forward over the rows, reduce)" + (" + all-reduce" if world > 1 else ""),
                 "ms_per_pass": sec * 1e3, "rows_per_s": world * rows_f / sec, "row_samples_per_s": world * rows_f * Sf / sec,
                 "flops_per_pass_per_rank": fl_f, "TFLOPs_per_gpu": fl_f / sec / 1e12}
    except Exception as e:
        fn_tc = {"error": repr(e)[:300]}

    # -------- cfg4 (lenet, M=200, S=10, B=128): one PSVI outer step through the convolutional kernels, T=20 as in the
    # reference's own lenet runs ------------------------------------------------------------------------
    lenet = None
    try:
        from tests.fake_mnist import FakeMNIST
        ltr, lte = FakeMNIST(1024, 0), FakeMNIST(256, 1)
        lkw = dict(mc_samples=10, num_epochs=0, data_minibatch=128, D=784, N=len(ltr), inner_it=20, trainer="nested",
                   log_every=150, lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=200,
                   seed=rank, architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=ltr,
                   test_dataset=lte, dnm="MNIST", nc=10, compute_weights_entropy=False, register_elbos=False, quiet=True)
        from psvi.inference.psvi_classes import PSVILearnV
        lobj = PSVILearnV(**lkw)
        lobj.run_psvi(**lkw)
        pc._dist_info = lambda: (None, 0, 1)
        lxb, lyb = lobj._next_minibatch()
        lobj.nested_step(lxb, lyb)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(3):
            lobj.nested_step(lxb, lyb)
        b.record(stream)
        torch.cuda.synchronize()
        lms = a.elapsed_time(b) / 3
        # the same step at T = 100 (a new graph is captured for the new T)
        lobj.inner_it = 100
        lobj.nested_step(lxb, lyb)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(2):
            lobj.nested_step(lxb, lyb)
        b.record(stream)
        torch.cuda.synchronize()
        lms100 = a.elapsed_time(b) / 2
        pc._dist_info = real_dist_info
        lflops = 2.0 * 10 * (28 * 28 * 25 * 6 + 100 * 150 * 16 + 400 * 120 + 120 * 84 + 84 * 10)   # fwd FLOPs per row, S=10
        lfl = 20 * 9 * lflops * 200 + 3 * lflops * 328
        lenet = {"what": "PSVILearnV.nested_step, lenet (P=61 706 per sample), M=200, S=10, B=128, T=20 (and T=100), fp32 "
                         "CUDA-core conv/fc kernels (psvi_lenet_pass), streaming engine, one CUDA graph per step",
                 "ms_per_outer_step": lms, "outer_steps_per_s": 1e3 / lms, "TFLOPs": lfl / (lms * 1e-3) / 1e12,
                 "T100_ms_per_outer_step": lms100}
    except Exception as e:
        lenet = {"error": repr(e)[:300]}

    # -------- cfg3 (fn2: full-covariance BNN, two hidden layers of 40 units, P = 1 356 286; synthetic 2-d data, M=100,
    # S=32, T=20, B=128): one PSVI outer step through the packed-triangle streaming path ---------------------------------
    fn2 = None
    try:
        from psvi.experiments.experiments_utils import SynthDataset as _SD, make_synthetic_rows as _msr
        from psvi.inference.psvi_classes import PSVILearnV as _PL
        X3, Y3 = _msr(100000, 2, 2, seed=0)
        t3, e3 = _SD(X3[:90000], Y3[:90000].float()), _SD(X3[90000:], Y3[90000:].float())
        k3 = dict(mc_samples=32, num_epochs=0, data_minibatch=128, D=2, N=90000, inner_it=20, trainer="nested", log_every=1000,
                  lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=100, seed=rank,
                  architecture="fn2", n_hidden=40, n_layers=1, logistic_regression=False, train_dataset=t3, test_dataset=e3,
                  dnm="synthetic", nc=2, compute_weights_entropy=False, register_elbos=False, quiet=True)
        o3 = _PL(**k3)
        o3.run_psvi(**k3)
        pc._dist_info = lambda: (None, 0, 1)
        x3b, y3b = o3._next_minibatch()
        o3.nested_step(x3b, y3b)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(3):
            o3.nested_step(x3b, y3b)
        b.record(stream)
        torch.cuda.synchronize()
        pc._dist_info = real_dist_info
        fn2 = {"what": "PSVILearnV.nested_step, fn2 dims [2, 40, 40, 2] (1 356 286 variational parameters, packed scale_tril), "
                       "N=90 000, M=100, S=32, B=128, T=20", "ms_per_outer_step": a.elapsed_time(b) / 3,
               "outer_steps_per_s": 3e3 / a.elapsed_time(b)}
    except Exception as e:
        fn2 = {"error": repr(e)[:300]}

    # -------- cfg5 (fn D=256 H=1024 C=10, M=1000, S=64, B=128, T=10): one PSVI outer step through the batched tcgen05 GEMM
    # path of the large regime, in both split-precision arithmetics -----------------------------------
    fn5 = None
    try:
        from psvi.experiments.experiments_utils import SynthDataset as _SD5, make_synthetic_rows as _msr5
        from psvi.inference.psvi_classes import PSVILearnV as _PL5
        X5, Y5 = _msr5(20000, 256, 10, seed=0)
        t5, e5 = _SD5(X5[:16000], Y5[:16000].float()), _SD5(X5[16000:], Y5[16000:].float())
        k5 = dict(mc_samples=64, num_epochs=0, data_minibatch=128, D=256, N=16000, inner_it=10, trainer="nested", log_every=1000,
                  lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=1000, seed=rank,
                  architecture="fn", n_hidden=1024, n_layers=1, logistic_regression=False, train_dataset=t5, test_dataset=e5,
                  dnm="synthetic", nc=10, compute_weights_entropy=False, register_elbos=False, quiet=True)
        fn5 = {"what": "PSVILearnV.nested_step, fn D=256 H=1024 C=10 (P = 273 418 per sample), M=1000, S=64, B=128, T=10; "
                       "batched TMA + tcgen05 GEMMs; algorithmic FLOPs = T 9 F(M) + 3 F(M+B), F(R) = 2 S R (D H + H C); 'mixed' (the default "
                       "arithmetic: tf32x3 gradient passes, split-bf16 Hessian-vector passes) is also timed at T=100"}
        f5 = lambda R_: 2.0 * 64 * R_ * (256 * 1024 + 1024 * 10)
        fl5 = 10 * 9 * f5(1000) + 3 * f5(1128)
        for prec in ("tf32x3", "mixed", "bf16x3"):
            o5 = _PL5(**k5)
            o5.large_precision = prec
            o5.run_psvi(**k5)
            pc._dist_info = lambda: (None, 0, 1)
            x5b, y5b = o5._next_minibatch()
            o5.nested_step(x5b, y5b)
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(2):
                o5.nested_step(x5b, y5b)
            b.record(stream)
            torch.cuda.synchronize()
            pc._dist_info = real_dist_info
            ms5 = a.elapsed_time(b) / 2
            fn5[prec] = {"ms_per_outer_step": ms5, "outer_steps_per_s": 1e3 / ms5, "algorithmic_TFLOPs": fl5 / ms5 / 1e9}
            if prec == "mixed":
                # the unroll length BASELINE configs[4] implies (flow default inner_it = 100): one timed step after one warm-up
                o5.inner_it = 100
                pc._dist_info = lambda: (None, 0, 1)
                o5.nested_step(x5b, y5b)
                torch.cuda.synchronize()
                a.record(stream)
                o5.nested_step(x5b, y5b)
                b.record(stream)
                torch.cuda.synchronize()
                pc._dist_info = real_dist_info
                ms100 = a.elapsed_time(b)
                fn5["mixed_T100"] = {"ms_per_outer_step": ms100, "outer_steps_per_s": 1e3 / ms100,
                                     "algorithmic_TFLOPs": (100 * 9 * f5(1000) + 3 * f5(1128)) / ms100 / 1e9}
            del o5
            torch.cuda.empty_cache()
    except Exception as e:
        fn5 = {"error": repr(e)[:300]}

    # -------- the part of the hot path that SHARDS, at the size BASELINE configs[4] names ("cfg5"): full-data pass of fn
    # D=256 H=1024 C=10 with S=64 MC samples over N = 10 M rows (bf16, 5.12 GB), M=1000 pseudo-points.  STRONG scaling: the
    # 10 M rows are split contiguously over the ranks, every rank runs the tcgen05 forward over its shard and ONE all-reduce
    # of the 8-float result closes the pass.  Timed per pass with CUDA events, median of the passes, max over ranks.
    sharded_fulldata = None
    try:
        import gc
        fx = fy = fscr = None      # release the weak-scaling section's buffers
        gc.collect()
        torch.cuda.empty_cache()
        n_total5 = 10_000_000
        lo5, hi5 = rank * n_total5 // world, (rank + 1) * n_total5 // world
        rows5 = hi5 - lo5
        gg = torch.Generator(device=dev).manual_seed(23 + rank)
        x5 = torch.empty(rows5, Df, device=dev, dtype=torch.bfloat16)
        for r0 in range(0, rows5, 1 << 20):     # generate in chunks: no fp32 copy of the whole shard
            r1 = min(rows5, r0 + (1 << 20))
            x5[r0:r1] = torch.randn(r1 - r0, Df, device=dev, generator=gg, dtype=torch.bfloat16)
        y5 = torch.randint(0, Cf, (rows5,), device=dev, dtype=torch.int32, generator=gg)
        s5 = torch.zeros(_native.fn_tc_scratch_floats(fmodel, rows5, Mf), device=dev)
        o5 = torch.zeros(8, device=dev)

        def pass5():
            _native.fn_predictive_tc(fmodel, fnoise, fmu, frho, fu, fz, fv, x5, y5, 0, float(n_total5), 1, 0.0, 0, o5, s5)
            if world > 1:
                dist.all_reduce(o5)
        pass5()
        barrier()
        per5, ar5 = [], []
        for _ in range(5):
            a.record(stream)
            _native.fn_predictive_tc(fmodel, fnoise, fmu, frho, fu, fz, fv, x5, y5, 0, float(n_total5), 1, 0.0, 0, o5, s5)
            b.record(stream)
            if world > 1:
                dist.all_reduce(o5)
            c2 = torch.cuda.Event(enable_timing=True)
            c2.record(stream)
            torch.cuda.synchronize()
            per5.append(a.elapsed_time(c2))
            ar5.append(b.elapsed_time(c2))
        barrier()
        t5 = torch.tensor([sorted(per5)[len(per5) // 2], sorted(ar5)[len(ar5) // 2]], device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        sec5 = t5[0].item() * 1e-3
        fl5_total = 2.0 * Sf * (n_total5 + world * Mf) * (Df * Hf + Hf * Cf)
        sharded_fulldata = {
            "what": f"cfg5 full-data pass, STRONG scaling: psvi_fn_predictive_tc over {n_total5} bf16 rows x D=256 (fn H=1024, "
                    f"C=10, S=64, M=1000) split over {world} rank(s) ({rows5} rows on rank 0), one all-reduce of 8 floats; "
                    "median of 5 passes, max over ranks",
            "scaling": "strong", "n_gpus": world, "rows_total": n_total5, "ms_per_pass": sec5 * 1e3,
            "passes_per_s": 1.0 / sec5, "row_samples_per_s": n_total5 * Sf / sec5,
            "allreduce_ms": t5[1].item() if world > 1 else 0.0,
            "TFLOPs_aggregate": fl5_total / sec5 / 1e12, "TFLOPs_per_gpu": fl5_total / world / sec5 / 1e12,
            "bound": "tensor", "frac_of_sustained_bf16_peak_per_gpu":
                fl5_total / world / sec5 / 1e12 / json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0)
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None}
        del s5
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        sharded_fulldata = {"error": repr(e)[:300]}

    # -------- the sharded outer-loss GRADIENT at cfg5 (north_star (4)): psvi_elbo value and
    # gradient wrt phi_T, u, v with the data term over the SAME 10 M rows split over the ranks (psvi_fn_data_grad_tc: two
    # tcgen05 passes per shard), pseudo-data terms replicated, and ONE all-reduce of [loss | dL/dphi_T | direct du | direct da].
    sharded_outer_grad = None
    try:
        from psvi.experiments.experiments_utils import SynthDataset as _SDg, make_synthetic_rows as _msrg
        from psvi.inference.psvi_classes import PSVILearnV as _PLg
        Xg, Yg = _msrg(4096, Df, Cf, seed=0)
        tg, eg = _SDg(Xg[:3072], Yg[:3072].float()), _SDg(Xg[3072:], Yg[3072:].float())
        kg = dict(mc_samples=Sf, num_epochs=0, data_minibatch=128, D=Df, N=n_total5, inner_it=1, trainer="nested", log_every=1000,
                  lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=Mf, seed=0,
                  architecture="fn", n_hidden=Hf, n_layers=1, logistic_regression=False, train_dataset=tg, test_dataset=eg,
                  dnm="synthetic", nc=Cf, compute_weights_entropy=False, register_elbos=False, quiet=True)
        og = _PLg(**kg)
        og.run_psvi(**kg)
        eng = og._stream(og.model)
        phi_g = eng.fam.get_phi()
        eps_g = og._noise_tensor(1, eng.Pt, Sf)[0]
        ug, _ = og._uv()
        zg, ag_ = og._z32(), og._a()

        def outer_once():
            lo_, pb_, ub_, ab_, _d = eng.outer_grad(phi_g, eps_g, ug, zg, ag_, None, y5, float(n_total5), kappa=1.0 / world,
                                                    n_total=n_total5, xb_bf16=x5)
            return torch.cat([lo_.reshape(1), pb_.reshape(-1), ub_.reshape(-1), ab_.reshape(-1)]).float()
        flat = outer_once()
        if world > 1:
            dist.all_reduce(flat)
        barrier()
        pg, arg = [], []
        for _ in range(3):
            a.record(stream)
            flat = outer_once()
            b.record(stream)
            if world > 1:
                dist.all_reduce(flat)           # the ONE collective of the sharded outer step (NCCL over NVLink)
            c3 = torch.cuda.Event(enable_timing=True)
            c3.record(stream)
            torch.cuda.synchronize()
            pg.append(a.elapsed_time(c3))
            arg.append(b.elapsed_time(c3))
        barrier()
        tg_ = torch.tensor([sorted(pg)[1], sorted(arg)[1]], device=dev)
        if world > 1:
            dist.all_reduce(tg_, op=dist.ReduceOp.MAX)
        secg = tg_[0].item() * 1e-3
        flg = 3 * 2.0 * Sf * n_total5 * (Df * Hf + Hf * Cf)      # data term: forward + backward = 3 F_fwd(N)
        sharded_outer_grad = {
            "what": f"psvi_elbo value + gradient at cfg5 shapes (fn D=256 H=1024 C=10, S=64, M=1000) with the data term over "
                    f"{n_total5} bf16 rows split over {world} rank(s): pseudo-data passes replicated (tf32x3), data rows through "
                    "psvi_fn_data_grad_tc (bf16 tcgen05, W1 adjoints accumulated in TMEM), ONE all-reduce of "
                    f"{flat.numel()} floats; median of 3, max over ranks",
            "scaling": "strong", "n_gpus": world, "rows_total": n_total5, "ms_per_outer_gradient": secg * 1e3,
            "allreduce_ms": tg_[1].item() if world > 1 else 0.0, "allreduce_bytes": int(flat.numel()) * 4,
            "data_term_algorithmic_TFLOPs_aggregate": flg / secg / 1e12,
            "data_term_algorithmic_TFLOPs_per_gpu": flg / world / secg / 1e12}
    except Exception as e:  # noqa: BLE001
        sharded_outer_grad = {"error": repr(e)[:300]}
    x5 = y5 = None
    torch.cuda.empty_cache()

    # -------- extra: several independent chains on ONE GPU (the reference's multi-trial mode) ------------------------
    # (single-process runs only: a chain built here would take the sharded code path under torch.distributed and wait for
    # collectives the other ranks never join)
    replicas = bench_replicas(c, dev, 8, min(K, 40), 3) if world == 1 else None

    # -------- reduce over ranks ---------------------------------------------------------------------------------------
    tt = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)      # the slowest rank sets the job's time
    value = world * K / (tt[0].item() * 1e-3)
    e2e = world * K / (tt[1].item() * 1e-3)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        fl = flops_outer_step(c)
        ach = fl / (kernel_ms * 1e-3) / 1e12
        sm = _native.lib().psvi_device_sm_count()
        n_cta = c["S"]                                  # one CTA (SM) per MC sample in the cluster
        sm_ghz = (clk.get("sm_mhz") or 1965.0) * 1e-3
        fp32_peak_used = n_cta * 128 * 2 * sm_ghz * 1e-3      # TFLOP/s of the SMs the cluster occupies (128 FMA lanes per SM)
        n_barriers = 4 * c["T"] + 8                     # two cluster barriers per inner step and per reverse step
        barrier_floor_ms = n_barriers * 380.0 / (sm_ghz * 1e6)   # B300_MICROARCH: cluster barrier ~380 cycles
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": tt[0].item() / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": make_config(world),
                "clocks": clk,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": c["B"] * c["D"] * 4 + c["B"] * 4,
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches,
                "roofline": {"bound": "fp32-issue/latency", "achieved": ach, "peak": fp32_peak_used, "unit": "TFLOP/s",
                             "frac": ach / fp32_peak_used, "traffic": None,
                             "kernel": "psvi_mf_fn1_kernel<2,2,7>", "kernel_ms": kernel_ms,
                             "flops_per_launch": fl, "sms_used": n_cta, "sms_total": sm,
                             "peak_source": f"fp32 FMA peak of the {n_cta} SMs the cluster occupies at the sampled SM clock "
                                            f"({sm_ghz:.3f} GHz): {n_cta} x 128 lanes x 2 FLOP",
                             "frac_of_tensor_peak": ach / peak_tf,
                             "tensor_peak": peak_tf,
                             "barrier_floor_ms": barrier_floor_ms,
                             "note": "one cluster launch per outer step; the products have K = D = 2 / N = C = 2 and the step is a "
                                     "chain of 2 T = 200 dependent phases, each closed by two cluster barriers, so neither HBM "
                                     "(0.3 MB of DRAM traffic per launch) nor the tensor pipe can bind (SURVEY.md section 8d): the "
                                     "roof that applies is fp32 issue on the SMs in use plus the barrier chain "
                                     "(barrier_floor_ms = barriers x ~380 cycles); frac_of_tensor_peak is kept because the "
                                     "contract names the tensor roof; ncu: profiles/r2_fn1_engine_ncu_summary.md"},
                "roofline_fulldata": None if not fulldata or "error" in fulldata else {
                    "bound": "hbm", "achieved": fulldata["GBps"] / world, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                    "frac": fulldata["GBps"] / world / peaks.get("hbm_gbs", 6650.0),
                    "traffic": None, "kernel": "psvi_lr_predictive_tc_kernel<12>",
                    "note": "per GPU; algorithmic bytes = rows*(D*2+4) = 4.128 GB per launch; DRAM traffic is not measured in this "
                            "run (one ncu --set full capture of the same launch read 4.135 GB: profiles/r1_lr_tc_ncu_summary.md); "
                            "peak = measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)"},
                "roofline_fn_tc": None if not fn_tc or "error" in fn_tc else {
                    "bound": "tensor", "achieved": fn_tc["TFLOPs_per_gpu"], "peak": peaks.get("bf16_tflops_sustained", 1400.0),
                    "unit": "TFLOP/s", "frac": fn_tc["TFLOPs_per_gpu"] / peaks.get("bf16_tflops_sustained", 1400.0),
                    "traffic": None, "kernel": "psvi_fn_forward_tc_kernel",
                    "note": "per GPU; algorithmic FLOPs = 2 S (rows + M) (D H + H C) with C = 10 (the kernel pads C to 16); "
                            "time = whole psvi_fn_predictive_tc call incl. weight sampling; peak = measured SUSTAINED bf16 "
                            "matmul throughput (the kernel runs for tens of ms back to back); ncu: "
                            "profiles/r1_fn_tc_ncu_summary.md"},
                "sharded_fulldata": sharded_fulldata,
                "sharded_outer_grad": sharded_outer_grad,
                "extra": {"fulldata_fn_tc": fn_tc, "lenet_cfg4": lenet, "fn2_cfg3": fn2, "fn_large_cfg5": fn5, "replicas_one_gpu": replicas, "fulldata_lr_tc": fulldata, "mc_loglik_evals_per_s": {"pseudo_data_elbo_fwd_bwd_fn_M50": inner_evals,
                                                    "full_data_predictive_passes_200rows": pred_evals},
                          "sharded": sharded, "kernel_only_ms": kernel_ms,
                          "per_step_ms_min_med_max": [min(ms), sorted(ms)[len(ms) // 2], max(ms)]}}
        if not args.no_cpu_baseline and world == 1:
            # the reference's own CPU path on this box's host cores, bounded sample (it replaces OUR psvi package in
            # sys.modules, so it runs last); the oracle port stands in only if the reference tree did not travel
            r = reference_sample(8, 1, "cpu", False, budget_s=15.0)
            if "error" not in r:
                line["cpu_baseline"] = {"value": r["steps_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                                        "sample": f"{r['steps']} full cfg2 outer steps of the unmodified reference "
                                                  f"PSVILearnV.nested_step ({r['root']}) on this box's host CPU, "
                                                  f"{r['threads']} torch threads on {r['cores']} cores"}
                g = reference_sample(12, 2, "cuda", False, budget_s=25.0)
                if "error" not in g:
                    line["extra"]["eager_b200"] = {
                        "steps_per_s": g["steps_per_s"], "ms_per_step": g["s_per_step"] * 1e3, "steps": g["steps"],
                        "what": "the same unmodified reference class running eagerly on this GPU (ATen kernels + autograd "
                                "double backward), timed like e2e",
                        "speedup_e2e_over_eager_b200": e2e / g["steps_per_s"]}
                else:
                    line["extra"]["eager_b200"] = g
            else:
                sps, sec = time_cpu(c, 12, 1)
                line["cpu_baseline"] = {"value": sps, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": "12 full cfg2 outer steps of oracle/psvi_oracle.py (numpy fp32, single "
                                                  f"thread): reference tree not importable ({r['error']})"}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
