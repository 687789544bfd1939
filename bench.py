"""Benchmark of the hot path: BrainEncoder forward + CLIPLoss + backward on synthetic Gwilliams2022-shaped batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config gwilliams] [--impl megb200|reference]

One JSON line on stdout (rank 0).  `value` = train samples/s with the batch resident in HBM; `e2e` = the same step
through the public modules with HOST (pinned) buffers, H2D copies and the loss read-back inside the timed region;
`roofline` = the dominant kernel (tcgen05 tap-GEMM: conv forward + data gradients + 1x1 convs) timed per launch with
CUDA events inside the timed region against the measured bf16 peak; `cpu_baseline` = the oracle port of the reference
on the host cores (bounded sample).  `--impl reference` times that CPU port alone (the reference is CPU-runnable pure
PyTorch; /root/reference does not exist on the GPU box, so its pinned restatement is what runs).
"""
import argparse
import gc
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "meg-decoding_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train samples/sec (fwd+bwd+CLIP)"
UNIT = "samples/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"),
                "hbm": d.get("hbm_gbs"), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.err = index, [], set(), False, None, None
        self.h = None
        try:       # import + NVML init on the caller's thread, BEFORE the timed region (the import holds the GIL for ms)
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if self.h is None:
            return
        try:
            import pynvml
            h = self.h
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step_time(cfg, B_sample, steps, warmup, threads=None, seed=0, keep_first=False):
    """fwd + CLIPLoss + bwd of the oracle port (fp32, all host threads) on the first B_sample samples of the workload's
    batch.  Returns the per-step times; with keep_first also (Z, loss) of the first step (weights at their initial
    values: the parity reference of the bench line)."""
    from megb200 import synth
    from oracle import restatement as R
    if threads:
        torch.set_num_threads(threads)
    P = R.state_to_torch(synth.make_state(cfg), requires_grad=True)
    temp = torch.tensor([float(cfg.init_temperature)], requires_grad=True)
    X, Y, subj = synth.make_batch(cfg, seed)           # the workload's batch; the sample is its first B_sample samples
    X, Y, subj = (torch.from_numpy(a[:B_sample].copy()) for a in (X, Y, subj))
    loc = synth.sensor_locations(cfg.C)
    center = synth.drop_center(cfg)
    times, first = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        Z, loss, _ = R.train_step(P, temp, X, Y, subj, loc, K=cfg.K, d_drop=cfg.d_drop, seq2seq=cfg.seq2seq,
                                  drop_center=center)
        dt = time.perf_counter() - t0
        if i == 0 and keep_first:
            first = (Z, float(loss))
        del Z, loss
        if i >= warmup:
            times.append(dt)
    return (times, first) if keep_first else times


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (its pinned port, oracle/restatement.py --
    /root/reference does not exist on the GPU box and is not pip-installable) on the box's host cores, at the FULL batch
    of the config, with the requested warm-up and step counts."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from megb200 import synth
    cfg = synth.get_config(args.config)
    if args.batch:
        cfg["B"] = args.batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B_sample = min(cfg.B, args.cpu_sample) if args.cpu_sample > 0 else cfg.B
    # safety net only: the whole run must end within a few minutes even on a slow host (a B=256 step is ~3 s on the
    # B200 box); the sample that was actually timed is stated in the line
    probe = cpu_reference_step_time(cfg, B_sample, 1, 0)[0]
    budget = 280.0
    while B_sample > 8 and probe * (args.steps + args.warmup) > budget:
        B_sample //= 2
        probe = cpu_reference_step_time(cfg, B_sample, 1, 0)[0]
    times = cpu_reference_step_time(cfg, B_sample, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    value = B_sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, 1, cfg.B),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{B_sample} of {cfg.B} samples per step (same shapes), {args.warmup} warm-up + "
                                   f"{args.steps} timed steps, fwd+CLIP+bwd, fp32, oracle/restatement.py pinned to the "
                                   f"reference by tests/golden"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(cfg, world, b_loc):
    from megb200 import synth
    return {"workload": f"{cfg.name}: BrainEncoder+CLIPLoss fwd+bwd, C={cfg.C} sensors, S={cfg.num_subjects} subjects, "
                        f"T={cfg.T}, D1={cfg.D1}, D2={cfg.D2}, F={synth.out_features(cfg)}, seq2seq={cfg.seq2seq}",
            "batch_per_gpu": b_loc, "global_batch": b_loc * world,
            "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": "inputs larger than L2: X+Y per step = "
                  f"{(b_loc * cfg.C * cfg.T + b_loc * synth.out_features(cfg) * (cfg.T if cfg.seq2seq else 1)) * 4 / 1e6:.0f} MB f32, "
                  "activations ~3 GB per step"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_megb200(args):
    import torch.distributed as dist
    import megb200
    from megb200 import lib, parallel, synth
    from tests import helpers as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for "
                           "the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        parallel.init()
    megb200.set_precision(args.precision)
    cfg = synth.get_config(args.config)
    if args.batch:
        cfg["B"] = args.batch
    B = cfg.B
    enc, lossf = H.make_modules(cfg, dev)
    enc.train()
    enc.drop_center = synth.drop_center(cfg)
    params = list(enc.parameters()) + list(lossf.parameters())
    from megb200.optim import Adam as FusedAdam
    opt = FusedAdam(params, lr=3e-4)                                      # train.py:160-162, one launch ("next" row N4)
    opt_torch = torch.optim.Adam(params, lr=3e-4)

    # synthetic batch: each rank its own shard (seed = rank), pinned host copies for the end-to-end leg
    Xh, Yh, subj_h = synth.make_batch(cfg, seed=rank)
    Xh, Yh, subj_h = torch.from_numpy(Xh).pin_memory(), torch.from_numpy(Yh).pin_memory(), torch.from_numpy(subj_h)
    Xd, Yd = Xh.to(dev), Yh.to(dev)
    # The speech targets are PRECOMPUTED wav2vec2 features (north_star): the data pipeline stores them once in the form
    # the loss contracts -- position-major bf16 rows + f32 norms (megb200.data.PackedTargets; same rounding point as the
    # dense f32 path, so the same loss).  The CUDA-graph step and the end-to-end leg take them in that form: no per-step
    # transpose/cast pass and half the host->device bytes.  The eager timed region (per-launch roofline events) and the
    # parity step keep the reference collator's dense f32 tensor.
    from megb200.data import PackedTargets
    packed_ok = args.precision.startswith("bf16") and args.targets == "packed"
    if packed_ok:
        Yph = PackedTargets.from_dense(Yh, dtype=torch.bfloat16).pin_memory()
        Ypd = Yph.to(dev)

    def step(X, Y, subj):
        for p in params:
            p.grad = None
        lossf.prefetch_targets(Y)          # target packing (+ all-gather in data parallel) overlaps the encoder forward
        Z = enc(X, subj)
        loss = lossf(Y, Z)
        loss.backward()
        if world > 1:
            parallel.allreduce_grads(params)
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # parity step: the very first step, weights at their initial values; its loss / embeddings are compared with the
    # CPU oracle's first step on the same inputs further down (the `parity` object of the line).  The CPU leg runs the
    # FULL batch when one oracle step is ~4 TFLOP or less (config 2: B = 256), otherwise the first cpu_B samples of it
    # (GOD: 512 of 2048) -- BatchNorm statistics depend on the batch, so the GPU parity step uses the same samples.
    cpu_B = min(B, args.cpu_sample) if args.cpu_sample > 0 else min(B, max(16, int(4.2e12 // synth.encoder_flops_per_sample(cfg)) // 64 * 64))
    for p in params:
        p.grad = None
    Z0 = enc(Xd[:cpu_B].contiguous(), subj_h[:cpu_B])
    loss0 = lossf(Yd[:cpu_B].contiguous(), Z0)
    loss0.backward()
    parity_gpu = (Z0.detach().float().cpu(), float(loss0.detach()))
    del Z0, loss0
    if cpu_B != B:                    # the parity step ran on a prefix: restore the BatchNorm buffers it touched
        sd0 = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state(cfg).items() if "running_" in k or "num_batches" in k}
        enc.load_state_dict(sd0, strict=False)

    # priming (setup, not warm-up): ~1 s of steps so that clocks, the caching allocator and NCCL channels are in steady
    # state; then EXACTLY max(W, 3) warm-up steps, then the K timed steps.  No optimizer step is taken before the timed
    # regions: the timed weights are the initial weights, so `loss` in the line is comparable with the oracle's.
    # The NUMBER of priming steps must be the same on every rank (each step issues collectives): it is derived from the
    # slowest rank's measured step time, not from each rank's own clock -- a per-rank "run for one second" loop lets
    # ranks disagree by a step and deadlocks the job in its next collective (seen at 8 GPUs).
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    for _ in range(4):
        loss = step(Xd, Yd, subj_h)
    torch.cuda.synchronize()
    t_step = max_over_ranks((time.perf_counter() - t_warm) / 4)
    n_prime = 4 + int(min(400, max(2, 1.0 / max(t_step, 1e-4))))
    for i in range(n_prime - 4):
        loss = step(Xd, Yd, subj_h)
        if i % 4 == 3:
            torch.cuda.synchronize()
    barrier()
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        loss = step(Xd, Yd, subj_h)
    barrier()

    if args.ncu_step:
        # exactly one step between cudaProfilerStart/Stop: `ncu --profile-from-start off` captures only these launches
        torch.cuda.profiler.start()
        step(Xd, Yd, subj_h)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"ncu_step": True, "launches": lib.launch_count()}), flush=True)
        return 0

    # ---- timed region: K steps, batch resident in HBM, per-launch CUDA events on the contraction kernels
    ops = megb200.get_ops()
    ops.prof = []
    n0 = lib.launch_count()
    nvml_index = local_rank
    try:
        nvml_index = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank])
    except (ValueError, IndexError):
        pass
    sampler = ClockSampler(nvml_index)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()                 # no collector pauses on the enqueueing thread inside the timed regions
    barrier()
    sampler.start()
    ev0.record()
    th0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(Xd, Yd, subj_h)
    host_ms = (time.perf_counter() - th0) * 1e3 / args.steps      # host time to ENQUEUE a step (no sync inside)
    ev1.record()
    barrier()
    clocks = sampler.result()
    launches = (lib.launch_count() - n0) // args.steps
    ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    prof, ops.prof = ops.prof, None
    loss_value = float(loss.detach())

    kern = {}
    for name, flops, a, b in prof:
        k = kern.setdefault(name, [0.0, 0.0, 0])
        k[0] += flops
        k[1] += a.elapsed_time(b)
        k[2] += 1
    peaks = measured_peaks()
    roofline = None
    # DRAM traffic of the dominant kernel comes from an `ncu --set full` capture (never taken inside this run: profiling
    # replays kernels).  It is only quoted while the kernel sources are the ones the capture was taken on (source hash
    # stored by scripts/ncu_extract.py); a stale capture reads as null.
    traffic, traffic_src, tensor_pipe_pct = None, None, None
    import glob
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        from ncu_extract import source_sha
        cur_sha = source_sha()
    except Exception:
        cur_sha = None
    for tpath in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_tapgemm2h_ncu.json")), reverse=True):
        try:
            d = json.load(open(tpath))
            if d.get("source_sha") is None or d.get("source_sha") != cur_sha:
                traffic_src = f"{os.path.basename(tpath)} is stale (kernel sources changed since the capture)"
                continue
            ls = [e for e in d["launches"] if "dram_read_bytes" in e]
            if ls:
                traffic = sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in ls) / len(ls)
                tp = [e["tensor_pipe_active_pct"] for e in ls if "tensor_pipe_active_pct" in e]
                tensor_pipe_pct = sum(tp) / len(tp) if tp else None
                traffic_src = (f"profiles/{os.path.basename(tpath)}: mean dram read+write bytes per captured launch "
                               f"(ncu --set full, config 2, same kernel sources: {cur_sha})")
                break
        except Exception:
            continue
    if "tapgemm" in kern and kern["tapgemm"][1] > 0:
        fl, tms, n = kern["tapgemm"]
        ach = fl / (tms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tc_tapgemm2h_kernel (k=3 convs fwd + dgrad: 29 of 35 launches) + tc_tapgemm2_kernel (1x1 heads) + tc_tapgemm_kernel (grouped subject GEMM, ragged N)",
                    "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_sustained"], "traffic": traffic,
                    "traffic_source": traffic_src, "ncu_tensor_pipe_active_pct": tensor_pipe_pct,
                    "peak_source": peaks["source"] + " sustained",
                    "launches_per_step": n // args.steps, "avg_launch_ms": tms / n,
                    "share_of_step": tms / args.steps / ms}      # share of the EAGER timed region the events were taken in
        if "wgrad" in kern:
            fl2, t2, n2 = kern["wgrad"]
            roofline["wgrad"] = {"achieved": fl2 / (t2 * 1e-3) / 1e12, "launches_per_step": n2 // args.steps,
                                 "avg_launch_ms": t2 / n2, "share_of_step": t2 / args.steps / ms}

    # ---- the same K steps replayed from a CUDA graph (megb200.graph.GraphedTrainStep: the step recorded once through
    # the same C-ABI launches).  One cudaGraphLaunch per step, so the number does not depend on how fast this host can
    # enqueue ~145 launches; it is the headline `value` when it is available, the eager loop above stays
    # in the line as `ms_per_step_eager` and is where the per-launch roofline events were taken.
    graph_ms, graph_note, eager_ms = None, None, ms
    use_graph = not args.no_graph and (world == 1 or os.environ.get("MEGB200_BENCH_DP_GRAPH", "1") != "0")
    if use_graph:
        gstep, ok = None, 1
        try:
            from megb200.graph import GraphedTrainStep
            gstep = GraphedTrainStep(enc, lossf, Xd, Ypd if packed_ok else Yd, subj_h, drop_center=enc.drop_center)
        except Exception as e:          # capture is an optimisation of the launch path, never a different compute path
            graph_note, ok = f"capture failed: {type(e).__name__}: {e}"[:200], 0
            if os.environ.get("MEGB200_GRAPH_DEBUG"):
                import traceback
                traceback.print_exc()
        if world > 1:                   # all ranks replay or none does: the recording contains collectives
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
            if not ok and graph_note is None:
                graph_note = "capture failed on another rank"
        if not ok:
            gstep = None
    if use_graph and gstep is not None:
        try:
            for _ in range(3):
                gstep()
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for _ in range(args.steps):
                gloss = gstep()
            q1.record()
            barrier()
            graph_ms = max_over_ranks(q0.elapsed_time(q1) / args.steps)
            gl = float(gloss)
            if not math.isfinite(gl) or abs(gl - loss_value) > 0.05 * abs(loss_value):
                graph_note, graph_ms = f"graph loss {gl} disagrees with eager {loss_value}", None
            else:
                launches = gstep.launches_per_step
        except Exception as e:
            graph_note, graph_ms = f"replay failed: {type(e).__name__}: {e}"[:200], None
        if world > 1:                   # one verdict for all ranks
            flag = torch.tensor([1 if graph_ms is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                graph_ms = None
    if graph_ms is not None:
        ms = graph_ms

    # ---- the UNCHANGED reference loop API (enc(X, s); lossf(Y, Z); loss.backward()) with its launches replayed from two
    # CUDA graphs (megb200.graph.ModuleGraphs, opt-in): what `train.py` gets with MEGB200_MODULE_GRAPHS=1
    module_graph_ms, module_graph_note = None, None
    if world == 1 and not args.no_graph:
        try:
            from megb200.graph import accelerate
            runner = accelerate(enc)
            Yt = Ypd if packed_ok else Yd

            def loop_step():
                for p in params:
                    p.grad = None
                Z = enc(Xd, subj_h)
                loss = lossf(Yt, Z)
                loss.backward()
                return loss
            for _ in range(runner.WARMUP + 3):
                lm = loop_step()
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for _ in range(args.steps):
                lm = loop_step()
            m1.record()
            barrier()
            module_graph_ms = m0.elapsed_time(m1) / args.steps
            lmv = float(lm.detach())
            if not math.isfinite(lmv) or abs(lmv - loss_value) > 0.05 * abs(loss_value):
                module_graph_note, module_graph_ms = f"module-graph loss {lmv} disagrees with eager {loss_value}", None
            elif runner.stats["graph_steps"] < args.steps:
                module_graph_note = f"only {runner.stats['graph_steps']} steps were replayed"
        except Exception as e:
            module_graph_note = f"{type(e).__name__}: {e}"[:200]
        finally:
            accelerate(enc, False)

    # ---- optional per-kernel timeline of ONE replayed step on rank 0 (torch.profiler / CUPTI; never part of a timed region)
    if args.timeline and use_graph and gstep is not None and graph_ms is not None:
        # every rank replays (the recording holds collectives); rank 0 does it under the profiler
        if rank != 0:
            gstep()
            torch.cuda.synchronize()
        else:
            try:
                from torch.profiler import ProfilerActivity, profile
                torch.cuda.synchronize()
                with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                    gstep()
                    torch.cuda.synchronize()
                evs = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
                evs.sort(key=lambda e: e.time_range.start)
                t0 = evs[0].time_range.start if evs else 0
                rows = [{"t_us": round(e.time_range.start - t0, 1), "dur_us": round(e.time_range.end - e.time_range.start, 1),
                         "name": e.name[:90]} for e in evs]
                agg = {}
                for r in rows:
                    a = agg.setdefault(r["name"], [0, 0.0])
                    a[0] += 1
                    a[1] += r["dur_us"]
                span = (max(r["t_us"] + r["dur_us"] for r in rows) if rows else 0)
                json.dump({"n_gpus": world, "config": cfg.name, "span_us": span,
                           "by_kernel": sorted(([k, v[0], round(v[1], 1)] for k, v in agg.items()), key=lambda x: -x[2]),
                           "events": rows}, open(args.timeline, "w"))
            except Exception as e:
                print(f"timeline failed: {type(e).__name__}: {e}", file=sys.stderr)
        barrier()

    # ---- optimizer step, reported separately; taken AFTER the timed regions so they see the initial weights
    barrier()
    for _ in range(2):                       # Adam's lazy state init must not land in its own timed region
        opt.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        opt.step()
    e1.record()
    barrier()
    adam_ms = e0.elapsed_time(e1) / 3
    for _ in range(2):
        opt_torch.step()
    barrier()
    e0.record()
    for _ in range(3):
        opt_torch.step()
    e1.record()
    barrier()
    adam_torch_ms = e0.elapsed_time(e1) / 3

    # ---- end to end: host buffers in, loss out, every step.  Public API: megb200.data.DevicePrefetcher (pinned host
    # batches -> device on a copy stream, two batches in flight) + the drop-in modules + loss.item() (train.py:208).
    from megb200.data import AsyncScalarReader, DevicePrefetcher

    def host_batches(n):
        for _ in range(n):
            yield (Xh, Yph.rows, Yph.sumsq, subj_h) if packed_ok else (Xh, Yh, subj_h)

    def unpack(batch):
        if packed_ok:
            X, rows, sumsq, subj = batch
            return X, PackedTargets(rows, sumsq, Yph.feature_shape), subj
        return batch

    prefetcher = DevicePrefetcher(host_batches(0), dev, keep_on_host=(3,) if packed_ok else (2,))
    PRIME = 2
    fill = {}

    gstep_ok = graph_ms is not None

    def e2e_time(lag, fn=None):
        # ONE continuous run of PRIME + K + PRIME steps; the K steps in the middle are timed, so the window is the
        # pipeline's steady state and holds K steps, K batch copies (of the batches two ahead) and K loss reads.
        # lag = 0: `loss.item()` after every step, the reference's loop (train.py:208); lag = 1: the same read
        # pipelined one step behind (megb200.data.AsyncScalarReader) so the stream never drains between steps.
        n = PRIME + args.steps + PRIME
        reader = AsyncScalarReader(dev, lag=lag)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        got, t0, t1 = [], 0.0, 0.0
        barrier()
        ta = time.perf_counter()
        for i, batch in enumerate(prefetcher.reset(host_batches(n))):
            X, Y, subj = unpack(batch)
            if i == PRIME:
                g0.record()
                t0 = time.perf_counter()
            v = reader.push((fn or step)(X, Y, subj))    # device -> host read of the loss, every step
            if v is not None:
                got.append(v)
            if i == PRIME + args.steps - 1:
                g1.record()
                t1 = time.perf_counter()
        got += reader.drain()
        barrier()
        fill[lag] = (time.perf_counter() - ta) * 1e3 / n      # whole run incl. the first batch's exposed copy
        assert len(got) == n and all(math.isfinite(v) for v in got)
        return max_over_ranks(max(g0.elapsed_time(g1), (t1 - t0) * 1e3) / args.steps)

    e2e_time(1)
    # three repetitions of the K-step loop each (PCIe traffic makes single runs noisy); the median is reported
    reps = 1 if args.no_e2e else 3           # --no-e2e: one repetition of each (profiling runs that only need the field)
    sync_runs = [e2e_time(0) for _ in range(reps)]
    pipe_runs = [e2e_time(1) for _ in range(reps)]
    e2e_sync_ms, e2e_ms = statistics.median(sync_runs), statistics.median(pipe_runs)
    e2e_eager_ms, e2e_path, graph_runs = e2e_ms, "eager", None
    if gstep_ok:
        # same loop with the step replayed from the CUDA graph: the prefetched device batch is copied into the graph's
        # static inputs (device -> device, inside the timed window), then one cudaGraphLaunch
        try:
            e2e_time(1, gstep)
            graph_runs = [e2e_time(1, gstep) for _ in range(reps)]
            gsync = statistics.median([e2e_time(0, gstep) for _ in range(reps)])
            if statistics.median(graph_runs) < e2e_ms:
                e2e_ms, e2e_sync_ms, e2e_path = statistics.median(graph_runs), min(gsync, e2e_sync_ms), "cuda_graph"
        except Exception as e:
            graph_note = f"graph e2e failed: {type(e).__name__}: {e}"[:200]
    # the H2D copy of one batch on its own (nothing else running): the floor PCIe sets for the end-to-end step
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        Xd.copy_(Xh, non_blocking=True)
        if packed_ok:
            Ypd.rows.copy_(Yph.rows, non_blocking=True)
            Ypd.sumsq.copy_(Yph.sumsq, non_blocking=True)
        else:
            Yd.copy_(Yh, non_blocking=True)
    c1.record()
    barrier()
    h2d_alone_ms = c0.elapsed_time(c1) / 3
    y_bytes = (Yph.rows.numel() * 2 + Yph.sumsq.numel() * 4) if packed_ok else Yh.numel() * 4
    h2d = Xh.numel() * 4 + y_bytes + subj_h.numel() * 4
    e2e = {"value": B * world / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": e2e_ms, "loss_read": "every step, pipelined one step behind (AsyncScalarReader)",
           "window": "K steps in the steady state of one continuous run (2 priming steps before, 2 after)",
           "ms_per_step_sync_read": e2e_sync_ms, "launch_path": e2e_path, "ms_per_step_eager": e2e_eager_ms,
           "runs_ms_graph": graph_runs, "runs_ms": pipe_runs, "runs_ms_sync_read": sync_runs,
           "ms_per_step_whole_run_incl_fill": fill.get(1), "h2d_alone_ms": h2d_alone_ms, "h2d_alone_GBps": h2d / (h2d_alone_ms * 1e-3) / 1e9,
           "inputs": ("X f32 (B,C,T) + targets as megb200.data.PackedTargets (bf16 position-major rows + f32 norms, the "
                      "stored form of the precomputed features) + int32 subject indices" if packed_ok else
                      "X f32 (B,C,T) + Y f32 (B,F,T) + int32 subject indices (the reference collator's tensors)"),
           "h2d_bytes_per_step_f32_dense_targets": Xh.numel() * 4 + Yh.numel() * 4 + subj_h.numel() * 4}

    # ---- optional extra leg (--e2e-resident): the label side of the dataset resident in HBM (megb200.data.DeviceTargetStore:
    # the precomputed targets of every segment, packed), a batch's targets are a row gather by segment id; per step the
    # host sends the brain data, the segment ids and the subject indices.  Reported as e2e["resident_targets"], never as e2e.value.
    if args.e2e_resident and packed_ok and gstep_ok:
        try:
            from megb200.data import DeviceTargetStore
            nseg = 4 * B
            store = DeviceTargetStore(PackedTargets(Ypd.rows.repeat(4, 1), Ypd.sumsq.repeat(4), Yph.feature_shape), dev)
            ids_h = [torch.randint(0, nseg, (B,), generator=torch.Generator().manual_seed(100 + i)).pin_memory()
                     for i in range(8)]
            pre2 = DevicePrefetcher(iter(()), dev, keep_on_host=(1, 2))      # ids stay on the host: range-checked there, copied by gather()

            def batches2(n):
                for i in range(n):
                    yield (Xh, ids_h[i % len(ids_h)], subj_h)

            def run_resident():
                n = PRIME + args.steps + PRIME
                reader = AsyncScalarReader(dev, lag=1)
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                for i, (X, ids, subj) in enumerate(pre2.reset(batches2(n))):
                    if i == PRIME:
                        g0.record()
                    reader.push(gstep(X, store.gather(ids), subj))
                    if i == PRIME + args.steps - 1:
                        g1.record()
                reader.drain()
                barrier()
                return max_over_ranks(g0.elapsed_time(g1) / args.steps)
            run_resident()
            r_ms = statistics.median([run_resident() for _ in range(reps)])
            e2e["resident_targets"] = {"ms_per_step": r_ms, "value": B * world / (r_ms / 1e3), "unit": UNIT,
                                       "h2d_bytes_per_step": Xh.numel() * 4 + B * 8 + subj_h.numel() * 4,
                                       "store": f"{nseg} segments x {Ypd.rows.shape[1] * 2} B packed bf16 in HBM"}
            del store
        except Exception as e:
            e2e["resident_targets"] = {"error": f"{type(e).__name__}: {e}"[:200]}

    flops_step = B * synth.encoder_flops_per_sample(cfg) + synth.clip_flops(cfg, B, B * world)
    line = {
        "metric": METRIC, "value": B * world / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "priming_steps": n_prime, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision.startswith("bf16") else "f32", "data": "synthetic",
        "config": workload_config(cfg, world, B), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline,
        "step_tflops": flops_step / (ms * 1e-3) / 1e12,
        "step_frac_of_peak": flops_step / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
        "launch_path": "cuda_graph" if graph_ms is not None else "eager", "ms_per_step_eager": eager_ms,
        "targets": "packed bf16 (megb200.data.PackedTargets)" if (packed_ok and graph_ms is not None) else "dense f32",
        "ms_per_step_graph": graph_ms, "graph_note": graph_note,
        "ms_per_step_unchanged_loop_module_graphs": module_graph_ms, "module_graphs_note": module_graph_note,
        "dp_target_gather": (parallel.context().gather_mode if parallel.context() is not None else None),
        "dp_target_gather_note": (parallel.context().gather_note if parallel.context() is not None else None),
        "adam_ms": adam_ms, "adam_ms_torch_optim": adam_torch_ms, "host_enqueue_ms_per_step": host_ms, "loss": loss_value, "precision": args.precision,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU leg: the oracle port at the FULL batch (a B=256 step is ~3 s on the box's cores): 1 warm-up + 2 timed steps.
        # Its first step (initial weights, same inputs) is the reference of the `parity` object.
        Bs = cpu_B
        times, first = cpu_reference_step_time(cfg, Bs, 2, 1, threads=os.cpu_count(), seed=rank, keep_first=True)
        v = Bs / (sum(times) / len(times))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{Bs} of {B} samples per step, 1 warm-up + 2 timed steps, fwd+CLIP+bwd fp32"}
        Zr, lr = first
        zg, lg = parity_gpu
        zerr = float((zg.double() - Zr.double()).norm() / Zr.double().norm())
        tol = 2e-2 if args.precision.startswith("bf16") else 1e-3
        line["parity"] = {"loss_rel": abs(lg - lr) / abs(lr), "z_rel": zerr, "loss": lg, "loss_oracle": lr,
                          "tolerance": tol, "ok": bool(zerr <= tol and abs(lg - lr) / abs(lr) <= tol),
                          "config": f"{cfg.name} B={Bs}" + ("" if Bs == B else f" (the first {Bs} samples of the batch)") +
                                    ": first step at the initial weights vs oracle/restatement.py on the host CPU, same "
                                    "inputs; z_rel = ||Z - Z_ref||_F / ||Z_ref||_F"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown order matters: a CUDA graph that holds NCCL kernels must be destroyed BEFORE the communicator
        # (destroy_process_group otherwise waits forever on it -- measured: the job printed its line and then hung).
        barrier()
        gstep = None          # noqa: F841  (drops the recording)
        prefetcher = None     # noqa: F841
        parallel.shutdown()   # unmaps the peer slabs
        gc.enable()
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        backstop = threading.Timer(20.0, lambda: os._exit(0))      # never let teardown outlive the measurement
        backstop.daemon = True
        backstop.start()
        dist.destroy_process_group()
        backstop.cancel()
    return 0


def main():
    # a hung collective must never burn the GPU budget: dump every thread's stack and exit after --max-seconds
    import faulthandler
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-seconds", type=float, default=1500.0, help="watchdog: dump stacks and exit after this long")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="megb200", choices=["megb200", "reference"])
    ap.add_argument("--config", default="gwilliams")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (default: the config's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16_simt"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="samples per CPU-baseline step (0 = the full batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time only the eager launch path")
    ap.add_argument("--no-e2e", action="store_true", help="one repetition of the end-to-end (host buffers) loops instead of three")
    ap.add_argument("--targets", default="packed", choices=["packed", "dense"],
                    help="target format of the CUDA-graph step and the end-to-end leg: PackedTargets (bf16) or dense f32")
    ap.add_argument("--ncu-step", action="store_true", help="run one step inside cudaProfilerStart/Stop and exit")
    ap.add_argument("--e2e-resident", action="store_true",
                    help="extra end-to-end leg with the targets resident in HBM (DeviceTargetStore); reported beside e2e")
    ap.add_argument("--timeline", default="", help="rank 0 writes a per-kernel timeline of one replayed step to this JSON file")
    args = ap.parse_args()
    if args.max_seconds > 0:
        faulthandler.dump_traceback_later(args.max_seconds, exit=True)
    if args.impl == "reference":
        return run_reference(args)
    return run_megb200(args)


if __name__ == "__main__":
    sys.exit(main())
