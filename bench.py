#!/usr/bin/env python
"""Benchmark of the ScalableFHVAE train step (BASELINE.json metric: train segments/sec, fwd+bwd+Adam).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode f32|bf16x3|bf16]

One "step" = the loop body of train_model.py:446-454 on one batch of synthetic 80-dim fbank segments
(SURVEY.md §8d).  Workload at every N: BASELINE config 1 per GPU -- FHVAE (z1 conditioned on z2, 2x256
LSTMs, z dims 32), batch 256 x 20 x 80 per GPU, 1000-row mu2 table -- i.e. weak scaling (config 2 at N=8).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU implementation of the same step
(the oracle port: the reference's FHVAE is a stub, fhvae.py:14) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=256, T=20, F=80, H=256, L=2, Z=32, N=1000, alpha=10.0)
FLOP_PER_SEG_TRAIN = 320_073_216          # SURVEY.md §8d canonical (nn.LSTM/nn.Linear count, N=1000)
METRIC = "train_segments_per_sec"
UNIT = "segments/s"


def synth(B, T, F, N, seed):
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, F, generator=g)
    lens = np.random.default_rng(7).integers(200, 1601, size=N)
    nsegs_u = (lens - 20) // 8 + 1                          # datasets.py:176
    idx = torch.from_numpy(np.random.default_rng(seed).choice(N, size=B, p=nsegs_u / nsegs_u.sum())).long()
    return x, idx, torch.from_numpy(nsegs_u)[idx].long()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_step_time(seconds_budget, threads, steps_cap=200, warmup=2):
    """Oracle port (oracle/fhvae_oracle.py FHVAEOracle, nn.LSTM) of the same step on the host cores."""
    from oracle import fhvae_oracle as O
    torch.set_num_threads(threads)
    c = CFG
    torch.manual_seed(0)
    m = O.FHVAEOracle(c["T"] * c["F"], [c["H"]] * c["L"], [c["H"]] * c["L"], c["Z"], c["Z"], [c["H"]] * c["L"],
                      seg_len=c["T"], num_seqs=c["N"])
    opt = O.make_adam(m.parameters())
    x, idx, nsegs = synth(c["B"], c["T"], c["F"], c["N"], 1234)
    times = []
    t_end = time.perf_counter() + seconds_budget
    i = 0
    while i < warmup + steps_cap:
        t0 = time.perf_counter()
        O.train_step(m, opt, x, idx, c["N"], nsegs, c["alpha"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
        if i > warmup and time.perf_counter() > t_end:
            break
    times.sort()
    return times[len(times) // 2], len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    c = CFG
    # each "step" of this arm = one CPU train step on the same 256-segment batch
    per_step, n = cpu_step_time(seconds_budget=max(10.0, 0.6 * args.steps), threads=threads,
                                steps_cap=args.steps, warmup=min(args.warmup, 3))
    v = c["B"] / per_step
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": min(args.warmup, 3), "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, "cpu"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} timed steps of one {c['B']}-segment batch (median), oracle FHVAEOracle "
                                       "(nn.LSTM restatement; the reference's fhvae.py is a stub)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, mode):
    c = CFG
    return {"workload": f"BASELINE config 1 per GPU: FHVAE LSTM {c['L']}x{c['H']}, z1/z2 dim {c['Z']}, batch "
                        f"{c['B']}x{c['T']}x{c['F']} per GPU, {c['N']}-row mu2 table, fwd+bwd+Adam(lr 1e-3, "
                        f"betas .95/.999), alpha_dis {c['alpha']}",
            "global_batch": c["B"] * n_gpus, "gemm_mode": mode, "parallelism": f"dp{n_gpus}",
            "l2": "no flush: per-step working set (activations+saved gates ~330 MB, params/Adam ~55 MB) exceeds the 126 MB L2"}


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("FHVAE_MODE", "bf16x3"), choices=["f32", "bf16x3", "bf16"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-kernel-family time split to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import pytorch_scalablefhvae_b200 as P
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    c = CFG
    mode = {"f32": P.MODE_F32_SIMT, "bf16x3": P.MODE_BF16X3, "bf16": P.MODE_BF16}[args.mode]
    torch.manual_seed(0)
    m = P.FHVAE(c["T"] * c["F"], [c["H"]] * c["L"], [c["H"]] * c["L"], c["Z"], c["Z"], [c["H"]] * c["L"],
                seg_len=c["T"], num_seqs=c["N"], gemm_mode=mode, use_cuda_graphs=not args.no_graphs).to(dev)
    opt = P.FusedAdam(m.parameters(), lr=1e-3, betas=(0.95, 0.999), grad_scale=1.0 / world)
    x, idx, nsegs = synth(c["B"], c["T"], c["F"], c["N"], 1234 + rank)
    xd, idd, nsd = x.to(dev), idx.to(dev), nsegs.to(dev)
    xh, idh, nsh = x.pin_memory(), idx.pin_memory(), nsegs.pin_memory()
    allreduce = None
    if world > 1:
        from pytorch_scalablefhvae_b200.parallel import DataParallel
        dp = DataParallel(m, opt)                  # broadcasts rank 0's parameters, grad_scale = 1/world
        allreduce = dp.allreduce_                  # ONE NCCL all-reduce of the flat gradient buffer per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = P._lib.load()

    # ---- (1) device-resident inputs: the fused step
    for _ in range(args.warmup):
        m.train_step(xd, idd, nsd, opt, c["alpha"], allreduce=allreduce)
    barrier()
    l0 = lib.fhvae_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        loss = m.train_step(xd, idd, nsd, opt, c["alpha"], allreduce=allreduce)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    final_loss = float(loss)
    # launches per step: counted by the library on an un-graphed replay of the same call lists
    m.use_cuda_graphs = False
    l0 = lib.fhvae_launch_count()
    m.train_step(xd, idd, nsd, opt, c["alpha"], allreduce=allreduce)
    launches_per_step = lib.fhvae_launch_count() - l0
    m.use_cuda_graphs = not args.no_graphs

    # ---- (2) end to end through the reference-facing nn.Module API, pinned host buffers
    def e2e_step():
        opt.zero_grad()
        out = m(xh.to(dev, non_blocking=True), idh, c["N"], nsh)           # H2D inside (train_model.py:444)
        lss = P.loss_function(out[0], out[1], c["alpha"])
        lss.backward()
        if allreduce is not None:
            allreduce(m.packed_grads())
        opt.step()
        return float(lss.detach())                                         # D2H (train_model.py:453)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- (3) dominant-kernel roofline, measured live with CUDA events on the launching stream
    roof, breakdown = dominant_kernel_roofline(m, P, xd, idd, nsd, opt, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        per_step, n = cpu_step_time(seconds_budget=15.0, threads=threads, steps_cap=40)
        cpu = {"value": c["B"] / per_step, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} timed steps of one {c['B']}-segment batch (median {per_step * 1e3:.1f} ms), "
                         "oracle FHVAEOracle nn.LSTM restatement (reference fhvae.py is a stub)"}
    gb = c["B"] * world
    value = gb * args.steps / (ms / 1e3)
    pk = peaks()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"f32": "f32", "bf16x3": "bf16x3 split (fp32-parity)", "bf16": "bf16"}[args.mode],
        "data": "synthetic", "config": workload_config(world, args.mode),
        "e2e": {"value": gb * args.steps / (ms_e2e / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": int(x.numel() * 4 + idx.numel() * 8 + nsegs.numel() * 8),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                "api": "nn.Module forward -> loss_function -> backward -> FusedAdam.step -> loss.item()"},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "step_tflops": FLOP_PER_SEG_TRAIN * value / 1e12,
        "step_frac_of_bf16_sustained": FLOP_PER_SEG_TRAIN * value / 1e12 / pk["tf_sus"],
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "final_loss": final_loss,
        "peaks": pk["src"], "breakdown_ms": breakdown,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(m, P, xd, idd, nsd, opt, args):
    """Time every kernel family of one un-graphed step with CUDA events (same stream), pick the
    dominant one and report its algorithmic work / time against the measured peak."""
    c = CFG
    plan = m._plan(c["B"], c["T"], c["F"])
    was = m.use_cuda_graphs
    m.use_cuda_graphs = False
    fam = {}
    stream = torch.cuda.current_stream().cuda_stream

    def timed_run(cl):
        evs = []
        for f, name, a, _side in cl.calls:
            if name == "join":
                continue
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if f is None:
                a()
            else:
                P._lib.check(f(*a, stream), name)
            e.record()
            evs.append((name, s, e))
        torch.cuda.synchronize()
        for name, s, e in evs:
            d = fam.setdefault(name, [0.0, 0])
            d[0] += s.elapsed_time(e); d[1] += 1

    reps = 3
    for _ in range(reps):
        plan.load_inputs(xd, idd, nsd, None)
        timed_run(plan.fwd)
        plan.gout.zero_(); plan.gout[0].fill_(-1.0 / c["B"]); plan.gout[5].fill_(-c["alpha"] / c["B"])
        plan._gout_rows = plan._gout_train = None          # the plan's own bookkeeping of what gout holds
        if plan.bwd[0] is None:
            plan.bwd[0] = plan._build_bwd(m._grad_buffer(0))
        timed_run(plan.bwd[0])
    m.use_cuda_graphs = was
    breakdown = {k: round(v[0] / reps, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])}
    pk = peaks()
    H, T, B = c["H"], c["T"], c["B"]
    top = max(fam.items(), key=lambda kv: kv[1][0])[0]
    if top.startswith("fhvae_lstm_wave"):
        # one call = one 2-layer stack, T wavefront steps: two recurrent products (h W_hh^T, or dgates W_hh) and the
        # in-kernel cross-layer product (layer-1 input projection / its data gradient), 2*B*4H*H flops each per step.
        # Canonical single-pass count; the bf16x3 mode issues three tcgen05.mma per product.
        calls = fam[top][1] / reps
        ms_call = fam[top][0] / reps / calls
        flops = 3 * 2.0 * B * 4 * H * H * T
        ach = flops / (ms_call * 1e-3) / 1e12
        roof = {"kernel": top + " (one 2-layer LSTM stack, T dependent wavefront steps, 128 CTAs)", "bound": "tensor",
                "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": ach / pk["tf_sus"], "traffic": None,
                "avg_call_ms": ms_call, "peak_kind": "bf16 sustained, " + pk["src"],
                "note": "latency-bound recurrence: T strictly dependent steps of a 256x1024x256 product; "
                        "canonical flops (x3 MMAs executed in bf16x3 mode)"}
    elif top in ("fhvae_lstm_fwd", "fhvae_lstm_bwd"):
        # one call = T recurrent steps of one layer: 2 * B * 4H * H flops per step (fwd: h W_hh^T, bwd: dg W_hh)
        calls = fam[top][1] / reps
        ms_call = fam[top][0] / reps / calls
        flops = 2.0 * B * 4 * H * H * T
        ach = flops / (ms_call * 1e-3) / 1e12
        roof = {"kernel": top + " (one LSTM layer, T recurrent steps)", "bound": "tensor", "achieved": ach,
                "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": ach / pk["tf_sus"], "traffic": None,
                "avg_call_ms": ms_call, "peak_kind": "bf16 sustained, " + pk["src"]}
    else:
        # grouped GEMMs: canonical non-recurrent contraction flops of the step / total GEMM time
        g = fam.get("fhvae_gemm_batch", [0.0, 1])
        rec = 2.0 * 31_457_280 * 3 * B           # recurrent MACs (fwd + 2x bwd) live in the lstm kernels
        flops = FLOP_PER_SEG_TRAIN * B - rec
        ach = flops / (g[0] / reps * 1e-3) / 1e12
        roof = {"kernel": "fhvae_gemm_batch (all grouped GEMM launches of the step)", "bound": "tensor",
                "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": ach / pk["tf_sus"],
                "traffic": None, "peak_kind": "bf16 sustained, " + pk["src"]}
    wg = fam.get("fhvae_wgrad_planes_batch")
    if wg:
        # executed MMA work of the weight gradients: 3 stacks x (3 x [4H,H,TB] + [4H,F or 2F->H...]) -- counted from the
        # plan's problems: per stack 3*H + F (z1,z2) or 3*H + 2F*H/(4H) (decoder head) columns of a [4H x TB] contraction
        cols = 3 * (3 * H) + 2 * c["F"] + (2 * c["F"]) * H / (4.0 * H)
        canon = 2.0 * 4 * H * T * B * cols
        passes = 3 if args.mode == "bf16x3" else 1
        t_s = wg[0] / reps * 1e-3
        roof["wgrad_gemm"] = {"kernel": "wgrad_tma_kernel (fhvae_wgrad_planes_batch, all launches of the step)",
                              "ms_per_step": wg[0] / reps, "canonical_tflops": canon / t_s / 1e12,
                              "executed_mma_tflops": passes * canon / t_s / 1e12,
                              "executed_frac_of_bf16_sustained": passes * canon / t_s / 1e12 / pk["tf_sus"]}
    # HBM-bound kernels of the step (north star item 2: "reported as achieved HBM GB/s"): algorithmic bytes per launch
    #  over the device time of one launch.  A lone eager launch of a 5-15 us kernel cannot be
    # event-timed (the host launch latency lands inside the bracket), so each kernel is replayed REPS times from a
    # CUDA graph; its 5-78 MB working set is then L2-resident, i.e. these are warm-L2 figures (the cold in-step
    # durations are in profiles/r01_step_timeline_cupti.txt: ELBO fwd 6.4 us, bwd 4.1 us, Adam 15.5 us).
    REPS = 20
    Bsz, TF, Z = c["B"], c["T"] * c["F"], c["Z"]
    alg = {"fhvae_elbo_fwd": Bsz * (3 * TF * 4 + 6 * Z * 4 + 20), "fhvae_elbo_bwd": Bsz * (5 * TF * 4 + 10 * Z * 4 + 16)}
    calls = {name: (f, a) for cl in (plan.fwd, plan.bwd[0]) for f, name, a, _s in cl.calls if name in alg}
    flat = m._ensure_flat()
    st_ = opt._state_for(m)
    snap = (flat.clone(), st_["m"].clone(), st_["v"].clone(), st_["step"].clone())
    gbuf = m._grad_buffer(0)
    alg["fhvae_adam_flat"] = 28 * flat.numel()

    def graph_us(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(REPS):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / REPS

    hbm = {}
    for name, nbytes in alg.items():
        if name == "fhvae_adam_flat":
            us = graph_us(lambda: opt.step_flat(m, gbuf))
        elif name in calls:
            f, a = calls[name]
            us = graph_us(lambda: P._lib.check(f(*a, torch.cuda.current_stream().cuda_stream), name))
        else:
            continue
        hbm[name] = {"bytes": nbytes, "us": us, "gbs": nbytes / us / 1e3, "frac_of_hbm": nbytes / us / 1e3 / pk["hbm"],
                     "cache": "warm L2 (graph loop)"}
    flat.copy_(snap[0]); st_["m"].copy_(snap[1]); st_["v"].copy_(snap[2]); st_["step"].copy_(snap[3])
    roof["hbm_kernels"] = hbm
    # DRAM traffic of the dominant kernel: taken from the committed ncu --set full capture of this round
    # (never measured under a profiler here)
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(top)
        if t:
            roof["traffic"] = t["bytes_per_launch"]
            roof["traffic_source"] = "profiles/r01_ncu_traffic.json (ncu dram__bytes_read+write per launch)"
    if args.breakdown:
        sys.stderr.write(json.dumps(breakdown, indent=1) + "\n")
    return roof, breakdown


if __name__ == "__main__":
    main()
