#!/usr/bin/env python
"""bench.py -- MSENet14 training throughput (plots/s) on N B200s, BASELINE.json's headline metric.

One "step" = one optimisation step of MSENet14 on a batch of 32 synthetic Danish-NFI-shaped plots per GPU
(BASELINE.json configs[1]): voxel quantisation of the raw points -> coordinate hash -> strided + kernel maps
-> forward -> loss -> backward (dgrad + wgrad) -> [gradient all-reduce] -> AdaBelief.  Nothing is cached
across steps: every step sees a different batch and rebuilds every coordinate structure.

  value  : whole-job plots/s with the raw points already resident in HBM
  e2e    : the same step driven from pinned HOST buffers (H2D of points inside the timed region, loss read back)
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section "Measurement"

`--impl reference` times the CPU arm instead: the reference's MinkowskiEngine CPU build cannot be compiled
here (source not in /root/reference, no network), so it is the oracle port on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "roundup_power2_divisions:8")  # row counts differ every step: bucket sizes so cached blocks are reused
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "MSENet14 train plots/sec"
UNIT = "plots/s"
PLOTS_PER_GPU = 32
POINTS_PER_PLOT = 16000
GRID = 0.0125
BOUNDS = ((0, 0, 0), (80, 80, 100))      # integer grid of positions normalised to [0,1]^2 x [0,1.25]
NUM_DISTINCT_BATCHES = 6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="SENet14")
    ap.add_argument("--plots-per-gpu", type=int, default=PLOTS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-plots", type=int, default=0,
                    help="plots per CPU-arm step; 0 = the GPU arm's batch (--plots-per-gpu): like for like")
    ap.add_argument("--precision", default="bf16x2", choices=["bf16x2", "tf32"],
                    help="operand mode of the tensor-core convolutions: bf16x2 = split-bf16 pairs (default; holds "
                         "1e-3 on every gradient end to end against the fp32 oracle), tf32 = single TF32 operands")
    return ap.parse_args()


def workload(args):
    return {"workload": f"{args.model} biomass-regression training step, batch {args.plots_per_gpu} synthetic "
                        f"NFI-shaped plots per GPU x {POINTS_PER_PLOT} points, GridSampling3D size {GRID} "
                        f"(BASELINE.json configs[1]); raw points -> quantise -> hash/maps -> fwd -> bwd -> AdaBelief",
            "plots_per_gpu": args.plots_per_gpu, "points_per_plot": POINTS_PER_PLOT, "grid_size": GRID,
            "optimizer": "AdaBelief lr 5e-3 wd 1e-2 clip 100 (fused flat buffer)", "parallelism": f"dp{args.gpus}",
            "l2": "distinct batch every step; per-step working set (stem neighbour table alone ~0.6 GB) >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port) -- used for cpu_baseline and for --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_arm(args, steps, warmup, sample_plots):
    from dpcr_agb_b200 import msenet, plots
    from oracle import me_cpu
    from oracle import train as otrain
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = msenet.build(me_cpu, args.model, drop_path=0.01)
    opt = otrain.AdaBelief(model.parameters(), lr=5e-3, weight_decay=1e-2)
    center, scale = torch.tensor([107.0, 200.0]), torch.tensor([103.0, 194.0])
    batches = [plots.synth_batch(2, 1000 + i * sample_plots, sample_plots, n_points=POINTS_PER_PLOT)
               for i in range(min(steps + warmup, 3))]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        otrain.cpu_training_step(model, opt, batches[i % len(batches)], GRID, center, scale)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = float(np.median(times))
    return {"value": sample_plots / med, "unit": UNIT, "cores": cores, "kind": "port",
            "plots_per_step": sample_plots,
            "sample": f"{steps} timed step(s) (median) of a {sample_plots}-plot batch x {POINTS_PER_PLOT} points "
                      f"{'(the GPU arm batch size)' if sample_plots == args.plots_per_gpu else '(SMALLER than the GPU arm batch of ' + str(args.plots_per_gpu) + ')'} "
                      f"after {warmup} warm-up, whole step from raw points (quantise+maps+fwd+bwd+AdaBelief), fp32, "
                      f"oracle restatement of the ME-CPU algorithm on torch-CPU with {cores} threads",
            "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample_plots if args.cpu_sample_plots > 0 else args.plots_per_gpu
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cb = cpu_arm(args, steps, warmup, sample)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload(args),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "plots_per_step")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference MinkowskiEngine (CPU build, env_cpu.yml) is an un-vendored pip dependency and cannot "
                    "be built offline; this arm is the oracle port of its algorithm on the host cores"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.path = tempfile.mktemp(prefix="b2s_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(device_index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    from dpcr_agb_b200 import MinkowskiEngine as ME
    from dpcr_agb_b200 import graph_step, lib, msenet, plots, train
    from dpcr_agb_b200.MinkowskiEngine import functional as Fn
    from dpcr_agb_b200.quantize import GridSampling3D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    lib.set_tuning("precise", 1 if args.precision == "bf16x2" else 0)
    B = args.plots_per_gpu

    torch.manual_seed(0)
    model = msenet.build(ME, args.model, drop_path=0.01).to(dev)
    trainer = train.Trainer(model, ME)
    trainer.broadcast_parameters()
    gs = GridSampling3D(GRID)

    # ---- synthetic input: NUM_DISTINCT_BATCHES different batches per rank, cycled (pinned host + device copies)
    nb = min(NUM_DISTINCT_BATCHES, args.steps + args.warmup)
    host, devb = [], []
    for i in range(nb):
        b = plots.synth_batch(2, (rank * nb + i) * B, B, n_points=POINTS_PER_PLOT)
        h = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in ("pos", "feats", "batch", "perm", "target")}
        host.append(h)
        devb.append({k: v.to(dev) for k, v in h.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def eager_step(d):
        vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=BOUNDS)
        return trainer.step(vox["coords"], vox["tensors"][0], d["target"], dense_index=vox["index"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic work per step (untimed statistics pass over every distinct batch, eager exact-size path)
    Fn.WORK_STATS = {}
    for d in devb:
        eager_step(d)
    torch.cuda.synchronize()
    work = {k: {kk: vv / len(devb) for kk, vv in v.items()} for k, v in Fn.WORK_STATS.items()}
    Fn.WORK_STATS = None

    # ---- the product path: the whole step captured as one CUDA graph at fixed row capacities (graph_step.py)
    caps = graph_step.plan_capacities(gs, ME, model, devb, B, BOUNDS)
    if world > 1:                                       # same capacities on every rank (max over ranks)
        keys = sorted(caps)
        t = torch.tensor([caps[k] for k in keys], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        caps = {k: int(v) for k, v in zip(keys, t.tolist())}
    gstep = graph_step.GraphStep(trainer, gs, B, B * POINTS_PER_PLOT, BOUNDS, caps).capture()

    def step_from_device(d):
        gstep.load(d)
        return gstep.step()

    last_loss = [None]

    def step_from_host(h):
        gstep.load(h)                                             # pinned host -> device copies of this step's inputs
        last_loss[0] = float(gstep.step())                        # device -> host read of the step's result
        return last_loss[0]

    def timed(fn, items, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(items[i % len(items)])
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)               # max over ranks, timed on the device
        return float(ms.item())

    # ---- warm-up
    for i in range(args.warmup):
        step_from_device(devb[i % nb])
    barrier()

    # ---- timed region 1: inputs resident in HBM
    clocks = ClockSampler(local)
    ms_total = timed(step_from_device, devb, args.steps)
    clk = clocks.stop()
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    calls = gstep.launches_per_step * args.steps
    gstep.verify()

    # ---- timed region 2: end to end from pinned host buffers
    for i in range(min(2, args.warmup)):
        step_from_host(host[i % nb])
    ms_e2e = timed(step_from_host, host, args.steps)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)
    gstep.verify()

    # ---- per-entry-point CUDA-event timing on the launching stream (eager exact-size path, same kernels, same
    #      batches): conv kernels alone first (their events do not perturb each other much), then everything
    for d in devb[:2]:
        eager_step(d)
    lib.profile_start(["b2s_conv_gather_gemm", "b2s_conv_wgrad"])
    psteps = min(args.steps, 6)
    for i in range(psteps):
        eager_step(devb[i % nb])
    prof = {k: (n * args.steps / psteps, t * args.steps / psteps) for k, (n, t) in lib.profile_stop().items()}
    lib.profile_start(None)
    bsteps = min(3, args.steps)
    for i in range(bsteps):
        eager_step(devb[i % nb])
    breakdown = {k: {"calls_per_step": n / bsteps, "ms_per_step": t / bsteps} for k, (n, t) in lib.profile_stop().items()}

    def shutdown():
        """Leave the process group without ever hanging the launcher: the captured graphs go first (NCCL does not
        let a communicator die while a graph references it), and a watchdog ends the process with status 0 if the
        teardown still blocks -- every number has been printed by then."""
        if world == 1:
            return
        sys.stdout.flush()
        import threading
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        gstep.release()
        dist.destroy_process_group()
        t.cancel()

    if rank != 0:
        shutdown()
        return

    # ---- roofline of the dominant kernel family (conv gather-GEMM fwd+dgrad; wgrad reported beside it)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_sust = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tcgen05 kind::tf32 runs at half the bf16 rate)"
    if bf16_sust is None:
        bf16_sust, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md) / 2"
    tf32_peak = bf16_sust / 2.0
    kinds = {"fwd": "b2s_conv_gather_gemm:fwd", "dgrad": "b2s_conv_gather_gemm:dgrad", "wgrad": "b2s_conv_wgrad"}
    per_kind = {}
    for kind, key in kinds.items():
        n, t = prof.get(key, (0, 0.0))
        w = work.get(kind, {"flops": 0, "bytes": 0, "launches": 0, "pairs": 0})
        ms = t / args.steps
        per_kind[kind] = {"launches_per_step": n / args.steps, "ms_per_step": ms,
                          "algorithmic_gflop_per_step": w["flops"] / 1e9, "algorithmic_gb_per_step": w["bytes"] / 1e9,
                          "tflops": (w["flops"] / 1e12) / (ms * 1e-3) if ms > 0 else None}
    gg_ms = per_kind["fwd"]["ms_per_step"] + per_kind["dgrad"]["ms_per_step"]
    gg_flops = work.get("fwd", {"flops": 0})["flops"] + work.get("dgrad", {"flops": 0})["flops"]
    gg_launch = per_kind["fwd"]["launches_per_step"] + per_kind["dgrad"]["launches_per_step"]
    achieved = (gg_flops / 1e12) / (gg_ms * 1e-3) if gg_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, weighted by the
    # launches per step of the captured instances); null when the capture file is absent
    traffic, traffic_note = None, None
    try:
        tr_ = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        inst = tr_["instances"]
        n_l = sum(i["launches_per_step"] for i in inst)
        traffic = sum(i["dram_bytes"] * i["launches_per_step"] for i in inst) / n_l
        traffic_note = (f"bytes per launch, mean over {n_l} of the {int(gg_launch)} launches per step "
                        f"({'; '.join(i['instance'] for i in inst)}); {tr_['source']}")
    except Exception:
        pass
    roofline = {"kernel": "gather_gemm_tc_kernel + gather_gemm_tc2_kernel (b2s_conv_gather_gemm: conv forward + dgrad, tcgen05 kind::tf32; M = 128 and M = 256 tiles)",
                "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": achieved / tf32_peak if tf32_peak else None, "traffic": traffic, "traffic_note": traffic_note,
                "algorithmic_bytes_per_launch": (work.get("fwd", {"bytes": 0})["bytes"]
                                                 + work.get("dgrad", {"bytes": 0})["bytes"]) / gg_launch if gg_launch else None,
                "peak_source": peak_src,
                "avg_launch_ms": gg_ms / gg_launch if gg_launch else None, "launches_per_step": gg_launch,
                "algorithmic_gflop_per_launch": gg_flops / 1e9 / gg_launch if gg_launch else None,
                "share_of_step": gg_ms / ms_step if ms_step else None,
                "conv_impl": {0: "auto (tcgen05 where covered, SIMT otherwise)", 1: "SIMT", 2: "tcgen05"}[Fn.CONV_IMPL],
                "per_kind": per_kind}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_arm(args, steps=1, warmup=1,
                     sample_plots=args.cpu_sample_plots if args.cpu_sample_plots > 0 else args.plots_per_gpu)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "plots_per_step")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("bf16x2 (fp32 storage; conv operands as split-bf16 pairs, 3 kind::f16 products, fp32 accumulate)"
                                       if args.precision == "bf16x2" else "tf32 (fp32 storage, fp32 accumulate)"), "data": "synthetic",
            "config": workload(args), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "last_loss": last_loss[0], "gpu_launches": calls, "gpu_launches_note": "C-ABI calls into libb200sparse.so recorded in the captured "
                                                        "step graph x steps (each launches 1-4 kernels of ours)",
            "row_capacities": caps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "breakdown_ms_per_step": breakdown}
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
