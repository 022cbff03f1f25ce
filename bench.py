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
NUM_DISTINCT_BATCHES = 6

# BASELINE.json configs -> workloads.  cfg2 (the default) is configs[1], the configuration the metric is quoted on;
# cfg4 is cfg2 under torchrun (--gpus N).  bounds = integer voxel grid of positions normalised to [0,1]^2 x [0,1.25].
WORKLOADS = {
    "cfg1": dict(model="SENet14", plots=1, points=16000, grid=0.0125, bounds=((0, 0, 0), (80, 80, 100)), canopy=40.0,
                 train=False, metric="MSENet14 inference plots/sec (one 16k-point plot per call)",
                 what="BASELINE.json configs[0] on the GPU: MSENet14 eval forward of ONE synthetic plot per call"),
    "cfg2": dict(model="SENet14", plots=32, points=16000, grid=0.0125, bounds=((0, 0, 0), (80, 80, 100)), canopy=40.0,
                 train=True, metric=METRIC, what="BASELINE.json configs[1]"),
    "cfg3": dict(model="SENet50", plots=32, points=16000, grid=0.0125, bounds=((0, 0, 0), (80, 80, 100)), canopy=40.0,
                 train=True, metric="MSENet50 train plots/sec", what="BASELINE.json configs[2]"),
    "cfg5": dict(model="SENet50", plots=8, points=200000, grid=0.005, bounds=((0, 0, 0), (200, 200, 250)), canopy=40.0,
                 train=True, metric="MSENet50 stress train plots/sec (200k-point plots, 0.2 m voxels)",
                 what="BASELINE.json configs[4]: 8 plots per GPU (batch 64 across 8 GPUs)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = BASELINE.json configs[1] (default, the headline); cfg1 / cfg3 / cfg5 = configs[0] / [2] / [4]")
    ap.add_argument("--model", default=None, help="override the workload's network (SENet14 / SENet50 / ...)")
    ap.add_argument("--plots-per-gpu", type=int, default=None, help="override the workload's plots per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-plots", type=int, default=0,
                    help="plots per CPU-arm step; 0 = the GPU arm's batch (--plots-per-gpu): like for like")
    ap.add_argument("--precision", default="bf16x2", choices=["bf16x2", "tf32"],
                    help="operand mode of the tensor-core convolutions: bf16x2 = split-bf16 pairs (default; holds "
                         "1e-3 on every gradient end to end against the fp32 oracle), tf32 = single TF32 operands")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.model:
        w["model"] = args.model
        if args.workload == "cfg2" and args.model != "SENet14":
            w["metric"] = f"M{args.model} train plots/sec"
    if args.plots_per_gpu:
        w["plots"] = args.plots_per_gpu
    args.w = w
    args.model, args.plots_per_gpu = w["model"], w["plots"]
    return args


def workload(args, cpu_plots=None):
    """``config`` of the JSON line.  ``cpu_plots``: the CPU arm states the batch it really ran."""
    w = args.w
    step = ("raw points -> quantise -> hash/maps -> fwd -> bwd -> AdaBelief" if w["train"]
            else "raw points -> quantise -> hash/maps -> eval forward -> prediction")
    cfg = {"workload": f"{w['model']} {'biomass-regression training step' if w['train'] else 'inference'}, batch "
                       f"{cpu_plots if cpu_plots is not None else w['plots']} synthetic NFI-shaped plots per "
                       f"{'step (CPU arm)' if cpu_plots is not None else 'GPU'} x {w['points']} points, GridSampling3D "
                       f"size {w['grid']} ({w['what']}); {step}",
           "plots_per_gpu": w["plots"], "points_per_plot": w["points"], "grid_size": w["grid"],
           "optimizer": "AdaBelief lr 5e-3 wd 1e-2 clip 100 (fused flat buffer)" if w["train"] else None,
           "parallelism": f"dp{args.gpus}" + ("" if args.gpus == 1 else
                                              " (plots of every global batch dealt to the ranks by voxel count)"),
           "l2": "distinct batch every step; per-step working set (activations + maps, several GB) >> 126 MB L2"}
    if cpu_plots is not None:
        cfg["plots_per_step_cpu_arm"] = cpu_plots
        cfg["parallelism"] = "host cores"
    return cfg


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port) -- used for cpu_baseline and for --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_arm(args, steps, warmup, sample_plots):
    from dpcr_agb_b200 import msenet, plots
    from oracle import me_cpu
    from oracle import train as otrain
    w = args.w
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = msenet.build(me_cpu, w["model"], drop_path=0.01)
    opt = otrain.AdaBelief(model.parameters(), lr=5e-3, weight_decay=1e-2)
    center, scale = torch.tensor([107.0, 200.0]), torch.tensor([103.0, 194.0])
    batches = [plots.synth_batch(2, 1000 + i * sample_plots, sample_plots, n_points=w["points"],
                                 canopy_max_m=w["canopy"]) for i in range(min(steps + warmup, 3))]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if w["train"]:
            otrain.cpu_training_step(model, opt, batches[i % len(batches)], w["grid"], center, scale)
        else:
            otrain.cpu_inference_step(model, batches[i % len(batches)], w["grid"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = float(np.median(times))
    same = sample_plots == w["plots"]
    return {"value": sample_plots / med, "unit": UNIT, "cores": cores, "kind": "port",
            "plots_per_step": sample_plots,
            "sample": f"{steps} timed step(s) (median) of a {sample_plots}-plot batch x {w['points']} points "
                      f"{'(the GPU arm batch size)' if same else '(SMALLER than the GPU arm batch of ' + str(w['plots']) + ')'} "
                      f"after {warmup} warm-up, whole step from raw points "
                      f"({'quantise+maps+fwd+bwd+AdaBelief' if w['train'] else 'quantise+maps+eval forward'}), fp32, "
                      f"oracle restatement of the ME-CPU algorithm on torch-CPU with {cores} threads",
            "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample_plots if args.cpu_sample_plots > 0 else args.plots_per_gpu
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cb = cpu_arm(args, steps, warmup, sample)
    line = {"impl": "reference", "metric": args.w["metric"], "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload(args, cpu_plots=sample),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "plots_per_step")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference MinkowskiEngine (CPU build, env_cpu.yml) is an un-vendored pip dependency and cannot "
                    "be built offline; this arm is the oracle port of its algorithm on the host cores, ONE process "
                    "whatever --gpus says"}
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
    from dpcr_agb_b200.MinkowskiEngine import coordinate_manager as CM
    from dpcr_agb_b200.MinkowskiEngine import functional as Fn
    from dpcr_agb_b200.quantize import GridSampling3D

    w = args.w
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
    precise = args.precision == "bf16x2"
    lib.set_tuning("precise", 1 if precise else 0)
    B, NPTS, GRID, BOUNDS, TRAIN = w["plots"], w["points"], w["grid"], w["bounds"], w["train"]

    torch.manual_seed(0)
    model = msenet.build(ME, w["model"], drop_path=0.01).to(dev)
    trainer = train.Trainer(model, ME)
    trainer.broadcast_parameters()
    gs = GridSampling3D(GRID)

    # ---- synthetic input: NUM_DISTINCT_BATCHES different batches per rank, cycled (pinned host + device copies)
    nb = min(NUM_DISTINCT_BATCHES, args.steps + args.warmup)
    keys_in = ("pos", "feats", "batch", "perm", "target") if TRAIN else ("pos", "feats", "batch", "perm")
    host, devb = [], []
    balance = None
    for i in range(nb):
        if world == 1:
            b = plots.synth_batch(2, i * B, B, n_points=NPTS, canopy_max_m=w["canopy"])
        else:
            # data-parallel step i: a global batch of world x B plots, dealt to the ranks by voxel count so that the
            # per-rank work (and with it the max-over-ranks step time) stays within ~1 % (train.shard_plots);
            # every rank computes the same deal from the same seeds
            ids = list(range(i * world * B, (i + 1) * world * B))
            weights = [plots.voxel_count(plots.synth_plot(2000 + p, NPTS, w["canopy"])[0], GRID) for p in ids]
            mine = train.shard_plots(weights, rank, world)
            loads = [sum(weights[j] for j in train.shard_plots(weights, r, world)) for r in range(world)]
            balance = max(loads) / (sum(loads) / world) if balance is None else max(balance, max(loads) / (sum(loads) / world))
            b = plots.synth_batch_from_ids(2, [ids[j] for j in mine], n_points=NPTS, canopy_max_m=w["canopy"])
        h = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in keys_in}
        host.append(h)
        devb.append({k: v.to(dev) for k, v in h.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def eager_step(d):
        vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=BOUNDS)
        if TRAIN:
            return trainer.step(vox["coords"], vox["tensors"][0], d["target"], dense_index=vox["index"])
        model.eval()
        with torch.no_grad():
            return model(ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"],
                                         dense_index=vox["index"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic work per step (untimed statistics pass over every distinct batch, eager exact-size path):
    #      conv FLOPs / bytes per launch kind (Fn.WORK_STATS) and the integer stages' bytes (CM.MAP_STATS)
    Fn.WORK_STATS, CM.MAP_STATS = {}, {}
    rows_seen = {}
    for d in devb:
        eager_step(d)
    torch.cuda.synchronize()
    work = {k: {kk: vv / len(devb) for kk, vv in v.items()} for k, v in Fn.WORK_STATS.items()}
    map_work = {k: {kk: vv / len(devb) for kk, vv in v.items()} for k, v in CM.MAP_STATS.items()}
    Fn.WORK_STATS, CM.MAP_STATS = None, None

    # ---- the product path: the whole step captured as one CUDA graph at fixed row capacities (graph_step.py)
    caps = graph_step.plan_capacities(gs, ME, model, devb, B, BOUNDS)
    if world > 1:                                       # same capacities on every rank (max over ranks)
        keys = sorted(caps)
        t = torch.tensor([caps[k] for k in keys], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        caps = {k: int(v) for k, v in zip(keys, t.tolist())}
    # training: the step pipelined across batches (the coordinate graph of batch i+1 replays beside the training graph
    # of batch i: graph_step.PipelinedGraphStep); B2S_PIPELINE=0 selects the single-graph step
    PIPE = TRAIN and os.environ.get("B2S_PIPELINE", "1") == "1"
    if PIPE:
        gstep = graph_step.PipelinedGraphStep(trainer, gs, B, B * NPTS, BOUNDS, caps).capture()
    elif TRAIN:
        gstep = graph_step.GraphStep(trainer, gs, B, B * NPTS, BOUNDS, caps).capture()
    else:
        gstep = graph_step.GraphForward(model, ME, gs, B, B * NPTS, BOUNDS, caps).capture()

    def step_from_device(d):
        gstep.load(d)
        return gstep.step()

    def run_pipelined(items, steps, read_loss):
        """steps training steps, every batch fed one step ahead; with ``read_loss`` the host reads the loss of step
        i-1 after launching step i (every step's loss is read before the function returns)."""
        gstep.reset_feed()
        gstep.feed(items[0])
        pending = None
        for i in range(steps):
            gstep.feed(items[(i + 1) % len(items)])
            if read_loss:
                handle = gstep.step_async()
                if pending is not None:
                    last_out[0] = pending.result()
                pending = handle
            else:
                gstep.step()
        if pending is not None:
            last_out[0] = pending.result()

    last_out = [None]

    def step_from_host(h):
        gstep.load(h)                                             # pinned host -> device copies of this step's inputs
        if TRAIN:
            last_out[0] = float(gstep.step())                     # device -> host read of the step's result (loss)
        else:
            gstep.step()
            gstep.pred_host.copy_(gstep.pred)                     # device -> host read of the predictions (blocking)
            last_out[0] = float(gstep.pred_host[0, 0])
        return last_out[0]

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

    def timed_pipelined(items, steps, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_pipelined(items, steps, read_loss)
        torch.cuda.current_stream().wait_stream(gstep.prep_stream)      # the one coordinate graph fed ahead of the end
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up
    if PIPE:
        run_pipelined(devb, args.warmup, False)
    else:
        for i in range(args.warmup):
            step_from_device(devb[i % nb])
    barrier()

    # ---- timed region 1: inputs resident in HBM
    ncu_range = os.environ.get("B2S_NCU_RANGE") == "1"     # tools/gpu_ncu.sh: profile exactly these replays
    if ncu_range:
        torch.cuda.profiler.start()
    clocks = ClockSampler(local)
    ms_total = timed_pipelined(devb, args.steps, False) if PIPE else timed(step_from_device, devb, args.steps)
    clk = clocks.stop()
    if ncu_range:
        torch.cuda.profiler.stop()
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    calls = gstep.launches_per_step * args.steps
    status = gstep.verify()

    # ---- timed region 2: end to end from pinned host buffers.  Every step's inputs are copied host -> device inside
    #      the timed region and its result is read back; for the training step the copy of step i+1 is started (on a
    #      copy stream, into staging buffers) while step i computes (GraphStep.prefetch / take_prefetched), and the
    #      host reads the loss of step i-1 after launching step i (GraphStep.step_async): the host never runs more than
    #      one step ahead, and every step's inputs and result cross PCIe inside the timed region
    if PIPE:
        run_pipelined(host, min(2, args.warmup), True)
        ms_e2e = timed_pipelined(host, args.steps, True)
        h2d_bytes = h2d_bytes * (args.steps + 1) / args.steps     # one more batch is fed than steps are run
    else:
        for i in range(min(2, args.warmup)):
            step_from_host(host[i % nb])
    if PIPE:
        pass
    elif TRAIN:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gstep.prefetch(host[0])
        pending = None
        for i in range(args.steps):
            gstep.take_prefetched()
            gstep.prefetch(host[(i + 1) % nb])
            handle = gstep.step_async()                           # replay + asynchronous device -> host copy of the loss
            if pending is not None:
                last_out[0] = pending.result()                    # host reads step i-1's loss while step i runs
            pending = handle
        last_out[0] = pending.result()                            # ... and the last step's before the clock stops
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
        h2d_bytes = h2d_bytes * (args.steps + 1) / args.steps     # one more batch is staged than steps are run
    else:
        ms_e2e = timed(step_from_host, host, args.steps)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)
    gstep.verify()

    # ---- per-entry-point CUDA-event timing on the launching stream (eager exact-size path, same kernels, same
    #      batches): the convolution entry points alone first (their events do not perturb each other much), then all
    CONV_FWD = ("b2s_conv_gather_gemm:fwd", "b2s_conv_lines_fwd")
    CONV_DGRAD = ("b2s_conv_gather_gemm:dgrad", "b2s_conv_dgrad_strided")
    CONV_WGRAD = ("b2s_conv_wgrad", "b2s_conv_lines_wgrad")
    for d in devb[:2]:
        eager_step(d)
    lib.profile_start(["b2s_conv_gather_gemm", "b2s_conv_wgrad", "b2s_conv_dgrad_strided", "b2s_conv_lines_fwd",
                       "b2s_conv_lines_wgrad"])
    psteps = min(args.steps, 6)
    for i in range(psteps):
        eager_step(devb[i % nb])
    prof = {k: (n / psteps, t / psteps) for k, (n, t) in lib.profile_stop().items()}     # per step
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

    # ---- rooflines.  Tensor: the contraction's algorithmic FLOPs (2 P c_in c_out per pass) against the kind::tf32
    #      rate -- the rate at which ONE product per term would run -- taken as half the measured sustained bf16 rate.
    #      In split-bf16 mode every term costs three kind::f16 products (five for the stem), so the tensor pipe executes
    #      ~3x the algorithmic FLOPs at twice that rate: `executed_frac_of_bf16_peak` reports that side.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_sust = peaks.get("bf16_tflops_sustained")
    hbm_peak = peaks.get("hbm_gbs")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tcgen05 kind::tf32 runs at half the bf16 rate)"
    if bf16_sust is None:
        bf16_sust, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md) / 2"
    hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    if hbm_peak is None:
        hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    tf32_peak = bf16_sust / 2.0

    def family(keys, kinds):
        n = sum(prof.get(k, (0, 0.0))[0] for k in keys)
        ms = sum(prof.get(k, (0, 0.0))[1] for k in keys)
        fl = sum(work.get(k, {"flops": 0})["flops"] for k in kinds)
        by = sum(work.get(k, {"bytes": 0})["bytes"] for k in kinds)
        return {"entry_points": {k: {"launches_per_step": prof[k][0], "ms_per_step": prof[k][1]} for k in keys if k in prof},
                "launches_per_step": n, "ms_per_step": ms, "algorithmic_gflop_per_step": fl / 1e9,
                "algorithmic_gb_per_step": by / 1e9, "tflops": (fl / 1e12) / (ms * 1e-3) if ms > 0 else None}

    per_kind = {"fwd": family(CONV_FWD, ("fwd",)), "dgrad": family(CONV_DGRAD, ("dgrad",)),
                "wgrad": family(CONV_WGRAD, ("wgrad",))}
    gg = family(CONV_FWD + CONV_DGRAD, ("fwd", "dgrad"))
    achieved = gg["tflops"] or 0.0
    # DRAM traffic of the captured instances of the family (ncu --set full, cold cache, one launch each), PER INSTANCE
    # next to that instance's algorithmic bytes; `traffic` = their mean per launch
    traffic, traffic_instances, traffic_note = None, None, None
    try:
        tr_ = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_instances.json")))
        inst = [i for i in tr_["instances"] if i["tag"].startswith(("conv fwd", "conv dgrad"))]
        traffic_instances = [{"instance": i["tag"], "dram_bytes": i["dram_bytes"], "algorithmic_bytes": i["alg_bytes"],
                              "ratio": i["dram_over_algorithmic"], "time_us": i["time_us"],
                              "tensor_pipe_pct": i["top_kernel_tensor_pipe_pct"]} for i in inst]
        traffic = sum(i["dram_bytes"] for i in inst) / len(inst)
        traffic_note = (f"mean over the {len(inst)} captured conv fwd/dgrad instances, one launch each, listed in "
                        f"traffic_instances with their own algorithmic bytes; {tr_['source']} "
                        f"(profiles/r02_ncu_instances.json)")
    except Exception:
        pass
    launches = gg["launches_per_step"]
    roofline = {"kernel": "conv forward + dgrad: gather_gemm_tc_kernel / gather_gemm_tc2_kernel (b2s_conv_gather_gemm), "
                          "the parity-plan dgrad of the stride-2 convolutions (b2s_conv_dgrad_strided) and the x-line stem "
                          "forward conv_lines_fwd_tmem_kernel (b2s_conv_lines_fwd); tcgen05, fp32 accumulation in TMEM",
                "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": achieved / tf32_peak if tf32_peak else None, "traffic": traffic, "traffic_note": traffic_note,
                "traffic_instances": traffic_instances,
                "algorithmic_bytes_per_launch": gg["algorithmic_gb_per_step"] * 1e9 / launches if launches else None,
                "peak_source": peak_src,
                "avg_launch_ms": gg["ms_per_step"] / launches if launches else None, "launches_per_step": launches,
                "algorithmic_gflop_per_launch": gg["algorithmic_gflop_per_step"] / launches if launches else None,
                "share_of_step": gg["ms_per_step"] / ms_step if ms_step else None,
                "operand_mode": args.precision,
                "executed_frac_of_bf16_peak": (3.0 * achieved / bf16_sust) if precise else (achieved / tf32_peak),
                "executed_note": "split-bf16: three kind::f16 products per term (five in the stem) -> executed tensor "
                                 "FLOPs ~ 3 x algorithmic, against the measured sustained bf16 rate" if precise else
                                 "tf32: one kind::tf32 product per term",
                "conv_impl": {0: "auto (tcgen05 where covered, SIMT otherwise)", 1: "SIMT", 2: "tcgen05"}[Fn.CONV_IMPL],
                "per_kind": per_kind}
    wg = per_kind["wgrad"]
    roofline_wgrad = {"kernel": "conv weight gradient: wgrad_group_kernel / wgrad_small (b2s_conv_wgrad) and the x-line "
                                "stem wgrad_lines_kernel (b2s_conv_lines_wgrad)",
                      "bound": "tensor", "achieved": wg["tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                      "frac": (wg["tflops"] / tf32_peak) if wg["tflops"] else None,
                      "launches_per_step": wg["launches_per_step"], "ms_per_step": wg["ms_per_step"],
                      "share_of_step": wg["ms_per_step"] / ms_step if ms_step else None, "peak_source": peak_src}

    # ---- HBM rooflines of the integer stages: algorithmic bytes / CUDA-event time of their entry points
    def bd(*names):
        return sum(breakdown.get(n, {"ms_per_step": 0.0})["ms_per_step"] for n in names)

    n_pts = B * NPTS
    m_rows = status.get("rows at tensor stride 1", 0) or 0
    stages = []

    def hbm_stage(stage, names, nbytes, note):
        ms = bd(*names)
        if ms > 0 and nbytes:
            gbps = nbytes / (ms * 1e-3) / 1e9
            stages.append({"stage": stage, "entry_points": list(names), "ms_per_step": ms,
                           "algorithmic_mb_per_step": nbytes / 1e6, "achieved": gbps, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gbps / hbm_peak, "bound": "hbm", "note": note})

    hbm_stage("voxel quantisation", ("b2s_quantize_points", "b2s_quantize_count", "b2s_quantize_fill", "b2s_gather_rows"),
              16 * n_pts + 20 * m_rows + 2 * 8 * 3 * m_rows,
              "12n + 4n (perm) + 12M + 8M + feature gather 4F(M + M) for x and pos (F = 3); M = largest voxel count of "
              "the timed steps")
    for name, note in (("b2s_kernel_map_lines", "x-line table of the k7 stem: 16 N_out + 8 P"),
                       ("b2s_kernel_map_dense", "kernel maps through the occupancy index (max pool): 16 N_out + 8 P"),
                       ("b2s_kernel_map", "hash-probed kernel maps and transposed tables: 16 N_out + 8 P")):
        if name in map_work:
            hbm_stage(name, (name,), map_work[name]["bytes"], note)
    if "strided map (b2s_coordmap_insert + _fill)" in map_work:
        hbm_stage("strided maps (coordinate hash insert + fill)", ("b2s_coordmap_insert", "b2s_coordmap_fill"),
                  map_work["strided map (b2s_coordmap_insert + _fill)"]["bytes"], "16 N_in + 16 N_out + 4 N_in per level")
    roofline_hbm = {"peak_source": hbm_src, "stages": stages,
                    "note": "entry-point times are CUDA events on the launching stream in the eager pass (launch latency "
                            "of the 1-4 small kernels behind an entry point included); ncu per-kernel DRAM bytes of one "
                            "instance per stage: profiles/r02_ncu_instances.txt"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_arm(args, steps=1, warmup=1,
                     sample_plots=args.cpu_sample_plots if args.cpu_sample_plots > 0 else args.plots_per_gpu)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "plots_per_step")}

    line = {"metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("bf16x2 (fp32 storage; conv operands as split-bf16 pairs, 3 kind::f16 products, fp32 accumulate)"
                                       if precise else "tf32 (fp32 storage, fp32 accumulate)"), "data": "synthetic",
            "config": workload(args), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4 if TRAIN else 8 * B, "ms_per_step": ms_e2e / args.steps,
                    "how": ("pinned host batch of step i+1 -> input buffers of the other buffer set + its coordinate graph "
                            "on a second stream while step i trains -> training graph -> loss copied to pinned host "
                            "memory, read by the host one step later") if PIPE else
                           (("pinned host batch -> staging buffers on a copy stream while the previous step computes -> "
                             "captured step -> loss copied to pinned host memory, read by the host one step later")
                            if TRAIN else "pinned host batch -> captured forward -> predictions read back (blocking)")},
            "step_form": "pipelined across batches (coordinate graph of batch i+1 beside the training graph of batch i)"
                         if PIPE else "single captured graph",
            "last_result": last_out[0], "gpu_launches": calls, "gpu_launches_note": "C-ABI calls into libb200sparse.so recorded in the captured "
                                                          "step graph x steps (each launches 1-4 kernels of ours)",
            "row_capacities": caps, "rank_work_imbalance_max_over_mean": balance,
            "roofline": roofline, "roofline_wgrad": roofline_wgrad, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu_baseline, "breakdown_ms_per_step": breakdown}
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
