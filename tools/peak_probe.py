"""Peak probes BASELINE.md section 2 asks for, written to gpurun_out/r02_peaks.json:
  * dense TF32 and bf16 matmul 8192^3 through cuBLAS (torch.matmul): the library's tensor-core rate for the two operand
    kinds the convolution kernels use (kind::tf32, kind::f16);
  * NCCL all-reduce bus bandwidth at 64 and 256 MB (only under torchrun with >= 2 ranks).
Run:  python tools/peak_probe.py      or      torchrun --nproc-per-node 2 tools/peak_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
out = {}


def mm_rate(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, reps = time.time(), 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / reps
    fl = 2.0 * n ** 3 / 1e12
    return {"burst_tflops": fl / (best * 1e-3), "sustained_tflops": fl / (sustained * 1e-3)}


if rank == 0:
    out["tf32_matmul_8192"] = mm_rate(torch.float32, True)
    out["bf16_matmul_8192"] = mm_rate(torch.bfloat16, False)
    out["fp32_matmul_8192_no_tf32"] = mm_rate(torch.float32, False)
    print(out)

if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    res = {}
    for mb in (64, 256):
        x = torch.ones(mb * 1024 * 1024 // 4, device=dev)
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            dist.all_reduce(x)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t = float(ms.item()) * 1e-3
        algbw = mb * 1024 * 1024 / t / 1e9
        res[f"{mb}MB"] = {"ms": t * 1e3, "algbw_GBps": algbw, "busbw_GBps": algbw * 2 * (world - 1) / world}
    # the exchange of the benchmark: one all-reduce of the flat MSENet14 gradient buffer (14.46 M fp32 = 57.8 MB)
    x = torch.ones(14_457_000, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res["msenet14_gradients_57.8MB"] = {"ms": float(ms.item())}
    out[f"nccl_allreduce_world{world}"] = res
    if rank == 0:
        print(res)
    dist.destroy_process_group()

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/r02_peaks{'_w' + str(world) if world > 1 else ''}.json"
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)
