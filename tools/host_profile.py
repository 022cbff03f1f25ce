"""cProfile of the host side of a few training steps (GPU box): where the Python time per step goes."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import lib, msenet, plots, train
from dpcr_agb_b200.quantize import GridSampling3D

dev = torch.device("cuda:0")
B = int(os.environ.get("PLOTS", "32"))
torch.manual_seed(0)
model = msenet.build(ME, "SENet14", drop_path=0.01).to(dev)
tr = train.Trainer(model, ME)
gs = GridSampling3D(0.0125)
batches = []
for i in range(3):
    b = plots.synth_batch(2, i * B, B, n_points=16000)
    batches.append({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()})


def step(d):
    vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=((0, 0, 0), (80, 80, 100)))
    return tr.step(vox["coords"], vox["tensors"][0], d["target"])


for i in range(6):
    step(batches[i % 3])
torch.cuda.synchronize()
import time
st0 = torch.cuda.memory_stats()
t0 = time.perf_counter()
for i in range(6):
    step(batches[i % 3])
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
st1 = torch.cuda.memory_stats()
print(f"[no profiler] alloc_conf={os.environ.get('PYTORCH_CUDA_ALLOC_CONF')} host-side {t_host / 6 * 1e3:.2f} ms/step, "
      f"with final sync {t_all / 6 * 1e3:.2f} ms/step; cudaMalloc calls during 6 steps: "
      f"{st1['num_device_alloc'] - st0['num_device_alloc']}, cudaFree: {st1['num_device_free'] - st0['num_device_free']}, "
      f"reserved {st1['reserved_bytes.all.current'] / 2**30:.1f} GiB, peak allocated {st1['allocated_bytes.all.peak'] / 2**30:.1f} GiB")
if os.environ.get("PROFILE", "1") != "1":
    sys.exit(0)
t0 = time.perf_counter()
N = 6
pr = cProfile.Profile()
pr.enable()
for i in range(N):
    step(batches[i % 3])
pr.disable()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host-side {t_host / N * 1e3:.2f} ms/step, with final sync {t_all / N * 1e3:.2f} ms/step, C-ABI calls/step {lib.launch_count / (N + 3):.0f}")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
