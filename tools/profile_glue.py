"""Times `bya_cfg_dpm_step` at the full latent size (13 x 16 x 60 x 90, CFG batch 2, 48-channel model input) with CUDA
events, flushing the 126 MB L2 between launches (in a real loop 440 ms of transformer step run between two launches, so
its inputs are cold).  Under ncu: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
-k regex:cfg_dpm --csv --log-file gpurun_out/glue.csv python tools/profile_glue.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bya_b200  # noqa: E402,F401
from bya_b200 import ops  # noqa: E402
from bya_b200.scheduler import CogVideoXDPMScheduler  # noqa: E402

F, C, H, W, B = 13, 16, 60, 90, 2
n = F * C * H * W
dev = "cuda"
sch = CogVideoXDPMScheduler.cogvideox_5b()
sch.set_timesteps(50)
ts = sch.timesteps.tolist()
coef = torch.tensor([sch.step_coefficients(t, ts[i - 1] if i else None, i > 0, 6.0) for i, t in enumerate(ts)],
                    dtype=torch.float32, device=dev)
noise = torch.randn(50, 2, n, device=dev, dtype=torch.bfloat16)
out = torch.randn(B, F, C, H, W, device=dev, dtype=torch.bfloat16)
lat = torch.randn(1, F, C, H, W, device=dev, dtype=torch.bfloat16)
pred = torch.zeros(1, F, C, H, W, device=dev, dtype=torch.float32)
x = torch.zeros(B, F, 48, H, W, device=dev, dtype=torch.bfloat16)
idx = torch.full((1,), 7, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for it in range(12):
    flush.fill_(it)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.cfg_dpm_step(out, lat, lat, pred, pred, noise, coef, step_index=idx, model_input=x)
    e.record()
    torch.cuda.synchronize()
    times.append(s.elapsed_time(e) * 1e3)
times = sorted(times[2:])
us = times[len(times) // 2]
# algorithmic bytes of a second-order CFG step: 2 x 2 B model output + 2 B x + 4 B old pred + 2 B noise in;
# 2 B x' + 2 x 2 B model input + 4 B pred out
alg = n * (4 + 2 + 4 + 2 + 2 + 4 + 4)
print(f"cfg_dpm_step: {us:.1f} us median of {len(times)}, algorithmic {alg / 1e6:.1f} MB -> {alg / us / 1e3:.0f} GB/s")
