import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench
from zkm_b200 import lib as zl
lib = zl.init(0)
seg = bench.Segment(lib, "U20")
for _ in range(2): seg.step_device()
def run(tag, prof, sampler):
    lib.zkm_b200_profile_enable(1 if prof else 0)
    seg.profile_reset(); seg.sync()
    if sampler:
        with bench.ClockSampler(0) as clk:
            t = seg.timed(lambda: [seg.step_device() for _ in range(2)])
    else:
        t = seg.timed(lambda: [seg.step_device() for _ in range(2)])
    lib.zkm_b200_profile_enable(0)
    print(tag, "ms/step", t/2, flush=True)
run("plain", 0, 0); run("prof", 1, 0); run("sampler", 0, 1); run("both", 1, 1); run("plain", 0, 0)
# H2D bandwidth
h = torch.empty(1 << 27, dtype=torch.int64, pin_memory=True)   # 1 GiB
d = torch.empty(1 << 27, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("torch pinned H2D GB/s", (1 << 30) / dt / 1e9)
seg.prepare_host()
seg.step_e2e()
t = seg.timed(lambda: seg.step_e2e())
print("e2e ms", t)
os.environ["ZKM_TRACE"] = "1"
seg.step_e2e()
