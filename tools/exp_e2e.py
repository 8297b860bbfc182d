import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench
from zkm_b200 import lib as zl
lib = zl.init(0)
seg = bench.Segment(lib, "U20")
for _ in range(2): seg.step_device()
seg.prepare_host()
for _ in range(2): seg.step_e2e()
print("device ms", seg.timed(lambda: seg.step_device()))
print("e2e ms", seg.timed(lambda: seg.step_e2e()))
t0=time.perf_counter(); seg.step_device(); print("device wall ms", (time.perf_counter()-t0)*1e3)
t0=time.perf_counter(); seg.step_e2e(); print("e2e wall ms", (time.perf_counter()-t0)*1e3)
os.environ["ZKM_TRACE"] = "1"
seg.step_e2e()
