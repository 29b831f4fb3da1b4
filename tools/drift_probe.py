"""Per-launch device time of the headline deform over a long run + NVML clocks / power sampled alongside: does the
kernel time drift with the GPU's clock / power state?"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from reze_engine_b200 import capi, synth  # noqa: E402

K, V, N = 4096, 100_000, int(os.environ.get("N", "300"))
wl = synth.make_workload(V, 512)
world = synth.make_palettes(wl.bones, K, np.random.default_rng(1))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
dw = torch.from_numpy(world).cuda()
E = lambda k, d=0: int(os.environ.get(k, d))
ctx = capi.DeformContext(max_instances=K, stream=stream.cuda_stream, vertices_per_lane=E("VPL"), instances_per_group=E("IPG"), threads=E("NT"),
                         store_mode=E("SB"), chunks=E("CHUNKS"))
ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
ctx.set_palettes_device(dw.data_ptr(), K)

samples = []
stop = False


def poll():
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        samples.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_MEM),
                        nv.nvmlDeviceGetPowerUsage(h) / 1000.0, nv.nvmlDeviceGetTemperature(h, 0),
                        int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
        time.sleep(0.005)


th = threading.Thread(target=poll, daemon=True)
th.start()
time.sleep(1.0)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
t_begin = time.time()
ev[0].record()
for i in range(N):
    ctx.deform()
    ev[i + 1].record()
torch.cuda.synchronize()
t_end = time.time()
stop = True
th.join()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(N)]
inside = [s for s in samples if t_begin <= s[0] <= t_end]
st = ctx.stats()
print(json.dumps({"vpl": st["verticesPerLane"], "I": st["instancesPerGroup"], "threads": st["threads"], "chunks": E("CHUNKS"),
                  "sm_mhz_last": [s_[1] for s_ in inside[-5:]], "power_last": [round(s_[3]) for s_ in inside[-5:]]}))
print(json.dumps({"ms_first10": ms[:10], "ms_20_30": ms[20:30], "ms_50_60": ms[50:60], "ms_100_110": ms[100:110], "ms_last10": ms[-10:],
                  "mean_first20": float(np.mean(ms[:20])), "mean_last100": float(np.mean(ms[-100:]))}))
step = max(1, len(inside) // int(os.environ.get("ROWS", "8")))
for s in inside[::step]:
    print("t=%.3f sm=%d mem=%d power=%.0fW temp=%d reasons=0x%x" % (s[0] - t_begin, s[1], s[2], s[3], s[4], s[5]))
idle = [s for s in samples if s[0] < t_begin][-3:]
print("idle before:", [(s[1], s[2], round(s[3])) for s in idle])
ctx.close()
