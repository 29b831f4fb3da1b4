"""Does the achievable HBM bandwidth itself follow the GPU's clock / power state?  A plain device-to-device copy and a fill
(write-only) of 4 GiB, repeated for ~2 s each, per-iteration device time next to the SM clock and power NVML reports."""
import json
import threading
import time

import torch

samples, stop = [], False


def poll():
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        samples.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.01)


th = threading.Thread(target=poll, daemon=True)
th.start()
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
a.fill_(1.0)
torch.cuda.synchronize()
for name, fn, bytes_ in (("copy (read+write)", lambda: b.copy_(a), 8 * n), ("fill (write only)", lambda: b.fill_(2.0), 4 * n)):
    time.sleep(2.0)                                   # cool down to the idle clock state
    N = 250 if name.startswith("copy") else 450
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    t0 = time.time()
    ev[0].record()
    for i in range(N):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(N)]
    gbs = [bytes_ / m / 1e6 for m in ms]
    ins = [s for s in samples if t0 <= s[0] <= t1]
    print(json.dumps({"what": name, "GBs_first10_median": sorted(gbs[1:11])[5], "GBs_last50_median": sorted(gbs[-50:])[25],
                      "sm_mhz_first": ins[0][1] if ins else None, "sm_mhz_last": ins[-1][1] if ins else None,
                      "power_first": round(ins[0][2]) if ins else None, "power_last": round(ins[-1][2]) if ins else None,
                      "seconds": round(t1 - t0, 2)}), flush=True)
stop = True
