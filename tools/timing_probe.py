"""Where does a frame's device time go?  Headline workload; CUDA events on the launching stream.
  step      : rz_set_palettes_device + rz_deform per step (graph: skin pass, counter reset, deform)
  deform    : rz_deform only, back to back (graph: counter reset, deform)
  per-launch: rz_get_stats().lastDeformMs after a synchronised single launch
Run once per environment: RZ_NO_GRAPH=1 for direct launches; --double for RZ_FLAG_DOUBLE_BUFFER."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from reze_engine_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--V", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--double", action="store_true")
    a = ap.parse_args()
    wl = synth.make_workload(a.V, 512)
    world = synth.make_palettes(wl.bones, a.K, np.random.default_rng(1))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dw = torch.from_numpy(world).cuda()
    ctx = capi.DeformContext(max_instances=a.K, stream=stream.cuda_stream, flags=capi.RZ_FLAG_DOUBLE_BUFFER if a.double else 0)
    ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def step():
        ctx.set_palettes_device(dw.data_ptr(), a.K)
        ctx.deform()

    out = {"graph": os.environ.get("RZ_NO_GRAPH") is None, "double": a.double, "K": a.K}
    out["step_ms"] = timed(step, a.steps)
    out["deform_only_ms"] = timed(ctx.deform, a.steps)
    out["step_ms_again"] = timed(step, a.steps)
    singles = []
    for _ in range(5):
        ctx.deform()
        singles.append(ctx.stats()["lastDeformMs"])
    out["single_launch_ms"] = singles
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
