"""Launch-shape sweep of the deform kernel on the headline workload (GPU box).  Writes gpurun_out/sweep.jsonl."""
import argparse
import itertools
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
    ap.add_argument("--verts", type=int, default=100_000)
    ap.add_argument("--bones", type=int, default=512)
    ap.add_argument("--instances", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--shapes", default="1:256:4,2:256:3,2:256:2,3:256:2,4:256:1,1:512:2,2:512:2,2:512:1,3:512:1,4:512:1,2:768:1,3:768:1,2:1024:1,3:1024:1")
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--pair-coherent", type=int, default=0, help="experiment: make groups of N consecutive vertices share joints (upper bound of lane packing)")
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0, help="rz_config.flags, e.g. 0x8 = RZ_FLAG_REORDER_VERTICES")
    ap.add_argument("--npz", default="", help="real model dumped by tests/golden/make_fixtures.py (e.g. tests/golden/_local/serqet2.npz) instead of the synthetic mesh")
    ap.add_argument("--own-palettes", action="store_true", help="one palette per instance (the headline) instead of 1024 shared ones")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    V, B, K = a.verts, a.bones, a.instances
    wl = synth.make_workload(V, B)
    if a.npz:
        z = np.load(a.npz)
        wl.vtx8, wl.joints, wl.weights, wl.invBind = z["vtx8"], z["joints"], z["weights"], z["invBind"]
        V, B = wl.vtx8.size // 8, wl.invBind.size // 16
        wl.bones = synth.make_workload(64, B).bones          # poses only need a skeleton with B bones
    if a.pair_coherent > 1:
        g = a.pair_coherent
        n = (V // g) * g
        J = wl.joints[:n].reshape(-1, g, 4)
        Wt = wl.weights[:n].reshape(-1, g, 4)
        J[:, 1:, :] = J[:, :1, :]
        Wt[:, 1:, :] = Wt[:, :1, :]
    P = K if a.own_palettes else min(K, 1024)
    world = synth.make_palettes(wl.bones, P, np.random.default_rng(1))
    dw = torch.from_numpy(world).cuda()
    i2p = torch.arange(K, dtype=torch.int32, device="cuda") % P
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rows = []
    L = lambda s: [int(x) for x in s.split(",")]
    shapes = [tuple(int(x) for x in sh.split(":")) for sh in a.shapes.split(",")]
    for shape, chunks in itertools.product(shapes, L(a.chunks)):
        I, nt, ctas = shape[:3]
        st = shape[3] if len(shape) > 3 else 0               # sub-batch of the two-vertex kernel
        vpl = shape[4] if len(shape) > 4 else 0               # 0 auto, 1 one vertex per lane, 2 two
        try:
            ctx = capi.DeformContext(max_instances=K, stream=stream.cuda_stream, instances_per_group=I, threads=nt,
                                     ctas_per_sm=ctas, chunks=chunks, flags=a.flags, store_mode=st, vertices_per_lane=vpl)
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            ctx.set_palettes_device(dw.data_ptr(), P, i2p.data_ptr(), K)
            for _ in range(2):
                ctx.deform()
            torch.cuda.synchronize()
            ms = []
            for _ in range(a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ctx.deform(); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            s = ctx.stats()
            ctx.close()
            med = float(np.median(ms))
            row = dict(vpl=s["verticesPerLane"], I=s["instancesPerGroup"], threads=s["threads"], store=s["storeMode"], ctas=s["ctas"], smem=s["smemBytes"], req=[I, nt, st, ctas, chunks],
                       ms=med, ms_min=float(min(ms)), flags=a.flags, fast_gathers=s["fastGatherPermille"] / 1000, perm=os.environ.get("RZ_PERM", "default"), gverts=K * V / med / 1e6, gbs=s["algorithmicBytes"] / med / 1e6)
        except Exception as e:  # noqa: BLE001
            row = dict(req=[I, nt, st, ctas, chunks], error=str(e))
        rows.append(row)
        print(json.dumps(row), flush=True)
        with open(a.out, "a") as f:
            f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
