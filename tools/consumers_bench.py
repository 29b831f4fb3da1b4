"""Cost of the fused consumers on the headline mesh (GPU box): plain planar output vs + AABB, + outline hull plane,
interleaved 32-byte stream, positions only.  Device time of rz_deform alone, palettes resident.
Writes gpurun_out/consumers.jsonl."""
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
    ap.add_argument("--verts", type=int, default=100_000)
    ap.add_argument("--bones", type=int, default=512)
    ap.add_argument("--instances", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--shapes", default="0:0:0")
    ap.add_argument("--only", default="", help="substring of the variant name to run alone (ncu captures)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "consumers.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    V, B, K = a.verts, a.bones, a.instances
    wl = synth.make_workload(V, B)
    P = min(K, 1024)
    world = synth.make_palettes(wl.bones, P, np.random.default_rng(1))
    dw = torch.from_numpy(world).cuda()
    i2p = torch.arange(K, dtype=torch.int32, device="cuda") % P
    edge = np.random.default_rng(2).uniform(0.2, 2.0, V).astype(np.float32)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    variants = [("planar pos+nrm", 0), ("+ AABB", capi.RZ_FLAG_BOUNDS), ("+ outline hull plane", capi.RZ_FLAG_OUTLINE),
                ("interleaved [pos,nrm,uv]", capi.RZ_FLAG_INTERLEAVED), ("positions only", capi.RZ_FLAG_NO_NORMALS)]
    for name, flags in variants:
        if a.only and a.only not in name:
            continue
        for sh in a.shapes.split(","):
            I, nt, ctas = (int(x) for x in sh.split(":"))
            row = dict(variant=name, flags=flags, req=[I, nt, ctas])
            try:
                with capi.DeformContext(max_instances=K, stream=stream.cuda_stream, flags=flags, instances_per_group=I, threads=nt,
                                        ctas_per_sm=ctas) as ctx:
                    ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
                    if flags & capi.RZ_FLAG_OUTLINE:
                        ctx.load_edge_size(edge)
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
                    med = float(np.median(ms))
                    row.update(vpl=s["verticesPerLane"], I=s["instancesPerGroup"], threads=s["threads"], ctas=s["ctas"], smem=s["smemBytes"], ms=med,
                               gverts=K * V / med / 1e6, alg_gbs=s["algorithmicBytes"] / med / 1e6,
                               bytes_per_vertex_instance=s["algorithmicBytes"] / (K * V))
            except Exception as e:  # noqa: BLE001
                row["error"] = str(e)
            print(json.dumps(row), flush=True)
            with open(a.out, "a") as f:
                f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
