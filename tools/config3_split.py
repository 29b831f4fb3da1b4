"""Where config 3's time goes (GPU box): the V=200k mesh timed plain, with morphs only, with SDEF only and with both,
over the launch shapes compiled for the feature kernels.  Parity of the full variant is checked on one instance.
Writes gpurun_out/config3_split.jsonl."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from reze_engine_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--verts", type=int, default=200_000)
    ap.add_argument("--instances", type=int, default=256)
    ap.add_argument("--morphs", type=int, default=64)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--shapes", default="0:0:0,2:256:2,2:512:1,4:512:1,1:256:2")
    ap.add_argument("--variants", default="plain,morph,sdef,both")
    ap.add_argument("--chunks", default="0", help="comma list of rz_config.tune_chunks values (0 = library default)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config3_split.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    V, K, M = a.verts, a.instances, a.morphs
    wl = synth.make_workload(V, 512, M=M, sdef=True)
    rng = np.random.default_rng(3)
    world = synth.make_palettes(wl.bones, K, rng)
    mw = rng.uniform(0, 1, (K, M)).astype(np.float32)
    dw = torch.from_numpy(world).cuda()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ref = {}
    for variant in a.variants.split(","):
        morph = variant in ("morph", "both")
        sdef = variant in ("sdef", "both")
        for sh, chunks in [(x, int(y)) for x in a.shapes.split(",") for y in a.chunks.split(",")]:
            I, nt, ctas = (int(x) for x in sh.split(":"))
            row = dict(variant=variant, req=[I, nt, ctas, chunks])
            try:
                with capi.DeformContext(max_instances=K, stream=stream.cuda_stream, flags=capi.RZ_FLAG_SDEF if sdef else 0,
                                        instances_per_group=I, threads=nt, ctas_per_sm=ctas, chunks=chunks) as ctx:
                    ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
                    if morph:
                        ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
                    if sdef:
                        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
                    ctx.set_palettes_device(dw.data_ptr(), K)
                    if morph:
                        ctx.set_morph_weights(mw, np.arange(M), K=K)
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
                    k = 100 % K
                    if variant not in ref:
                        ref[variant] = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[k], wl.invBind),
                                                  morph=(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta) if morph else None,
                                                  morphW=mw[k] if morph else None,
                                                  sdef=(wl.sdef.vertexIndex, wl.sdef.c_r0_r1) if sdef else None)
                    rp, rn = ref[variant]
                    gp, gn = ctx.read_instance(k)
                    err = max(float(np.abs(gp - rp).max() / max(np.abs(rp).max(), 1.0)), float(np.abs(gn - rn).max()))
                    med = float(np.median(ms))
                    row.update(I=s["instancesPerGroup"], threads=s["threads"], ctas=s["ctas"], smem=s["smemBytes"], ms=med, ms_min=float(min(ms)),
                               gverts=K * V / med / 1e6, gbs=s["algorithmicBytes"] / med / 1e6, max_rel_err=err,
                               fast_gathers=s["fastGatherPermille"] / 1000, sdef=s["sdefCount"], nnz=s["morphNnz"])
            except Exception as e:  # noqa: BLE001
                row["error"] = str(e)
            print(json.dumps(row), flush=True)
            with open(a.out, "a") as f:
                f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
