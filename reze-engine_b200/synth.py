"""Synthetic PMX-shaped workloads: mesh, skeleton, palettes, morphs, SDEF.

Everything is seeded (`numpy.random.default_rng(20260925)` by default) and shaped after
the reference's shipped model (web/app/tutorial/model.json: 28 789 verts, 471 bones;
influence mix 27.7/52.9/12.7/6.7 % for 1/2/3/4 bones, ~16 distinct bones per 256-vertex
tile, 9-15 % of vertices touched by morphs).  Used by bench.py and the parity tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .crowd import tween_pose_batch, world_matrices_batch
from .model import Bone, Model, SdefTable, Skeleton, Skinning, VertexMorphs
from .pmx import compute_inverse_bind

SEED = 20260925
GOLDEN = 0.6180339887498949


def make_skeleton(B: int, rng) -> List[Bone]:
    """Random tree; bone 0 is the root, parents precede children (PMX files mostly do)."""
    bones: List[Bone] = []
    for i in range(B):
        if i == 0:
            parent = -1
        else:
            lo = max(0, i - 24)
            parent = int(rng.integers(lo, i))
        bt = rng.normal(0, 0.6, 3)
        bt[1] = abs(bt[1]) * 0.8
        bones.append(Bone(name=f"bone{i}", parentIndex=parent,
                          bindTranslation=[float(np.float32(x)) for x in bt]))
    return bones


def _neighbours(bones: List[Bone]):
    B = len(bones)
    nb = [[] for _ in range(B)]
    for i, b in enumerate(bones):
        if b.parentIndex >= 0:
            nb[i].append(b.parentIndex)
            nb[b.parentIndex].append(i)
    for i in range(B):
        if not nb[i]:
            nb[i].append((i + 1) % B)
    return nb


def make_mesh(V: int, bones: List[Bone], rng, mix=(0.277, 0.529, 0.127, 0.067), region_mean: float = 96.0,
              set_lo: int = 3, set_hi: int = 6, p_keep_bones: float = 0.8, p_keep_count: float = 0.965, heavy_tail: bool = False):
    """Returns vtx8 [V,8] f32, joints [V,4] u16, weights [V,4] u8 (sum 255).

    Skin-weight structure follows what the reference's shipped model shows when walked in vertex-index order
    (web/app/tutorial/model.json, measured per 32-vertex warp / 256-vertex tile):

        statistic                                   fixture (塞尔凯特.pmx)      this generator (V=20k, B=512)
        influence mix 1/2/3/4                        27.7/52.9/12.7/6.7 %       same (sampled)
        max influence count per warp = 1/2/3/4       14.6/46.7/12.9/25.8 %      ~10/52/23/15 %   (mean 2.50 vs 2.44)
        distinct bones per warp, influence 0/1/2/3   3.5 / 4.5 / 2.4 / 2.1      4.4 / 3.7 / 2.8 / 2.8
        distinct bones per warp (all influences)     5.2                        5.5
        distinct bones per 256-vertex tile           15.6 (p95 61)              15.9 (p95 24)

    i.e. both the bone set AND the influence count are spatially coherent (an i.i.d. influence count would put a
    4-influence vertex into 90 % of all warps, which no real mesh does).  Model: the index range is cut into regions
    (geometric, mean `region_mean`); a region owns a small connected bone set (set_lo..set_hi bones); along the region
    the influence count is kept with probability p_keep_count and the bone tuple with p_keep_bones, else redrawn.
    """
    B = len(bones)
    nb = _neighbours(bones)
    vtx = np.empty((V, 8), np.float32)
    vtx[:, 0] = rng.uniform(-8, 8, V)
    vtx[:, 1] = rng.uniform(0, 22, V)
    vtx[:, 2] = rng.uniform(-2.5, 4, V)
    n = rng.normal(size=(V, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    vtx[:, 3:6] = n
    vtx[:, 6:8] = rng.uniform(0, 1, (V, 2))
    joints = np.zeros((V, 4), np.uint16)
    weights = np.zeros((V, 4), np.uint8)
    mixp = np.asarray(mix, np.float64) / np.sum(mix)
    # heavy_tail: the fixture's tiles are mostly quiet (a limb: ~10 bones per 256 vertices) with a few dense-rig stretches
    # (hair strands, skirt, fingers: many small bones side by side) where one tile sees 60-100 bones -- p95 61 / max 102
    # per tile on the fixture vs p95 26 / max 37 of the default generator.  Zones of ~2 000 vertices are drawn dense with
    # probability 0.065: regions there last ~10 vertices; elsewhere they last 1.5x longer than by default, so the mean stays ~16
    # (measured at V = 100 000, B = 512: mean 15.9, p95 57, max 97 bones per tile; influence mix unchanged).
    zone_len, zone_end, dense = 2000, 0, False
    v = 0
    while v < V:
        if heavy_tail:
            if v >= zone_end:
                dense = rng.random() < 0.065
                zone_end = v + int(rng.uniform(0.5, 1.5) * zone_len)
            rmean = 10.0 if dense else region_mean * 1.5
        else:
            rmean = region_mean
        run = min(int(rng.geometric(1.0 / rmean)), V - v)
        c = int(rng.integers(0, B))
        size = int(rng.integers(set_lo, set_hi + 1))
        S, frontier = [c], [c]
        while len(S) < size and frontier:
            f = frontier.pop(0)
            for nbr in nb[f]:
                if nbr not in S:
                    S.append(nbr)
                    frontier.append(nbr)
                    if len(S) >= size:
                        break
        while len(S) < min(4, B):                     # (skeletons with fewer than 4 bones: fewer influences per vertex)
            x = int(rng.integers(0, B))
            if x not in S:
                S.append(x)
        js = None
        k = 1
        for i in range(v, v + run):
            redraw = js is None
            if js is None or rng.random() > p_keep_count:
                k = min(int(rng.choice(4, p=mixp)) + 1, len(S))
                redraw = True
            if redraw or rng.random() > p_keep_bones:
                js = [S[t] for t in rng.permutation(len(S))[:k]]
            raw = rng.dirichlet(np.ones(k) * 2.0) if k > 1 else np.ones(1)
            w8 = np.maximum(np.floor(raw * 255 + 0.5).astype(np.int64), 1)
            w8[np.argmax(w8)] += 255 - int(w8.sum())
            joints[i, :k] = js
            weights[i, :k] = w8
        v += run
    assert (weights.astype(np.int64).sum(axis=1) == 255).all()
    return vtx, joints, weights


def make_pose_keys(B: int, rng, max_angle: float = 0.6):
    """Two keyframes per bone: identity-ish key A and a random-axis rotation B (angle <= max_angle)."""
    def rnd():
        ax = rng.normal(size=(B, 3))
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        ang = rng.uniform(-max_angle, max_angle, B)
        q = np.concatenate([ax * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
        return q
    return rnd(), rnd()


def make_palettes(bones: List[Bone], P: int, rng, stagger: bool = True, first: int = 0) -> np.ndarray:
    """World matrices [P,B,16] f32: instance p (global id first+p) sits at phase ((first+p)*0.618...) mod 1 of a
    two-key tween evaluated with the reference rule (quadratic ease + slerp) -- "staggered VMD phase"."""
    qa, qb, phase = make_crowd_tween(len(bones), P, rng, stagger=stagger, first=first)
    lr = tween_pose_batch(qa, qb, phase)
    return world_matrices_batch(bones, lr)


def make_crowd_tween(B: int, P: int, rng, stagger: bool = True, first: int = 0):
    """The two-key tween behind make_palettes: (start quats [B,4], target quats [B,4], phase [P] in [0,1))."""
    qa, qb = make_pose_keys(B, rng)
    phase = ((first + np.arange(P)) * GOLDEN) % 1.0 if stagger else np.zeros(P)
    return qa, qb, phase


def make_morphs(V: int, M: int, rng, face_frac: float = 0.12, touch_frac: float = 0.03) -> VertexMorphs:
    face0 = int(V * 0.05)
    faceN = min(max(int(V * face_frac), 8), V - face0)      # (tiny meshes: the face region is the mesh)
    offs = [0]
    vis, dls = [], []
    for m in range(M):
        n = max(int(V * touch_frac * rng.uniform(0.3, 1.6)), 1)
        n = min(n, faceN)
        start = face0 + int(rng.integers(0, faceN - n + 1))
        idx = np.arange(start, start + n)
        # contiguous-ish: drop a few, keep order
        idx = idx[rng.random(n) < 0.9]
        if idx.size == 0:
            idx = np.array([start])
        d = np.clip(rng.normal(0, 0.03, (idx.size, 3)), -0.85, 0.85).astype(np.float32)
        vis.append(idx.astype(np.uint32))
        dls.append(d)
        offs.append(offs[-1] + idx.size)
    return VertexMorphs([f"morph{m}" for m in range(M)], np.asarray(offs, np.uint32),
                        np.concatenate(vis) if vis else np.zeros(0, np.uint32),
                        np.concatenate(dls) if dls else np.zeros((0, 3), np.float32))


def make_sdef(vtx: np.ndarray, weights: np.ndarray, rng, frac: float = 0.2) -> SdefTable:
    two = np.nonzero((weights[:, 2] == 0) & (weights[:, 3] == 0) & (weights[:, 1] > 0))[0]
    pick = two[rng.random(two.size) < frac]
    C = vtx[pick, 0:3] + rng.normal(0, 0.1, (pick.size, 3))
    d = rng.normal(0, 0.2, (pick.size, 3))
    vec = np.concatenate([C, C + d, C - d], axis=1).astype(np.float32)
    return SdefTable(pick.astype(np.uint32), vec, (weights[pick, 0] / 255.0).astype(np.float32))


@dataclass
class Workload:
    vtx8: np.ndarray
    joints: np.ndarray
    weights: np.ndarray
    bones: List[Bone]
    invBind: np.ndarray
    morphs: VertexMorphs
    sdef: SdefTable

    @property
    def V(self):
        return self.vtx8.shape[0]

    @property
    def B(self):
        return len(self.bones)

    def model(self, clock=None) -> Model:
        return Model(self.vtx8.reshape(-1), np.zeros(0, np.uint32), [], [], Skeleton(self.bones, self.invBind),
                     Skinning(self.joints.reshape(-1), self.weights.reshape(-1)), morphs=self.morphs, sdef=self.sdef, clock=clock)


def make_workload(V: int, B: int, M: int = 0, sdef: bool = False, seed: int = SEED, heavy_tail: bool = False) -> Workload:
    rng = np.random.default_rng(seed)
    bones = make_skeleton(B, rng)
    vtx, joints, weights = make_mesh(V, bones, rng, heavy_tail=heavy_tail)
    inv = compute_inverse_bind(bones)
    morphs = make_morphs(V, M, rng) if M else VertexMorphs.empty()
    sd = make_sdef(vtx, weights, rng) if sdef else SdefTable.empty()
    return Workload(vtx, joints, weights, bones, inv, morphs, sd)
