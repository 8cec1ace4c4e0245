"""CPU model of the BVH8 walk of trace_engine.cuh (test infrastructure; runs without a GPU).

It decodes the 80-byte nodes exactly the way the kernel does — per-ray power-of-two time scale folded into 1/dir, near
planes through the 2^15 byte trick with the addend nudged towards the origin and clamped to [0, 1] (FFMA.SAT), far
planes plain, the 8-bit child hit mask turned into (a) the front-to-back mask of internal children through the
octant permutation table and (b) the triangle mask through the 3-bits-per-child expansion table AND Node8::triMask,
triangle slots addressed by popcount(triMask below the bit) — and culls with the closest t found so far (the kernel's
TRACE_CLOSEST_T mode).  tests/test_bvh8_host.py compares it with a brute-force loop over all triangles: same
(triangle, t, u, v), bit for bit.  It validates the LAYOUT CONTRACT between the builders and the kernel; the CUDA code
itself is checked on the GPU against the oracle (tests/test_gpu_parity.py).
"""
import numpy as np

F = np.float32


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


EXPAND = np.array([sum(7 << (3 * i) for i in range(8) if (h >> i) & 1) for h in range(256)], np.uint32)
PERM = np.array([[sum(1 << (i ^ o) for i in range(8) if (h >> i) & 1) for h in range(256)] for o in range(8)], np.uint32)


def moller_trumbore(o, d, v0, e1, e2):
    """S/Tri.hpp:38-68 in uncontracted float32 (traverse.cuh: mollerTrumbore), vectorised over triangles.
    Returns (hit mask, t, u, v)."""
    def cross(a, b):
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(F)

    def dot(a, b):
        return ((a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]).astype(F) + a[..., 2] * b[..., 2]).astype(F)
    with np.errstate(all="ignore"):
        pvec = cross(d[None, :], e2)
        det = dot(e1, pvec)
        inv = (F(1.0) / det).astype(F)
        ok = ~((det > F(-1e-7)) & (det < F(1e-7)))
        tvec = (o[None, :] - v0).astype(F)
        u = (dot(tvec, pvec) * inv).astype(F)
        ok &= ~((u < 0) | (u > 1))
        qvec = cross(tvec, e1)
        v = (dot(d[None, :], qvec) * inv).astype(F)
        ok &= ~((v < 0) | ((u + v).astype(F) > 1))
        t = (dot(e2, qvec) * inv).astype(F)
        ok &= ~(t < 0)
    return ok, t, u, v


def brute_force(slots, o, d):
    ok, t, u, v = moller_trumbore(o, d, slots["v0"], slots["e1"], slots["e2"])
    if not ok.any():
        return -1, F(0), F(0), F(0)
    tt = np.where(ok, t, np.inf)
    best = np.flatnonzero(tt == tt.min())
    k = best[np.argmin(slots["tri"][best])]
    return int(slots["tri"][k]), t[k], u[k], v[k]


def walk_closest_t(nodes, slots, o, d, count=None):
    """One ray through the tree, the kernel's way.  d must be normalised (float32)."""
    o, d = o.astype(F), d.astype(F)
    dd = np.where(np.abs(d) > F(1e-20), d, np.copysign(F(1e-20), d)).astype(F)
    root = nodes[0]
    rs = (root["e"].astype(np.uint32) << 23).view(F)
    far1 = F(F(F(abs(F(o[0] - root["p"][0])) + abs(F(o[1] - root["p"][1]))) + abs(F(o[2] - root["p"][2]))) + F(256.0) * F(F(rs[0] + rs[1]) + rs[2]))
    fe = min(max(int(np.array(far1, F).view(np.uint32)) >> 23, 64), 190)
    tscale = np.array((253 - fe) << 23, np.uint32).view(F)
    inv = ((F(1.0) / dd).astype(F) * tscale).astype(F)
    octinv = (4 if dd[0] >= 0 else 0) | (2 if dd[1] >= 0 else 0) | (1 if dd[2] >= 0 else 0)
    best = (-1, F(0), F(0), F(0))
    tcull = F(np.inf)
    stack = []
    ngroup = [0, 0x80000000]
    while True:
        if ngroup[1] <= 0x00ffffff:
            if not stack:
                break
            ngroup = stack.pop()
        imask = ngroup[1]
        bit = imask.bit_length() - 1
        ngroup[1] &= ~(1 << bit)
        if ngroup[1] > 0x00ffffff:
            stack.append(list(ngroup))
        slot = (bit - 24) ^ octinv
        rank = bin(imask & ~(0xffffffff << slot) & 0xffffffff).count("1")
        N = nodes[ngroup[0] + rank]
        if count is not None:
            count[0] += 1
        scale = (N["e"].astype(np.uint32) << 23).view(F)
        a = (scale * inv).astype(F)
        off = ((N["p"] - o).astype(F) * inv).astype(F)
        offN0 = _fma(F(-32768.0), a, off)
        offN = _fma(-np.abs(offN0), F(1.1920929e-7), offN0)
        qlo = np.stack([N["qlox"], N["qloy"], N["qloz"]], 1)
        qhi = np.stack([N["qhix"], N["qhiy"], N["qhiz"]], 1)
        neg = dd < 0
        qn = np.where(neg[None, :], qhi, qlo).astype(np.uint32)
        qf = np.where(neg[None, :], qlo, qhi).astype(F)
        magic = ((qn << 8) | 0x47000000).astype(np.uint32).view(F)                      # 2^15 + q (byteMagic15)
        tmin = np.clip(_fma(magic, a[None, :], offN[None, :]), F(0), F(1))            # FFMA.SAT
        tmax = _fma(qf, a[None, :], off[None, :])
        cmin = tmin.max(1)
        cmax = np.minimum(tmax.min(1), F(tcull * tscale))
        hit = cmin <= (cmax * F(1.000001)).astype(F)
        h8 = int(sum(1 << i for i in range(8) if hit[i]))
        im8 = int(N["imask"])
        ngroup = [int(N["childBase"]), (int(PERM[octinv][h8 & im8]) << 24) | im8]
        tris = int(EXPAND[h8]) & int(N["triMask"])
        tvalid = int(N["triMask"])
        while tris:
            tb = 1 << (tris.bit_length() - 1)
            tris &= ~tb
            k = int(N["triBase"]) + bin(tvalid & (tb - 1)).count("1")
            S = slots[k:k + 1]
            ok, t, u, v = moller_trumbore(o, d, S["v0"], S["e1"], S["e2"])
            if ok[0]:
                tri = int(S["tri"][0])
                if best[0] < 0 or t[0] < best[1] or (t[0] == best[1] and tri < best[0]):
                    best = (tri, t[0], u[0], v[0])
                    tcull = t[0]
    return best
