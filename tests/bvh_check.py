"""Host-side validator of the device-resident BVH8 (bvh8.h layout), used on the output of BOTH builders (host:
bvh8_build.cpp, device: bvh8_build_gpu.cuh).  Checks what the traversal kernel relies on:

  * every triangle sits in at least one leaf slot — exactly one unless the builder pre-split it (bvh8_build.cpp:
    presplitTriangles: a sliver triangle enters the build as several references, each with the box of the part of the
    triangle inside one cell; its slot is then repeated in several leaves); slots of a node are contiguous (<= 24) and
    packed in triMask bit order (slot of bit b = triBase + popcount(triMask below b)), internal children contiguous in
    ascending slot order (rank = popcount of imask below the slot);
  * every child box, decoded the way the kernel decodes it (p + q * 2^(e-127) in float32), encloses the boxes of all the
    unsplit triangles below it, and the leaf boxes of a split triangle together COVER it: every one of a few hundred
    points sampled on the triangle (corners, edges, interior) lies in the leaf box of one of its slots — a ray that hits
    the triangle at that point walks into that leaf;
  * per-node slack >= the shift bound of every triangle below.
Returns (sah_cost, depth, node_count): SAH cost = sum over child boxes of area x (1 for an internal child, #triangles for
a leaf child) / root area.
"""
import numpy as np


def _area(lo, hi):
    d = np.maximum(hi - lo, 0)
    return 2.0 * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2])


def validate_bvh8(nodes, slots, slack, tris):
    ntri = len(tris)
    n = len(slots)                                                    # >= ntri: pre-split triangles own several slots
    assert n >= ntri
    uses = np.bincount(slots["tri"], minlength=ntri) if n else np.zeros(ntri, int)
    assert len(uses) == ntri and (uses >= 1).all(), "every triangle must appear in at least one slot"
    split = uses > 1                                                  # per triangle
    v = tris["vertices"][slots["tri"]].astype(np.float32)            # (n, 3, 3) in slot order
    assert (slots["v0"] == v[:, 0]).all() and (slots["e1"] == v[:, 1] - v[:, 0]).all() and (slots["e2"] == v[:, 2] - v[:, 0]).all()
    slo, shi = v.min(1), v.max(1)
    whole = ~split[slots["tri"]]                                      # per slot: the leaf box must enclose the whole triangle
    if n == 0:
        return 0.0, 0, len(nodes)
    seen_nodes = np.zeros(len(nodes), bool)
    seen_slots = np.zeros(n, bool)
    sub_lo = np.full((len(nodes), 3), np.inf, np.float32)
    sub_hi = np.full((len(nodes), 3), -np.inf, np.float32)
    sub_shift = np.zeros(len(nodes), np.float32)
    order, depth_of = [], {0: 1}
    stack = [0]
    cost, root_area = 0.0, None
    boxes = {}
    decoded = {}
    while stack:
        ni = stack.pop()
        assert not seen_nodes[ni], "node %d reached twice" % ni
        seen_nodes[ni] = True
        order.append(ni)
        N = nodes[ni]
        scale = (N["e"].astype(np.uint32) << 23).view(np.float32)
        p = N["p"]
        qlo = np.stack([N["qlox"], N["qloy"], N["qloz"]], 1).astype(np.float32)      # (8, 3)
        qhi = np.stack([N["qhix"], N["qhiy"], N["qhiz"]], 1).astype(np.float32)
        blo = (p[None, :] + qlo * scale[None, :]).astype(np.float32)
        bhi = (p[None, :] + qhi * scale[None, :]).astype(np.float32)
        tri_off_expected, rank = 0, 0
        tri_mask, imask = int(N["triMask"]), int(N["imask"])
        decoded[ni] = (blo, bhi, imask, tri_mask, int(N["childBase"]), int(N["triBase"]))
        assert tri_mask >> 24 == 0
        assert N["slack"] == slack[ni], "Node8::slack must repeat the per-node slack array"
        used = []
        for s in range(8):
            unary = (tri_mask >> (3 * s)) & 7
            if (imask >> s) & 1:                                        # internal child
                assert unary == 0, "a child is internal or a leaf, not both"
                used.append(s)
                ci = int(N["childBase"]) + rank
                rank += 1
                assert 0 < ci < len(nodes)
                boxes[ci] = (ni, blo[s], bhi[s])
                depth_of[ci] = depth_of[ni] + 1
                stack.append(ci)
                cost += _area(blo[s], bhi[s])
            elif unary:
                used.append(s)
                cnt = {1: 1, 3: 2, 7: 3}[unary]
                off = bin(tri_mask & ((1 << (3 * s)) - 1)).count("1")    # the kernel's rank: popcount of triMask below the bit
                assert off == tri_off_expected
                tri_off_expected += cnt
                a = int(N["triBase"]) + off
                assert a + cnt <= n and not seen_slots[a:a + cnt].any()
                seen_slots[a:a + cnt] = True
                w = whole[a:a + cnt]
                assert (blo[s] <= slo[a:a + cnt][w]).all() and (bhi[s] >= shi[a:a + cnt][w]).all(), "leaf box does not enclose its triangles (node %d slot %d)" % (ni, s)
                if w.any():                                             # whole triangles must be enclosed all the way up; pieces are checked by point location below
                    sub_lo[ni] = np.minimum(sub_lo[ni], slo[a:a + cnt][w].min(0)); sub_hi[ni] = np.maximum(sub_hi[ni], shi[a:a + cnt][w].max(0))
                sub_shift[ni] = max(sub_shift[ni], slots["shiftBound"][a:a + cnt].max())
                cost += _area(blo[s], bhi[s]) * cnt
        assert tri_off_expected <= 24
        if ni == 0:
            root_area = _area(blo[used].min(0), bhi[used].max(0))
    assert seen_nodes.all(), "unreachable nodes"
    assert seen_slots.all(), "triangle slots not referenced by any leaf"
    for ni in reversed(order):                                        # children after parents in `order`: fold bottom-up
        if ni == 0:
            continue
        parent, lo, hi = boxes[ni]
        assert (lo <= sub_lo[ni]).all() and (hi >= sub_hi[ni]).all(), "internal child box does not enclose its subtree (node %d)" % ni
        sub_lo[parent] = np.minimum(sub_lo[parent], sub_lo[ni]); sub_hi[parent] = np.maximum(sub_hi[parent], sub_hi[ni])
        sub_shift[parent] = max(sub_shift[parent], sub_shift[ni])
    assert (slack >= sub_shift).all(), "per-node slack below a triangle's shift bound"
    # coverage of the pre-split triangles: POINT LOCATION.  For points sampled on the triangle (corners, edges, interior) there must be
    # a root-to-leaf chain of child boxes that all contain the point and ends in a leaf listing the triangle: a ray hitting the
    # triangle there intersects every box of that chain, so the walk reaches the slot.
    if split.any():
        rng = np.random.RandomState(7)
        lin = np.linspace(0, 1, 12)
        bary = np.concatenate([np.eye(3), rng.dirichlet((1, 1, 1), 40), np.stack([lin, 1 - lin, 0 * lin], 1), np.stack([0 * lin, lin, 1 - lin], 1), np.stack([1 - lin, 0 * lin, lin], 1)])

        def locate(P, t):
            todo = [0]
            while todo:
                ni = todo.pop()
                blo_, bhi_, imask_, tmask_, cbase, tbase = decoded[ni]
                inside = ((P >= blo_.astype(np.float64) - 1e-12) & (P <= bhi_.astype(np.float64) + 1e-12)).all(1)
                rank_ = 0
                for s_ in range(8):
                    if (imask_ >> s_) & 1:
                        if inside[s_]:
                            todo.append(cbase + rank_)
                        rank_ += 1
                    elif inside[s_]:
                        unary = (tmask_ >> (3 * s_)) & 7
                        if unary:
                            cnt_ = {1: 1, 3: 2, 7: 3}[unary]
                            a_ = tbase + bin(tmask_ & ((1 << (3 * s_)) - 1)).count("1")
                            if (slots["tri"][a_:a_ + cnt_] == t).any():
                                return True
            return False

        for t in np.flatnonzero(split):
            P = (bary[:, :, None] * tris["vertices"][t].astype(np.float64)[None]).sum(1)
            missing = [k for k in range(len(P)) if not locate(P[k], t)]
            assert not missing, "pre-split triangle %d: %d of %d sampled points are in no leaf that lists it" % (t, len(missing), len(P))
    return cost / max(root_area, 1e-30), max(depth_of.values()), len(nodes)
