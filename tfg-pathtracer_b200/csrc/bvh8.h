/*
 * bvh8.h — compressed wide BVH (8-wide, 80-byte nodes) and its host builder.
 *
 * Replaces the reference's acceleration structure: an implicit complete binary tree of fixed depth 18
 * (524 287 x 44-byte nodes whatever the scene, 57 % empty leaves at ClockCC0 scale) built by a
 * recursive single-threaded CPU builder (S/BVH.hpp:187-330, divideSAH :373-460; S/ = reference
 * src/tfg-pathtracer).  Here: top-down binned-SAH binary build (16 bins, in-place partition, task-parallel),
 * SAH-optimal collapse to 8-wide nodes, octant-ordered child slots, quantised child boxes
 * (layout after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide
 * BVHs", HPG 2017).  Triangles are re-laid out in leaf order as 48-byte records for 128-bit loads.
 */
#ifndef ELEVEN_BVH8_H
#define ELEVEN_BVH8_H

#include <stdint.h>
#include <vector>
#include "../../include/eleven_b200.h"

/* Leaf policy of both builders (binary binned-SAH stage): a node of <= EL_MAX_LEAF (2; the node layout allows up to 3) triangles stays a leaf unless
 * splitting it is cheaper by SAH, with a triangle test costing 1 and the split EL_LEAF_COST_NODE. */
#ifndef EL_MAX_LEAF
#define EL_MAX_LEAF 2
#endif
#ifndef EL_LEAF_COST_NODE
#define EL_LEAF_COST_NODE 1.0f
#endif

namespace eleven {

/* 80 bytes = 5 x 16-byte loads.
 * child box i: lo = p + 2^e * qlo[i], hi = p + 2^e * qhi[i] (per axis); child i sits in octant slot i.
 * imask: bit i set = child i is an internal node (its node index = childBase + popcount(imask below i)).
 * triMask: bit 3*i + k set = child i is a leaf holding more than k triangles (<= 3 per leaf).  The triangles of a node
 *   are contiguous in memory in (slot, k) order: the slot of bit b is triBase + popcount(triMask below b).  The fixed
 *   3-bits-per-child positions let the traversal turn its 8-bit child hit mask into the triangle mask with one table
 *   look-up and one AND (round 1 stored an (offset, unary count) byte per child and spent 5 ALU-pipe instructions per
 *   child on `count << offset`: the ALU pipe was the limiter of the node test).
 * slack: largest TriSlot::shiftBound below this node (the same value as Bvh8::nodeSlack[i], inside the node so that the
 *   KEY-mode cull needs no second load). */
struct alignas(16) Node8 {
    float   px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t childBase, triBase;
    uint32_t triMask;
    float    slack;
    uint8_t qlox[8], qloy[8];
    uint8_t qloz[8], qhix[8];
    uint8_t qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

/* 48 bytes = 3 x float4.  e1 = v1 - v0 and e2 = v2 - v0 are the float32 differences the reference forms per hit
 * test (S/Tri.hpp:42-43), precomputed once: identical bits, no per-test subtraction. */
struct alignas(16) TriSlot {
    float v0x, v0y, v0z, e1x;
    float e1y, e1z, e2x, e2y;
    float e2z; int32_t tri; int32_t material; float shiftBound;   /* bound on |hit.position - geometric position| for THIS triangle */
};
static_assert(sizeof(TriSlot) == 48, "TriSlot must be 48 bytes");

struct Bvh8 {
    std::vector<Node8>   nodes;       /* nodes[0] is the root */
    std::vector<TriSlot> slots;       /* triangles in leaf order */
    std::vector<float>   nodeSlack;   /* per node: max TriSlot::shiftBound in its subtree */
    float  keySlack;                  /* bound on |reference key - t| over the scene (see build) */
    float  boundsLo[3], boundsHi[3];
    double buildMs;
    uint32_t maxDepth;
};

/* One piece of a pre-split triangle (bvh8_build.cpp: presplitTriangles): the box of the part of triangle `tri` inside one cell. */
struct PresplitPiece { float lo[3], hi[3]; uint32_t tri; };
void presplitTriangles(const ElevenTri* tris, uint32_t n, std::vector<PresplitPiece>& out);
/* The two halves separately, for the device builder: it selects the candidates on the GPU (same rule, thresholds from presplitParams) and
 * has the host clip only those (ascending triangle indices). */
void presplitParams(float& areaFactor, float& sliverFactor, int& depth);
void presplitCandidates(const ElevenTri* tris, const uint32_t* cand, size_t nc, float meanArea, std::vector<PresplitPiece>& out);

/* Builds the BVH8 over `tris`.  triMaterial[i] = material id of triangle i (object -> material resolved).
 * presplit: sliver triangles enter the build as several references (their slots are repeated in the leaves). */
void buildBvh8(const ElevenTri* tris, uint32_t n, const int32_t* triMaterial, Bvh8& out, int threads, bool presplit = true);

} // namespace eleven
#endif
