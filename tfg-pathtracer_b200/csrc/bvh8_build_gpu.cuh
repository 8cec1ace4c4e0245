/*
 * bvh8_build_gpu.cuh — the BVH8 builder on the GPU.
 *
 * Replaces S/BVH.hpp:187-330 + divideSAH :373-460 (recursive, single-threaded host build: 1.4 s for ClockCC0, >60 s
 * estimated for 10 M triangles) and complements our own host builder (bvh8_build.cpp).  Same algorithm family as the
 * host builder — top-down binned SAH (16 bins x 3 axes), SAH-terminated leaves of <= EL_MAX_LEAF (2) triangles, collapse to
 * 8-wide nodes (SAH-optimal cut by dynamic programming), octant-ordered slots, outward-rounded 8-bit quantisation — restructured
 * for the device:
 *
 *   prep        one thread per triangle: box, shadow-terminator shift bound, scene bounds (ordered-uint atomics)
 *   per LEVEL of the binary tree (all nodes of a level at once, one host read-back of the active-node count):
 *     centroid  centroid bounds of every node being split      (warp REDUX + atomics)
 *     bin       16 x 3 bins per node: box, count, max shift     (shared-memory bins when a block lies inside one node,
 *                                                                 global atomics otherwise)
 *     split     one thread per node: SAH sweep, children allocated and boxed from the bin unions
 *     partition out-of-place scatter of the triangle indices with warp-aggregated cursors
 *   collapse    bottom-up per level of the binary tree: the SAH cost tables of the optimal 8-wide cut (k_collapseDp); then per
 *               LEVEL of the wide tree, one thread per BVH8 node: follow the recorded cut, assign octant slots, quantise, emit
 *               the 80-byte node + its 48-byte triangle slots at prefix sums; children of a node contiguous
 *
 * The traversal result does not depend on the tree (closest hit by the reference key, ties to the lowest triangle
 * index), so the parity contract holds for any valid tree; tests check validity (every triangle in exactly one leaf, every
 * child box encloses its triangles after the kernel's float decode) and that its SAH cost is the host builder's.
 * The emitted tree is deterministic: bins are min/max/count (order-independent), wide nodes and triangle slots are placed
 * at prefix sums (no atomics in the collapse), and the <= 3 triangles of a leaf — whose order in the index array depends
 * on the arrival order of the partition atomics — are emitted in triangle-id order.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>      // exclusive prefix sums of the per-node child/triangle counts (plumbing, not a hot op)

#include "bvh8.h"
#include "common.cuh"

namespace eleven {
namespace gpubvh {

enum { BINS = 16, MAX_LEAF = EL_MAX_LEAF, BIN_WORDS = 8, NODE_BIN_WORDS = 3 * BINS * BIN_WORDS };   // bin: lo xyz, hi xyz, count, max shift
enum { C_NODES = 0, C_NEXT_ACTIVE = 1, C_COUNT = 8 };
enum { SCENE_LO = 0, SCENE_HI = 3, SCENE_SHIFT = 6, SCENE_WORDS = 8 };

// order-preserving float <-> uint mapping, so that float min/max are native unsigned atomics / REDUX
__host__ __device__ __forceinline__ uint32_t encF(float f) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(f);
#else
    memcpy(&b, &f, 4);
#endif
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__host__ __device__ __forceinline__ float decF(uint32_t e) {
    const uint32_t b = e ^ ((e >> 31) ? 0x80000000u : 0xffffffffu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
#define EL_ENC_POS_INF 0xff800000u   /* encF(+inf) */
#define EL_ENC_NEG_INF 0x007fffffu   /* encF(-inf) */

struct Node2G {                      // binary node, 96 bytes
    float lo[3], hi[3];              // box of the (padded) triangle boxes below
    uint32_t clo[3], chi[3];         // centroid bounds, encF (filled by k_centroid at the level the node is split)
    int32_t left;                    // first child (right = left + 1); -1 = leaf
    uint32_t first, count;           // range in the index array
    int32_t binSlot;                 // >= 0 while the node is being split at the current level
    int32_t axis, splitBin;          // chosen split; axis -1 = median split by position
    uint32_t leftCount, curL, curR;  // partition cursors
    float maxShift;                  // max shadow-terminator shift bound below (nodeSlack of the wide node)
    uint32_t pad_[2];
};
static_assert(sizeof(Node2G) == 96, "Node2G layout");

__device__ __forceinline__ int binOf(float c, float cmin, float ext) {
    const float scale = (float)BINS / ext;
    return min(BINS - 1, max(0, (int)((c - cmin) * scale)));
}
__device__ __forceinline__ float boxArea(const float* lo, const float* hi) {
    const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    return (x < 0.f || y < 0.f || z < 0.f) ? 0.f : 2.f * (x * y + x * z + y * z);
}

// ---- prep ------------------------------------------------------------------------------------------------------------------
// boxLo[i] = (lo.xyz, shift bound), boxHi[i] = (hi.xyz, -): UNPADDED boxes; the pad is added where boxes are consumed
__global__ void __launch_bounds__(256) k_prep(const ElevenTri* __restrict__ tris, uint32_t n, float4* __restrict__ boxLo, float4* __restrict__ boxHi,
                                              uint32_t* __restrict__ idx, uint32_t* __restrict__ owner, uint32_t* __restrict__ scene) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float shift = 0.f;
    if (i < n) {
        const ElevenTri& T = tris[i];
        float v[3][3], nn[3][3];
        for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) { v[k][a] = T.vertices[k][a]; nn[k][a] = T.normals[k][a]; }
        for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], v[k][a]); hi[a] = fmaxf(hi[a], v[k][a]); }
        // bound on |shadingPosition - geometricPosition| (S/Tri.hpp:81-89), see bvh8_build.cpp: max_{i,j} |(v_j - v_i).n_i| |n_i|
        double mine = 0.0;
        for (int vi = 0; vi < 3; vi++) {
            const double nl = sqrt((double)nn[vi][0] * nn[vi][0] + (double)nn[vi][1] * nn[vi][1] + (double)nn[vi][2] * nn[vi][2]);
            for (int vj = 0; vj < 3; vj++) {
                if (vj == vi) continue;
                double d = 0.0;
                for (int a = 0; a < 3; a++) d += ((double)v[vj][a] - (double)v[vi][a]) * (double)nn[vi][a];
                mine = fmax(mine, fabs(d) * nl);
            }
        }
        shift = (float)(mine * 1.0001);
        if (!(shift >= 0.f)) shift = INFINITY;                      // NaN normals: no culling slack can be trusted
        boxLo[i] = make_float4(lo[0], lo[1], lo[2], shift);
        boxHi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
        idx[i] = i; owner[i] = 0u;
    }
    // scene bounds: warp REDUX then one atomic per warp and value
    const uint32_t FULL = 0xffffffffu;
    for (int a = 0; a < 3; a++) {
        const uint32_t l = __reduce_min_sync(FULL, encF(lo[a])), h = __reduce_max_sync(FULL, encF(hi[a]));
        if ((threadIdx.x & 31u) == 0u) { atomicMin(&scene[SCENE_LO + a], l); atomicMax(&scene[SCENE_HI + a], h); }
    }
    const uint32_t s = __reduce_max_sync(FULL, __float_as_uint(shift));     // shift >= 0: bit order == value order
    if ((threadIdx.x & 31u) == 0u) atomicMax(&scene[SCENE_SHIFT], s);
}

// ---- pre-split candidates on the device (bvh8_build.cpp: presplitTriangles states the rule) ---------------------------------------
// Box area per triangle, summed per block in a FIXED order (shared-memory tree) so that the mean — and with it the tree — is the same
// on every run and every GPU; the host adds the block sums in block order.
__device__ __forceinline__ float triBoxArea(const ElevenTri& T) {
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = fminf(fminf(T.vertices[0][a], T.vertices[1][a]), T.vertices[2][a]); hi[a] = fmaxf(fmaxf(T.vertices[0][a], T.vertices[1][a]), T.vertices[2][a]); }
    return boxArea(lo, hi);
}
__global__ void __launch_bounds__(256) k_presplitAreaSums(const ElevenTri* __restrict__ tris, uint32_t n, double* __restrict__ blockSum) {
    __shared__ double sh[256];
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    sh[threadIdx.x] = i < n ? (double)triBoxArea(tris[i]) : 0.0;
    __syncthreads();
    for (uint32_t w = 128u; w > 0u; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) blockSum[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_presplitSelect(const ElevenTri* __restrict__ tris, uint32_t n, float mean, float areaFactor, float sliverFactor,
                                                        uint32_t* __restrict__ cand, uint32_t cap, uint32_t* __restrict__ count) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const ElevenTri& T = tris[i];
    const float area = triBoxArea(T);
    if (!(area > areaFactor * mean)) return;
    double e1[3], e2[3];
    for (int a = 0; a < 3; a++) { e1[a] = (double)T.vertices[1][a] - (double)T.vertices[0][a]; e2[a] = (double)T.vertices[2][a] - (double)T.vertices[0][a]; }
    const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
    const double twiceTri = sqrt(cx * cx + cy * cy + cz * cz);
    if (!((double)area > (double)sliverFactor * twiceTri)) return;
    const uint32_t k = atomicAdd(count, 1u);
    if (k < cap) cand[k] = i;
}
// Selects the sliver triangles of d_tris on the GPU and has the host clip them (h_tris: the same triangles in host memory) into pieces.
// A scene without slivers costs two small kernels and one synchronisation.
static bool devicePresplit(const ElevenTri* d_tris, const ElevenTri* h_tris, uint32_t n, cudaStream_t st, std::vector<PresplitPiece>& pieces, std::string& err) {
    pieces.clear();
    float areaFactor, sliverFactor; int depth;
    presplitParams(areaFactor, sliverFactor, depth);
    if (n == 0 || depth <= 0) return true;
    const uint32_t blocks = (n + 255u) / 256u, cap = n / 8u + 4096u;
    double* dSum = nullptr; uint32_t* dCand = nullptr;
    if (cudaMallocAsync((void**)&dSum, (size_t)blocks * 8, st) != cudaSuccess || cudaMallocAsync((void**)&dCand, ((size_t)cap + 1) * 4, st) != cudaSuccess) { err = "device BVH build: out of memory (pre-split scratch)"; return false; }
    uint32_t* dCount = dCand + cap;
    std::vector<double> sums(blocks);
    k_presplitAreaSums<<<blocks, 256, 0, st>>>(d_tris, n, dSum);
    bool ok = cudaMemcpyAsync(sums.data(), dSum, (size_t)blocks * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaMemsetAsync(dCount, 0, 4, st) == cudaSuccess &&
              cudaStreamSynchronize(st) == cudaSuccess;
    double total = 0; for (double v : sums) total += v;
    const float mean = (float)(total / n);
    uint32_t count = 0;
    if (ok) {
        k_presplitSelect<<<blocks, 256, 0, st>>>(d_tris, n, mean, areaFactor, sliverFactor, dCand, cap, dCount);
        ok = cudaMemcpyAsync(&count, dCount, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    }
    std::vector<uint32_t> cand;
    if (ok && count > 0 && count <= cap) {                           // more candidates than the buffer holds: a scene OF slivers; build it unsplit
        cand.resize(count);
        ok = cudaMemcpyAsync(cand.data(), dCand, (size_t)count * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    }
    cudaFreeAsync(dSum, st); cudaFreeAsync(dCand, st);
    if (!ok) { err = std::string("device BVH build (pre-split): ") + cudaGetErrorString(cudaGetLastError()); return false; }
    std::sort(cand.begin(), cand.end());                              // the atomics' arrival order is not deterministic
    presplitCandidates(h_tris, cand.data(), cand.size(), mean, pieces);
    return true;
}

// Pre-split triangles (bvh8_build.cpp: presplitTriangles): piece k becomes reference refOf[k] — the triangle's own slot for its first
// piece, slot nTris + j for the others — with the piece's box and the triangle's shift bound.
__global__ void k_applyPieces(const PresplitPiece* __restrict__ pieces, const uint32_t* __restrict__ refOf, uint32_t nPieces,
                              float4* __restrict__ boxLo, float4* __restrict__ boxHi, uint32_t* __restrict__ idx, uint32_t* __restrict__ owner,
                              uint32_t nTris, uint32_t* __restrict__ extraTri) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nPieces) return;
    const PresplitPiece P = pieces[k];
    const uint32_t r = refOf[k];
    const float shift = boxLo[P.tri].w;                              // .w is never rewritten: no race with the first piece's thread
    if (r == P.tri) { float4 lo = make_float4(P.lo[0], P.lo[1], P.lo[2], shift); boxLo[r] = lo; }
    else { boxLo[r] = make_float4(P.lo[0], P.lo[1], P.lo[2], shift); extraTri[r - nTris] = P.tri; }
    boxHi[r] = make_float4(P.hi[0], P.hi[1], P.hi[2], 0.f);
    idx[r] = r; owner[r] = 0u;
}

// ---- per level ---------------------------------------------------------------------------------------------------------------
__global__ void k_initBins(uint32_t* __restrict__ bins, size_t words) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= words) return;
    const uint32_t w = (uint32_t)(i % BIN_WORDS);
    bins[i] = w < 3 ? EL_ENC_POS_INF : (w < 6 ? EL_ENC_NEG_INF : 0u);
}

__global__ void __launch_bounds__(256) k_centroid(const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, const uint32_t* __restrict__ idx,
                                                  const uint32_t* __restrict__ owner, Node2G* __restrict__ nodes, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t FULL = 0xffffffffu;
    uint32_t node = 0xffffffffu - (threadIdx.x & 31u);              // distinct dummy ids for lanes without work
    bool work = false;
    float c[3] = {0.f, 0.f, 0.f};
    if (i < n) {
        const uint32_t o = owner[i];
        if (nodes[o].binSlot >= 0) {
            node = o; work = true;
            const uint32_t t = idx[i];
            const float4 l = boxLo[t], h = boxHi[t];
            c[0] = 0.5f * (l.x + h.x); c[1] = 0.5f * (l.y + h.y); c[2] = 0.5f * (l.z + h.z);
        }
    }
    int same = 0;
    __match_all_sync(FULL, node, &same);
    if (same) {                                                      // the whole warp feeds one node: 6 REDUX + 6 atomics per warp
        for (int a = 0; a < 3; a++) {
            const uint32_t l = __reduce_min_sync(FULL, encF(c[a])), h = __reduce_max_sync(FULL, encF(c[a]));
            if ((threadIdx.x & 31u) == 0u) { atomicMin(&nodes[node].clo[a], l); atomicMax(&nodes[node].chi[a], h); }
        }
    } else if (work) {
        for (int a = 0; a < 3; a++) { atomicMin(&nodes[node].clo[a], encF(c[a])); atomicMax(&nodes[node].chi[a], encF(c[a])); }
    }
}

__device__ __forceinline__ void binUpdate(uint32_t* b, const float* lo, const float* hi, float shift, bool shared) {
    // (shared and global atomics are the same intrinsics; the flag only documents the call sites)
    (void)shared;
    atomicMin(&b[0], encF(lo[0])); atomicMin(&b[1], encF(lo[1])); atomicMin(&b[2], encF(lo[2]));
    atomicMax(&b[3], encF(hi[0])); atomicMax(&b[4], encF(hi[1])); atomicMax(&b[5], encF(hi[2]));
    atomicAdd(&b[6], 1u);
    atomicMax(&b[7], __float_as_uint(shift));
}

__global__ void __launch_bounds__(256) k_bin(const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, const uint32_t* __restrict__ idx,
                                             const uint32_t* __restrict__ owner, const Node2G* __restrict__ nodes, uint32_t* __restrict__ bins,
                                             uint32_t n, float pad, int slotBase, int slotCount) {
    __shared__ uint32_t sb[NODE_BIN_WORDS];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b0 = blockIdx.x * blockDim.x, b1 = min(n, b0 + blockDim.x) - 1u;
    const uint32_t nodeA = owner[b0], nodeB = owner[b1];
    const bool priv = nodeA == nodeB;                                // the block lies inside one node: bins in shared memory
    // the active nodes of a level are binned in chunks of `slotCount` nodes (bounded scratch): this launch serves the nodes
    // whose slot lies in [slotBase, slotBase + slotCount)
    if (priv) {
        const int sA = nodes[nodeA].binSlot - slotBase;
        if (nodes[nodeA].binSlot < 0 || sA < 0 || sA >= slotCount) return;   // ... which is not being split in this chunk: nothing to do for the block
        for (uint32_t w = threadIdx.x; w < NODE_BIN_WORDS; w += blockDim.x) { const uint32_t k = w % BIN_WORDS; sb[w] = k < 3 ? EL_ENC_POS_INF : (k < 6 ? EL_ENC_NEG_INF : 0u); }
        __syncthreads();
    }
    if (i < n) {
        const uint32_t o = priv ? nodeA : owner[i];
        const Node2G& N = nodes[o];
        const int slot = N.binSlot - slotBase;
        if (N.binSlot >= 0 && slot >= 0 && slot < slotCount) {
            const uint32_t t = idx[i];
            const float4 l = boxLo[t], h = boxHi[t];
            const float lo[3] = {l.x - pad, l.y - pad, l.z - pad}, hi[3] = {h.x + pad, h.y + pad, h.z + pad};
            const float c[3] = {0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z)};
            for (int a = 0; a < 3; a++) {
                const float cmin = decF(N.clo[a]), ext = decF(N.chi[a]) - cmin;
                if (!(ext > 0.f)) continue;
                const int k = binOf(c[a], cmin, ext);
                uint32_t* b = (priv ? sb : bins + (size_t)slot * NODE_BIN_WORDS) + (a * BINS + k) * BIN_WORDS;
                binUpdate(b, lo, hi, l.w, priv);
            }
        }
    }
    if (priv) {
        __syncthreads();
        uint32_t* g = bins + (size_t)(nodes[nodeA].binSlot - slotBase) * NODE_BIN_WORDS;
        for (uint32_t w = threadIdx.x; w < NODE_BIN_WORDS; w += blockDim.x) {
            const uint32_t k = w % BIN_WORDS, v = sb[w];
            if (k < 3) { if (v != EL_ENC_POS_INF) atomicMin(&g[w], v); }
            else if (k < 6) { if (v != EL_ENC_NEG_INF) atomicMax(&g[w], v); }
            else if (k == 6) { if (v) atomicAdd(&g[w], v); }
            else { if (v) atomicMax(&g[w], v); }
        }
    }
}

__global__ void __launch_bounds__(128) k_split(Node2G* __restrict__ nodes, const uint32_t* __restrict__ bins, const uint32_t* __restrict__ active,
                                               uint32_t activeCount, uint32_t* __restrict__ nextActive, uint32_t* __restrict__ counters, int forceMedian,
                                               int slotBase) {
    const uint32_t ai = blockIdx.x * blockDim.x + threadIdx.x;       // index inside the chunk; `active` points at the chunk
    if (ai >= activeCount) return;
    const uint32_t ni = active[ai];
    Node2G N = nodes[ni];
    const uint32_t* B = bins + (size_t)(N.binSlot - slotBase) * NODE_BIN_WORDS;
    float bestCost = INFINITY; int bestAxis = -1, bestSplit = -1;
    if (!forceMedian) {
        for (int a = 0; a < 3; a++) {
            const float ext = decF(N.chi[a]) - decF(N.clo[a]);
            if (!(ext > 0.f)) continue;
            const uint32_t* Ba = B + a * BINS * BIN_WORDS;
            float rightArea[BINS]; uint32_t rightCount[BINS];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}; uint32_t c = 0;
            for (int k = BINS - 1; k > 0; k--) {
                const uint32_t* b = Ba + k * BIN_WORDS;
                for (int d = 0; d < 3; d++) { lo[d] = fminf(lo[d], decF(b[d])); hi[d] = fmaxf(hi[d], decF(b[3 + d])); }
                c += b[6]; rightArea[k] = boxArea(lo, hi); rightCount[k] = c;
            }
            for (int d = 0; d < 3; d++) { lo[d] = INFINITY; hi[d] = -INFINITY; }
            c = 0;
            for (int k = 1; k < BINS; k++) {
                const uint32_t* b = Ba + (k - 1) * BIN_WORDS;
                for (int d = 0; d < 3; d++) { lo[d] = fminf(lo[d], decF(b[d])); hi[d] = fmaxf(hi[d], decF(b[3 + d])); }
                c += b[6];
                if (c == 0u || rightCount[k] == 0u) continue;
                const float cost = boxArea(lo, hi) * (float)c + rightArea[k] * (float)rightCount[k];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = k; }
            }
        }
    }
    // SAH leaf termination, as the host builder: a node of <= MAX_LEAF triangles stays a leaf unless splitting it is cheaper
    if (N.count <= (uint32_t)MAX_LEAF) {
        const float area = boxArea(N.lo, N.hi);
        if (bestAxis < 0 || area * (float)N.count <= area * EL_LEAF_COST_NODE + bestCost) return;      // cost of a triangle test = 1, of the split = EL_LEAF_COST_NODE (bvh8.h)
    }
    const int left = (int)atomicAdd(&counters[C_NODES], 2u);
    Node2G L, R;
    memset(&L, 0, sizeof L); memset(&R, 0, sizeof R);
    uint32_t lc;
    if (bestAxis >= 0) {
        const uint32_t* Ba = B + bestAxis * BINS * BIN_WORDS;
        float llo[3] = {INFINITY, INFINITY, INFINITY}, lhi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float rlo[3] = {INFINITY, INFINITY, INFINITY}, rhi[3] = {-INFINITY, -INFINITY, -INFINITY};
        uint32_t ls = 0, rs = 0; lc = 0;
        for (int k = 0; k < BINS; k++) {
            const uint32_t* b = Ba + k * BIN_WORDS;
            if (k < bestSplit) { for (int d = 0; d < 3; d++) { llo[d] = fminf(llo[d], decF(b[d])); lhi[d] = fmaxf(lhi[d], decF(b[3 + d])); } lc += b[6]; ls = max(ls, b[7]); }
            else { for (int d = 0; d < 3; d++) { rlo[d] = fminf(rlo[d], decF(b[d])); rhi[d] = fmaxf(rhi[d], decF(b[3 + d])); } rs = max(rs, b[7]); }
        }
        for (int d = 0; d < 3; d++) { L.lo[d] = llo[d]; L.hi[d] = lhi[d]; R.lo[d] = rlo[d]; R.hi[d] = rhi[d]; }
        L.maxShift = __uint_as_float(ls); R.maxShift = __uint_as_float(rs);
    } else {                                                         // all centroids coincide (or depth guard): split by position, children keep the parent box
        lc = N.count / 2u;
        for (int d = 0; d < 3; d++) { L.lo[d] = R.lo[d] = N.lo[d]; L.hi[d] = R.hi[d] = N.hi[d]; }
        L.maxShift = R.maxShift = N.maxShift;
    }
    L.first = N.first; L.count = lc; R.first = N.first + lc; R.count = N.count - lc;
    Node2G* ch[2] = {&L, &R};
    for (int s = 0; s < 2; s++) {
        Node2G& C = *ch[s];
        C.left = -1; C.axis = -1; C.splitBin = 0; C.leftCount = 0; C.curL = C.curR = 0;
        for (int d = 0; d < 3; d++) { C.clo[d] = EL_ENC_POS_INF; C.chi[d] = EL_ENC_NEG_INF; }
        C.binSlot = -1;
        if (C.count > 1u) {                                          // 2..MAX_LEAF triangles: k_split decides by SAH whether to split further
            const uint32_t slot = atomicAdd(&counters[C_NEXT_ACTIVE], 1u);
            nextActive[slot] = (uint32_t)(left + s);
            C.binSlot = (int)slot;
        }
        nodes[left + s] = C;
    }
    N.left = left; N.axis = bestAxis; N.splitBin = bestSplit; N.leftCount = lc; N.curL = 0; N.curR = 0;
    nodes[ni] = N;
}

__global__ void __launch_bounds__(256) k_partition(const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, const uint32_t* __restrict__ idxIn,
                                                   const uint32_t* __restrict__ ownerIn, uint32_t* __restrict__ idxOut, uint32_t* __restrict__ ownerOut,
                                                   Node2G* __restrict__ nodes, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const bool valid = i < n;
    uint32_t o = 0, t = 0, key = 0xffffffffu - lane;               // key: (child id) for scattered entries, unique otherwise
    bool scatter = false, goLeft = false;
    Node2G* N = nullptr;
    if (valid) {
        o = ownerIn[i]; t = idxIn[i];
        N = nodes + o;
        if (N->left >= 0 && N->binSlot >= 0) {                       // split at this level
            if (N->axis >= 0) {
                const float4 l = boxLo[t], h = boxHi[t];
                const float c = N->axis == 0 ? 0.5f * (l.x + h.x) : (N->axis == 1 ? 0.5f * (l.y + h.y) : 0.5f * (l.z + h.z));
                const float cmin = decF(N->clo[N->axis]), ext = decF(N->chi[N->axis]) - cmin;
                goLeft = binOf(c, cmin, ext) < N->splitBin;
                scatter = true;
                key = (uint32_t)N->left + (goLeft ? 0u : 1u);
            } else {                                                 // median split: positions stay, owners change
                goLeft = (i - N->first) < N->leftCount;
                idxOut[i] = t; ownerOut[i] = (uint32_t)N->left + (goLeft ? 0u : 1u);
            }
        } else { idxOut[i] = t; ownerOut[i] = o; }
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (scatter) {
        const uint32_t leader = __ffs(peers) - 1u;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(goLeft ? &N->curL : &N->curR, (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const uint32_t pos = N->first + (goLeft ? 0u : N->leftCount) + base + __popc(peers & ((1u << lane) - 1u));
        idxOut[pos] = t; ownerOut[pos] = key;
    }
}

// nodes split at this level stop being "active": their binSlot must not alias a slot of the next level
__global__ void k_retire(Node2G* __restrict__ nodes, const uint32_t* __restrict__ active, uint32_t activeCount) {
    const uint32_t ai = blockIdx.x * blockDim.x + threadIdx.x;
    if (ai < activeCount) nodes[active[ai]].binSlot = -1;
}

// ---- collapse to 8-wide + emission ----------------------------------------------------------------------------------------------
struct Item8 { uint32_t n2, n8, depth; };

__device__ __forceinline__ uint8_t quantExpDev(float extent) {       // smallest power-of-two cell (biased exponent byte) with 2^e * 255 >= extent
    if (!(extent > 0.f)) return 0;
    int e = (int)ceil(log2((double)extent / 255.0));
    while (ldexp(255.0, e) < (double)extent) e++;
    e = max(-126, min(127, e));
    return (uint8_t)(e + 127);
}

// SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1), bottom-up over the levels of the binary tree (the nodes of a
// level have contiguous ids: k_split allocates them level by level).  C(n, i) = lowest SAH cost of the subtree of binary node n
// when it may occupy i child slots of its wide parent:
//   leaf:      C(n, i) = A_n * count
//   internal:  D(n, j) = min_{0<k<j} C(l, k) + C(r, j-k);  C(n, 1) = D(n, 8) + A_n * c_node;  C(n, i) = min(D(n, i), C(n, i-1))
// dpK[n][j] (j >= 2) = the k of D(n, j), or 0 = "C(n, j-1) is as good"; dpK[n][1] = the k of D(n, 8), i.e. how n's own 8 slots
// split between its two subtrees.  Measured against the greedy open-the-largest-child collapse: 36 % fewer wide nodes, 6 % fewer
// node visits per ray, k_extend -5 %.
#define EL_COLLAPSE_COST_NODE 3.0f   /* a wide-node visit costs ~3 triangle tests (~300 vs ~100 instructions); the trees for 1..5 are the same */
__global__ void __launch_bounds__(128) k_collapseDp(const Node2G* __restrict__ nodes, uint32_t first, uint32_t count, float* __restrict__ dpC, uint8_t* __restrict__ dpK) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t i = first + t;
    const Node2G& N = nodes[i];
    const float area = boxArea(N.lo, N.hi);
    float* c = dpC + (size_t)i * 8; uint8_t* kk = dpK + (size_t)i * 8;
    if (N.left < 0) {
        const float v = area * (float)N.count;
        for (int j = 0; j < 8; j++) { c[j] = v; kk[j] = 0; }
        return;
    }
    float L[8], R[8];
    for (int j = 1; j < 8; j++) { L[j] = dpC[(size_t)N.left * 8 + j]; R[j] = dpC[(size_t)(N.left + 1) * 8 + j]; }
    float D[9]; uint8_t K[9];
    for (int j = 2; j <= 8; j++) {
        D[j] = INFINITY; K[j] = 1;
        for (int k = 1; k < j; k++) { const float v = L[min(k, 7)] + R[min(j - k, 7)]; if (v < D[j]) { D[j] = v; K[j] = (uint8_t)k; } }
    }
    float prev = D[8] + area * EL_COLLAPSE_COST_NODE;
    c[0] = INFINITY; kk[0] = 0; c[1] = prev; kk[1] = K[8];
    for (int j = 2; j <= 7; j++) {
        if (D[j] < prev) { prev = D[j]; kk[j] = K[j]; } else kk[j] = 0;
        c[j] = prev;
    }
}

// Pass 1 of a wide-tree level: choose the (up to 8) children of every wide node and their octant slots; count its internal
// children and triangles.  Pass 2 (k_collapseEmit) places nodes, triangle slots and next-level items at the exclusive
// prefix sums of those counts, so the emitted tree does not depend on any atomic arrival order.
__global__ void __launch_bounds__(128) k_collapseGather(const Node2G* __restrict__ nodes, const uint8_t* __restrict__ dpK, const Item8* __restrict__ items,
                                                        uint32_t itemCount, int32_t* __restrict__ childAtOut, uint32_t* __restrict__ nInt, uint32_t* __restrict__ nTri) {
    const uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= itemCount) return;
    const Item8 it = items[ii];
    const Node2G& root = nodes[it.n2];
    int ch[8]; int nc = 0;
    if (root.left < 0) { if (root.count > 0u) ch[nc++] = (int)it.n2; }
    else {
        // the SAH-optimal cut below this node (k_collapseDp): walk the recorded slot splits, budgets sum to 8
        int sm[8], sj[8], sp = 0;
        const int k0 = dpK[(size_t)it.n2 * 8 + 1];
        sm[sp] = root.left + 1; sj[sp++] = 8 - k0;
        sm[sp] = root.left; sj[sp++] = k0;
        while (sp > 0) {
            const int m = sm[--sp]; int j = sj[sp];
            const int ml = nodes[m].left;
            while (ml >= 0 && j > 1 && dpK[(size_t)m * 8 + j] == 0) j--;
            if (ml < 0 || j <= 1) { ch[nc++] = m; continue; }
            const int k = dpK[(size_t)m * 8 + j];
            sm[sp] = ml + 1; sj[sp++] = j - k;
            sm[sp] = ml; sj[sp++] = k;
        }
    }
    // octant-ordered slots (greedy assignment, as the host builder)
    const float cx[3] = {0.5f * (root.lo[0] + root.hi[0]), 0.5f * (root.lo[1] + root.hi[1]), 0.5f * (root.lo[2] + root.hi[2])};
    float d[8][3]; int slotOf[8]; uint32_t cUsed = 0, sUsed = 0;
    for (int c = 0; c < nc; c++) { const Node2G& cb = nodes[ch[c]]; for (int a = 0; a < 3; a++) d[c][a] = 0.5f * (cb.lo[a] + cb.hi[a]) - cx[a]; }
    for (int k = 0; k < nc; k++) {
        int bc = -1, bs = -1; float bv = -INFINITY;
        for (int c = 0; c < nc; c++) {
            if (cUsed & (1u << c)) continue;
            for (int s = 0; s < 8; s++) {
                if (sUsed & (1u << s)) continue;
                const float cost = ((s & 4) ? d[c][0] : -d[c][0]) + ((s & 2) ? d[c][1] : -d[c][1]) + ((s & 1) ? d[c][2] : -d[c][2]);
                if (cost > bv || bc < 0) { bv = cost; bc = c; bs = s; }
            }
        }
        cUsed |= 1u << bc; sUsed |= 1u << bs; slotOf[bc] = bs;
    }
    int childAt[8];
    for (int s = 0; s < 8; s++) childAt[s] = -1;
    for (int c = 0; c < nc; c++) childAt[slotOf[c]] = ch[c];
    uint32_t nInternal = 0, nTris = 0;
    for (int s = 0; s < 8; s++) {
        childAtOut[(size_t)ii * 8 + s] = childAt[s];
        if (childAt[s] < 0) continue;
        const Node2G& c = nodes[childAt[s]];
        if (c.left >= 0) nInternal++; else nTris += c.count;
    }
    nInt[ii] = nInternal; nTri[ii] = nTris;
}

__global__ void __launch_bounds__(128) k_collapseEmit(const Node2G* __restrict__ nodes, const uint32_t* __restrict__ idx, const ElevenTri* __restrict__ tris,
                                                      const int32_t* __restrict__ triMaterial, const float4* __restrict__ boxLo,
                                                      uint32_t nTris, const uint32_t* __restrict__ extraTri,
                                                      const Item8* __restrict__ items, uint32_t itemCount, const int32_t* __restrict__ childAtIn,
                                                      const uint32_t* __restrict__ offInt, const uint32_t* __restrict__ offTri, uint32_t n8Base, uint32_t slotBase,
                                                      Item8* __restrict__ nextItems, Node8* __restrict__ out8, TriSlot* __restrict__ slots,
                                                      float* __restrict__ nodeSlack) {
    const uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= itemCount) return;
    const Item8 it = items[ii];
    const Node2G& root = nodes[it.n2];
    int childAt[8];
    for (int s = 0; s < 8; s++) childAt[s] = childAtIn[(size_t)ii * 8 + s];
    const uint32_t childBase = n8Base + offInt[ii], triBase = slotBase + offTri[ii], itemBase = offInt[ii];

    Node8 N; memset(&N, 0, sizeof N);
    N.px = root.lo[0]; N.py = root.lo[1]; N.pz = root.lo[2];
    N.ex = quantExpDev(root.hi[0] - root.lo[0]); N.ey = quantExpDev(root.hi[1] - root.lo[1]); N.ez = quantExpDev(root.hi[2] - root.lo[2]);
    N.childBase = childBase; N.triBase = triBase; N.slack = root.maxShift;
    const uint8_t ebits[3] = {N.ex, N.ey, N.ez};
    const float pf3[3] = {N.px, N.py, N.pz};
    uint8_t* qlo[3] = {N.qlox, N.qloy, N.qloz}; uint8_t* qhi[3] = {N.qhix, N.qhiy, N.qhiz};
    uint32_t triOff = 0, rank = 0;
    for (int s = 0; s < 8; s++) {
        const int c2 = childAt[s];
        if (c2 < 0) continue;
        const Node2G& c = nodes[c2];
        for (int a = 0; a < 3; a++) {
            if (ebits[a] == 0) { qlo[a][s] = 0; qhi[a][s] = 0; continue; }
            const double sc = ldexp(1.0, (int)ebits[a] - 127), p = (double)pf3[a];
            int lo = (int)floor(((double)c.lo[a] - p) / sc), hi = (int)ceil(((double)c.hi[a] - p) / sc);
            lo = max(0, min(255, lo)); hi = max(0, min(255, hi));
            const float scf = (float)sc, pf = pf3[a];                // the float decode the traversal kernel performs must be conservative
            while (lo > 0 && __fadd_rn(pf, __fmul_rn((float)lo, scf)) > c.lo[a]) lo--;
            while (hi < 255 && __fadd_rn(pf, __fmul_rn((float)hi, scf)) < c.hi[a]) hi++;
            qlo[a][s] = (uint8_t)lo; qhi[a][s] = (uint8_t)hi;
        }
        if (c.left < 0) {
            const uint32_t cnt = c.count;                            // 1..3
            N.triMask |= ((1u << cnt) - 1u) << (3 * s);
            uint32_t t3[3];
            for (uint32_t k = 0; k < cnt; k++) t3[k] = idx[c.first + k];
            // arrival order of the partition atomics is not deterministic: emit in reference-id order
            if (cnt > 1 && t3[0] > t3[1]) { const uint32_t x = t3[0]; t3[0] = t3[1]; t3[1] = x; }
            if (cnt > 2 && t3[1] > t3[2]) { const uint32_t x = t3[1]; t3[1] = t3[2]; t3[2] = x; }
            if (cnt > 1 && t3[0] > t3[1]) { const uint32_t x = t3[0]; t3[0] = t3[1]; t3[1] = x; }
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t r = t3[k];
                const uint32_t t = r < nTris ? r : extraTri[r - nTris];      // reference -> triangle (pieces of pre-split triangles beyond the first)
                const ElevenTri& T = tris[t];
                TriSlot S;
                S.v0x = T.vertices[0][0]; S.v0y = T.vertices[0][1]; S.v0z = T.vertices[0][2];
                S.e1x = __fsub_rn(T.vertices[1][0], T.vertices[0][0]); S.e1y = __fsub_rn(T.vertices[1][1], T.vertices[0][1]); S.e1z = __fsub_rn(T.vertices[1][2], T.vertices[0][2]);
                S.e2x = __fsub_rn(T.vertices[2][0], T.vertices[0][0]); S.e2y = __fsub_rn(T.vertices[2][1], T.vertices[0][1]); S.e2z = __fsub_rn(T.vertices[2][2], T.vertices[0][2]);
                S.tri = (int32_t)t; S.material = triMaterial[t]; S.shiftBound = boxLo[r].w;
                slots[triBase + triOff + k] = S;
            }
            triOff += cnt;
        } else {
            N.imask |= (uint8_t)(1u << s);
            Item8 ni; ni.n2 = (uint32_t)c2; ni.n8 = childBase + rank; ni.depth = it.depth + 1u;     // internal children contiguous in ascending slot order
            nextItems[itemBase + rank] = ni;
            rank++;
        }
    }
    const float4* src = reinterpret_cast<const float4*>(&N);
    float4* dst = reinterpret_cast<float4*>(out8 + it.n8);
    for (int k = 0; k < 5; k++) dst[k] = src[k];
    nodeSlack[it.n8] = root.maxShift;
}

// shading records (9 x float4 per triangle, original order) straight from the uploaded ElevenTri array
__global__ void __launch_bounds__(256) k_shadeTris(const ElevenTri* __restrict__ tris, uint32_t n, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ElevenTri& T = tris[i];
    float* o = out + (size_t)i * 36;
    for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) { o[3 * k + a] = T.vertices[k][a]; o[9 + 3 * k + a] = T.normals[k][a]; o[18 + 3 * k + a] = T.tangents[k][a]; }
    o[27] = T.tangentsSign;
    o[28] = T.uv[0][0]; o[29] = T.uv[0][1]; o[30] = T.uv[1][0]; o[31] = T.uv[1][1]; o[32] = T.uv[2][0]; o[33] = T.uv[2][1];
    o[34] = __int_as_float(T.objectID); o[35] = 0.f;
}

// ---- host driver ---------------------------------------------------------------------------------------------------------------
struct DeviceBvh {
    float4* nodes = nullptr;       // nodeCount x 5 float4 (exact-size allocation, owned by the caller)
    float4* slots = nullptr;       // triCount x 3 float4
    float*  nodeSlack = nullptr;
    uint32_t nodeCount = 0, slotCount = 0, maxDepth = 0, levels = 0;
    float keySlack = 0.f;
    float boundsLo[3] = {0, 0, 0}, boundsHi[3] = {0, 0, 0};
    double buildMs = 0.0;
};

// Build scratch: ONE device allocation carved into the builder's arrays, owned by the context and kept for the next build.
// (cudaMalloc / cudaFree of the ~20 separate multi-GB buffers of the first version cost 5-10x the kernels: 155 ms + 190-590 ms
// against 56 ms of kernels for 10 M triangles.)
struct BuildArena { char* base = nullptr; size_t bytes = 0; };

enum { BIN_CHUNK_NODES = 131072 };   // nodes binned per launch: 131 072 x 1 536 B = 201 MB of bins whatever the scene

#define GB_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

// d_tris: the scene's triangles already on the device; d_triMaterial: per-triangle material.  n > 0.
static bool buildBvh8Device(const ElevenTri* d_tris, const int32_t* d_triMaterial, uint32_t nTris, const std::vector<PresplitPiece>& pieces, cudaStream_t st,
                            BuildArena& arena, DeviceBvh& out, std::string& err) {
    // references = triangles + the extra pieces of pre-split triangles; everything below the prep kernel works on references
    std::vector<uint32_t> refOf(pieces.size());
    uint32_t nExtra = 0;
    for (size_t k = 0; k < pieces.size(); k++) refOf[k] = (k == 0 || pieces[k].tri != pieces[k - 1].tri) ? pieces[k].tri : nTris + nExtra++;
    const uint32_t n = nTris + nExtra;
    const auto t0 = std::chrono::steady_clock::now();
    auto msSince = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    const bool verbose = getenv("ELEVEN_BVH_VERBOSE") != nullptr;

    for (int attempt = 0; attempt < 2; attempt++) {
    double tAlloc = 0, tPrep = 0, tLevels = 0, tCollapse = 0;
    const auto tA = std::chrono::steady_clock::now();
    // wide nodes: one per ~4.6 triangles in practice; n + 1 is the worst case (second attempt, after an overflow)
    // (ELEVEN_BVH_TEST_WIDE_DIV / ELEVEN_BVH_TEST_BIN_CHUNK: test knobs that force the overflow retry and the chunked binning on small scenes)
    const char* eDiv = getenv("ELEVEN_BVH_TEST_WIDE_DIV"); const char* eChunk = getenv("ELEVEN_BVH_TEST_BIN_CHUNK");
    const size_t wideDiv = eDiv ? std::max(2, atoi(eDiv)) : 2, binChunk = eChunk ? std::max(1, atoi(eChunk)) : (size_t)BIN_CHUNK_NODES;
    const size_t wideCap = attempt == 0 ? std::min<size_t>((size_t)n + 1, (size_t)n / wideDiv + (eDiv ? 2 : 1024)) : (size_t)n + 1;
    const size_t maxNodes = 2 * (size_t)n + 2, maxActive = (size_t)n / 2 + 2;
    const size_t binNodes = std::min<size_t>(maxActive, binChunk);
    size_t scanBytes = 0;
    GB_CK(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)wideCap, st));

    float4 *boxLo, *boxHi; uint32_t *idxA, *idxB, *ownA, *ownB, *scene, *counters, *actA, *actB, *bins; Node2G* nodes;
    Node8* out8; TriSlot* slots; float* slack; Item8 *itA, *itB; int32_t* childAt; uint32_t *nInt, *nTri, *offInt, *offTri; void* scanTmp;
    float* dpC; uint8_t* dpK; PresplitPiece* dPieces; uint32_t *dRefOf, *extraTri;
    auto layout = [&](char* base) -> size_t {
        size_t off = 0;
        auto take = [&](size_t bytes) -> char* { off = (off + 255) & ~(size_t)255; char* r = base ? base + off : nullptr; off += bytes; return r; };
        boxLo = (float4*)take((size_t)n * 16); boxHi = (float4*)take((size_t)n * 16);
        idxA = (uint32_t*)take((size_t)n * 4); idxB = (uint32_t*)take((size_t)n * 4); ownA = (uint32_t*)take((size_t)n * 4); ownB = (uint32_t*)take((size_t)n * 4);
        scene = (uint32_t*)take(SCENE_WORDS * 4); counters = (uint32_t*)take(C_COUNT * 4);
        nodes = (Node2G*)take(maxNodes * sizeof(Node2G));
        actA = (uint32_t*)take(maxActive * 4); actB = (uint32_t*)take(maxActive * 4);
        bins = (uint32_t*)take(binNodes * NODE_BIN_WORDS * 4);
        out8 = (Node8*)take(wideCap * sizeof(Node8)); slots = (TriSlot*)take((size_t)n * sizeof(TriSlot)); slack = (float*)take(wideCap * 4);
        itA = (Item8*)take(wideCap * sizeof(Item8)); itB = (Item8*)take(wideCap * sizeof(Item8));
        childAt = (int32_t*)take(wideCap * 8 * 4);
        nInt = (uint32_t*)take(wideCap * 4); nTri = (uint32_t*)take(wideCap * 4); offInt = (uint32_t*)take(wideCap * 4); offTri = (uint32_t*)take(wideCap * 4);
        scanTmp = take(scanBytes);
        dpC = (float*)take(maxNodes * 8 * 4); dpK = (uint8_t*)take(maxNodes * 8);
        dPieces = (PresplitPiece*)take(pieces.size() * sizeof(PresplitPiece)); dRefOf = (uint32_t*)take(pieces.size() * 4); extraTri = (uint32_t*)take((size_t)nExtra * 4 + 4);
        return off;
    };
    const size_t need = layout(nullptr);
    if (arena.bytes < need) {
        if (arena.base) cudaFree(arena.base);
        arena.base = nullptr; arena.bytes = 0;
        GB_CK(cudaMalloc((void**)&arena.base, need));
        arena.bytes = need;
    }
    layout(arena.base);
    tAlloc = msSince(tA);

    const uint32_t sceneInit[SCENE_WORDS] = {EL_ENC_POS_INF, EL_ENC_POS_INF, EL_ENC_POS_INF, EL_ENC_NEG_INF, EL_ENC_NEG_INF, EL_ENC_NEG_INF, 0u, 0u};
    GB_CK(cudaMemcpyAsync(scene, sceneInit, sizeof sceneInit, cudaMemcpyHostToDevice, st));
    const int gridN = (int)((n + 255) / 256);
    k_prep<<<(int)((nTris + 255) / 256), 256, 0, st>>>(d_tris, nTris, boxLo, boxHi, idxA, ownA, scene);
    if (!pieces.empty()) {
        GB_CK(cudaMemcpyAsync(dPieces, pieces.data(), pieces.size() * sizeof(PresplitPiece), cudaMemcpyHostToDevice, st));
        GB_CK(cudaMemcpyAsync(dRefOf, refOf.data(), refOf.size() * 4, cudaMemcpyHostToDevice, st));
        k_applyPieces<<<(unsigned)((pieces.size() + 255) / 256), 256, 0, st>>>(dPieces, dRefOf, (uint32_t)pieces.size(), boxLo, boxHi, idxA, ownA, nTris, extraTri);
    }
    uint32_t sceneHost[SCENE_WORDS];
    GB_CK(cudaMemcpyAsync(sceneHost, scene, sizeof sceneHost, cudaMemcpyDeviceToHost, st));
    GB_CK(cudaStreamSynchronize(st));
    float ext = 0.f, mag = 0.f, maxShift;
    memcpy(&maxShift, &sceneHost[SCENE_SHIFT], 4);
    for (int a = 0; a < 3; a++) {
        out.boundsLo[a] = decF(sceneHost[SCENE_LO + a]); out.boundsHi[a] = decF(sceneHost[SCENE_HI + a]);
        ext = std::max(ext, out.boundsHi[a] - out.boundsLo[a]);
        mag = std::max(mag, std::max(std::fabs(out.boundsLo[a]), std::fabs(out.boundsHi[a])));
    }
    const float pad = 4e-6f * std::max(ext, mag) + 1e-30f;            // same padding rule as the host builder
    out.keySlack = (float)((double)maxShift + 1e-5 * (double)std::max(ext, mag));

    Node2G rootN; memset(&rootN, 0, sizeof rootN);
    for (int a = 0; a < 3; a++) { rootN.lo[a] = out.boundsLo[a] - pad; rootN.hi[a] = out.boundsHi[a] + pad; rootN.clo[a] = EL_ENC_POS_INF; rootN.chi[a] = EL_ENC_NEG_INF; }
    rootN.left = -1; rootN.first = 0; rootN.count = n; rootN.axis = -1; rootN.maxShift = maxShift;
    uint32_t activeCount = n > 1u ? 1u : 0u;
    rootN.binSlot = activeCount ? 0 : -1;
    GB_CK(cudaMemcpyAsync(nodes, &rootN, sizeof rootN, cudaMemcpyHostToDevice, st));
    uint32_t cnt[C_COUNT] = {1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    GB_CK(cudaMemcpyAsync(counters, cnt, sizeof cnt, cudaMemcpyHostToDevice, st));
    const uint32_t zero = 0u;
    GB_CK(cudaMemcpyAsync(actA, &zero, 4, cudaMemcpyHostToDevice, st));
    tPrep = msSince(tA) - tAlloc;

    uint32_t level = 0;
    std::vector<std::pair<uint32_t, uint32_t>> levelRange;          // (first id, count) of the binary nodes created per level
    levelRange.push_back({0u, 1u});
    uint32_t nodesSoFar = 1;
    while (activeCount > 0) {
        if (level > 96) { err = "device BVH build: depth guard exceeded"; return false; }
        k_centroid<<<gridN, 256, 0, st>>>(boxLo, boxHi, idxA, ownA, nodes, n);
        for (uint32_t c0 = 0; c0 < activeCount; c0 += (uint32_t)binNodes) {
            const uint32_t cn = std::min<uint32_t>((uint32_t)binNodes, activeCount - c0);
            const size_t words = (size_t)cn * NODE_BIN_WORDS;
            k_initBins<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(bins, words);
            k_bin<<<gridN, 256, 0, st>>>(boxLo, boxHi, idxA, ownA, nodes, bins, n, pad, (int)c0, (int)cn);
            k_split<<<(cn + 127) / 128, 128, 0, st>>>(nodes, bins, actA + c0, cn, actB, counters, level >= 48 ? 1 : 0, (int)c0);
        }
        k_partition<<<gridN, 256, 0, st>>>(boxLo, boxHi, idxA, ownA, idxB, ownB, nodes, n);
        k_retire<<<(activeCount + 255) / 256, 256, 0, st>>>(nodes, actA, activeCount);
        uint32_t next = 0, nodeTotal = 0;
        GB_CK(cudaMemcpyAsync(&next, counters + C_NEXT_ACTIVE, 4, cudaMemcpyDeviceToHost, st));
        GB_CK(cudaMemcpyAsync(&nodeTotal, counters + C_NODES, 4, cudaMemcpyDeviceToHost, st));
        GB_CK(cudaMemcpyAsync(counters + C_NEXT_ACTIVE, &zero, 4, cudaMemcpyHostToDevice, st));
        GB_CK(cudaStreamSynchronize(st));
        GB_CK(cudaGetLastError());
        levelRange.push_back({nodesSoFar, nodeTotal - nodesSoFar});
        nodesSoFar = nodeTotal;
        std::swap(idxA, idxB); std::swap(ownA, ownB); std::swap(actA, actB);
        activeCount = next; level++;
    }
    out.levels = level;
    tLevels = msSince(tA) - tAlloc - tPrep;

    // ---- collapse + emission ------------------------------------------------------------------------------------------------
    for (size_t l = levelRange.size(); l-- > 0;)
        if (levelRange[l].second) k_collapseDp<<<(levelRange[l].second + 127) / 128, 128, 0, st>>>(nodes, levelRange[l].first, levelRange[l].second, dpC, dpK);
    Item8 first; first.n2 = 0; first.n8 = 0; first.depth = 1;
    GB_CK(cudaMemcpyAsync(itA, &first, sizeof first, cudaMemcpyHostToDevice, st));
    uint32_t itemCount = 1, n8Base = 1, slotBase = 0, depth = 0;
    bool overflow = false;
    while (itemCount > 0) {
        depth++;
        const unsigned grid = (itemCount + 127) / 128;
        k_collapseGather<<<grid, 128, 0, st>>>(nodes, dpK, itA, itemCount, childAt, nInt, nTri);
        GB_CK(cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, nInt, offInt, (int)itemCount, st));
        GB_CK(cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, nTri, offTri, (int)itemCount, st));
        uint32_t last[4];
        GB_CK(cudaMemcpyAsync(&last[0], offInt + itemCount - 1, 4, cudaMemcpyDeviceToHost, st));
        GB_CK(cudaMemcpyAsync(&last[1], nInt + itemCount - 1, 4, cudaMemcpyDeviceToHost, st));
        GB_CK(cudaMemcpyAsync(&last[2], offTri + itemCount - 1, 4, cudaMemcpyDeviceToHost, st));
        GB_CK(cudaMemcpyAsync(&last[3], nTri + itemCount - 1, 4, cudaMemcpyDeviceToHost, st));
        GB_CK(cudaStreamSynchronize(st));
        const uint32_t next = last[0] + last[1];
        if ((size_t)n8Base + next > wideCap || (size_t)slotBase + last[2] + last[3] > n) { overflow = true; break; }   // before anything is written out of bounds
        k_collapseEmit<<<grid, 128, 0, st>>>(nodes, idxA, d_tris, d_triMaterial, boxLo, nTris, extraTri, itA, itemCount, childAt, offInt, offTri, n8Base, slotBase,
                                             itB, out8, slots, slack);
        GB_CK(cudaGetLastError());
        n8Base += next; slotBase += last[2] + last[3];
        std::swap(itA, itB);
        itemCount = next;
    }
    if (overflow) {
        if (attempt == 0 && wideCap < (size_t)n + 1) continue;       // more wide nodes than n/2: rebuild with worst-case buffers
        err = "device BVH build: collapse overflow"; return false;
    }
    GB_CK(cudaStreamSynchronize(st));
    out.nodeCount = n8Base; out.slotCount = slotBase; out.maxDepth = depth;
    tCollapse = msSince(tA) - tAlloc - tPrep - tLevels;
    if (out.slotCount != n) { err = "device BVH build: emitted " + std::to_string(out.slotCount) + " triangle slots for " + std::to_string(n) + " references"; return false; }

    // exact-size results (the arena is scratch)
    void *rn = nullptr, *rs = nullptr, *rk = nullptr;
    cudaError_t e1 = cudaMalloc(&rn, (size_t)out.nodeCount * sizeof(Node8)), e2 = cudaMalloc(&rs, (size_t)n * sizeof(TriSlot)), e3 = cudaMalloc(&rk, (size_t)out.nodeCount * 4);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cudaFree(rn); cudaFree(rs); cudaFree(rk); err = "device BVH build: out of memory"; return false; }
    cudaMemcpyAsync(rn, out8, (size_t)out.nodeCount * sizeof(Node8), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(rs, slots, (size_t)n * sizeof(TriSlot), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(rk, slack, (size_t)out.nodeCount * 4, cudaMemcpyDeviceToDevice, st);
    cudaError_t es = cudaStreamSynchronize(st);
    if (es != cudaSuccess) { cudaFree(rn); cudaFree(rs); cudaFree(rk); err = std::string("device BVH build: ") + cudaGetErrorString(es); return false; }
    out.nodes = (float4*)rn; out.slots = (float4*)rs; out.nodeSlack = (float*)rk;
    out.buildMs = msSince(t0);
    if (verbose) fprintf(stderr, "[eleven] device BVH: %u tris, %u levels, %u wide nodes, depth %u, scratch %.0f MB: alloc %.2f ms, prep %.2f, levels %.2f, collapse %.2f, results %.2f, total %.2f ms\n",
                         n, level, out.nodeCount, depth, need / 1048576.0, tAlloc, tPrep, tLevels, tCollapse, msSince(tA) - tAlloc - tPrep - tLevels - tCollapse, out.buildMs);
    return true;
    }
    err = "device BVH build: unreachable";
    return false;
}
#undef GB_CK

} // namespace gpubvh
} // namespace eleven
