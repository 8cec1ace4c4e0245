/*
 * bvh8_build.cpp — host builder for the compressed 8-wide BVH (see bvh8.h).
 *
 * Replaces S/BVH.hpp:187-330 + divideSAH :373-460 (recursive, single-threaded, fixed depth 18, std::vector copies
 * of 156-byte structs at every level).  Differences by design: adaptive depth, SAH leaf termination (<= EL_MAX_LEAF = 2
 * triangles), in-place index partition, O(bins) sweep, task-parallel subtrees, wide collapse, quantisation.
 */
#include "bvh8.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <future>
#include <thread>

namespace eleven {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float area() const {
        float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return (x < 0 || y < 0 || z < 0) ? 0.f : 2.f * (x * y + x * z + y * z);
    }
};

struct Node2 { Box box; int left, right, first, count, span; };   // leaf iff count > 0; span = triangles in the subtree

static const int   MAXBINS = 64;
// ELEVEN_BVH_BINS (experiments; read once when the library is loaded, so concurrent builds never write it); 16 is what the device builder uses
static const int BINS = [] { const char* e = getenv("ELEVEN_BVH_BINS"); return e ? std::max(2, std::min(MAXBINS, atoi(e))) : 16; }();
static const int   MAX_LEAF = EL_MAX_LEAF;                 // bvh8.h
static const float COST_TRI = 1.0f, COST_NODE = EL_LEAF_COST_NODE;

struct Builder {
    const ElevenTri* tris; uint32_t n;
    std::vector<Box> tbox; std::vector<float> cen;   // cen[3*i+a]
    std::vector<int> idx;
    std::vector<Node2> nodes; std::atomic<int> nodeCount;
    int threads; std::atomic<int> liveTasks;

    int alloc() { return nodeCount.fetch_add(1); }

    void build(int ni, int first, int count, int depth) {
        Node2& N = nodes[ni];
        Box b, cb; b.reset(); cb.reset();
        for (int i = first; i < first + count; i++) { b.grow(tbox[idx[i]]); cb.grow(&cen[3 * idx[i]]); }
        N.box = b; N.left = N.right = -1; N.first = first; N.count = count; N.span = count;
        if (count <= 1) return;

        // binned SAH over the three axes
        float bestCost = INFINITY; int bestAxis = -1, bestSplit = -1;
        for (int a = 0; a < 3; a++) {
            float lo = cb.lo[a], ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0)) continue;
            Box bb[MAXBINS]; int bc[MAXBINS];
            for (int k = 0; k < BINS; k++) { bb[k].reset(); bc[k] = 0; }
            float scale = BINS / ext;
            for (int i = first; i < first + count; i++) {
                int t = idx[i];
                int k = std::min(BINS - 1, std::max(0, (int)((cen[3 * t + a] - lo) * scale)));
                bc[k]++; bb[k].grow(tbox[t]);
            }
            float rightArea[MAXBINS]; int rightCount[MAXBINS];
            Box acc; acc.reset(); int c = 0;
            for (int k = BINS - 1; k > 0; k--) { acc.grow(bb[k]); c += bc[k]; rightArea[k] = acc.area(); rightCount[k] = c; }
            acc.reset(); c = 0;
            for (int k = 1; k < BINS; k++) {
                acc.grow(bb[k - 1]); c += bc[k - 1];
                if (c == 0 || rightCount[k] == 0) continue;
                float cost = acc.area() * c + rightArea[k] * rightCount[k];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = k; }
            }
        }
        float leafCost = COST_TRI * count * b.area();
        float splitCost = COST_NODE * b.area() + COST_TRI * bestCost;
        if (count <= MAX_LEAF && (bestAxis < 0 || leafCost <= splitCost)) return;

        int mid;
        if (bestAxis >= 0) {
            float lo = cb.lo[bestAxis], scale = BINS / (cb.hi[bestAxis] - cb.lo[bestAxis]);
            int a = bestAxis, s = bestSplit;
            int* p = std::partition(&idx[first], &idx[first] + count, [&](int t) {
                int k = std::min(BINS - 1, std::max(0, (int)((cen[3 * t + a] - lo) * scale)));
                return k < s;
            });
            mid = (int)(p - &idx[0]);
        } else {
            mid = first + count / 2;   // all centroids coincide: split by index
        }
        if (mid == first || mid == first + count) mid = first + count / 2;

        int l = alloc(), r = alloc();
        N.left = l; N.right = r; N.count = 0;
        int lc = mid - first, rc = first + count - mid;
        bool fork = threads > 1 && std::min(lc, rc) > 8192 && liveTasks.load() < threads - 1;
        if (fork) {
            liveTasks.fetch_add(1);
            auto fut = std::async(std::launch::async, [=]() { build(l, first, lc, depth + 1); liveTasks.fetch_sub(1); });
            build(r, mid, rc, depth + 1);
            fut.get();
        } else {
            build(l, first, lc, depth + 1);
            build(r, mid, rc, depth + 1);
        }
    }
};

static inline bool isLeaf2(const Node2& n) { return n.count > 0; }

// largest power-of-two exponent e (biased by 127, stored in a byte) with 2^e * 255 >= extent
static uint8_t quantExp(float extent) {
    if (!(extent > 0)) return 0;
    int e = (int)std::ceil(std::log2((double)extent / 255.0));
    while (std::ldexp(255.0, e) < (double)extent) e++;
    e = std::max(-126, std::min(127, e));
    return (uint8_t)(e + 127);
}


// ---- triangle pre-splitting ("early split clipping") -------------------------------------------------------------------------
// A binned-SAH tree over whole-triangle boxes is helpless against slivers whose boxes overlap: the ~100 wedges of a disc modelled as a
// fan all own most of the disc's box, and a ray through the disc tests every one of them (the clock face of the benchmark scene: camera
// rays that hit it ran 80-99 triangle tests for 1 hit, 10 % of all camera rays more than 37; CPU walk of tests/bvh8_walk.py).  Such a
// triangle is handed to the builder as several REFERENCES, each with the box of the part of the triangle inside one cell of a recursive
// midpoint split of its box: the pieces' boxes are small and disjoint, the leaves simply list the triangle more than once (a ray may
// test it twice; Moeller-Trumbore gives the same answer twice).  Split rule, per triangle: box area > PRESPLIT_AREA x the mean box
// area of the scene AND box area > PRESPLIT_SLIVER x twice the triangle's own area (a well-shaped triangle fills its box; plain large
// triangles and the millions of similar triangles of a grid are left alone), pieces until their boxes are below half the mean, <= 2^5 each.
struct Ref { Box box; int tri; };

static void clipPoly(const std::vector<double>& in, int axis, double pos, bool keepBelow, std::vector<double>& out) {   // Sutherland-Hodgman against one axis-aligned plane
    out.clear();
    const size_t m = in.size() / 3;
    for (size_t i = 0; i < m; i++) {
        const double* a = &in[3 * i]; const double* b = &in[3 * ((i + 1) % m)];
        const bool ina = keepBelow ? a[axis] <= pos : a[axis] >= pos, inb = keepBelow ? b[axis] <= pos : b[axis] >= pos;
        if (ina) out.insert(out.end(), a, a + 3);
        if (ina != inb) {
            const double t = (pos - a[axis]) / (b[axis] - a[axis]);
            double q[3]; for (int k = 0; k < 3; k++) q[k] = a[k] + t * (b[k] - a[k]);
            q[axis] = pos;
            out.insert(out.end(), q, q + 3);
        }
    }
}
static void splitRec(const std::vector<double>& poly, const Box& box, int tri, float targetArea, int depth, std::vector<Ref>& out) {
    if (depth == 0 || box.area() <= targetArea || poly.size() < 9) { out.push_back({box, tri}); return; }
    int axis = 0; float ext = box.hi[0] - box.lo[0];
    for (int a = 1; a < 3; a++) if (box.hi[a] - box.lo[a] > ext) { ext = box.hi[a] - box.lo[a]; axis = a; }
    const double mid = 0.5 * ((double)box.lo[axis] + (double)box.hi[axis]);
    std::vector<double> part[2];
    clipPoly(poly, axis, mid, true, part[0]); clipPoly(poly, axis, mid, false, part[1]);
    if (part[0].size() < 9 || part[1].size() < 9) { out.push_back({box, tri}); return; }
    for (int h = 0; h < 2; h++) {
        Box b; b.reset();
        for (size_t i = 0; i < part[h].size() / 3; i++) for (int a = 0; a < 3; a++) {
            const double v = part[h][3 * i + a];
            b.lo[a] = std::min(b.lo[a], std::nextafter((float)v, -INFINITY)); b.hi[a] = std::max(b.hi[a], std::nextafter((float)v, INFINITY));   // float box of double points, outwards
        }
        for (int a = 0; a < 3; a++) { b.lo[a] = std::max(b.lo[a], box.lo[a]); b.hi[a] = std::min(b.hi[a], box.hi[a]); }
        splitRec(part[h], b, tri, targetArea, depth - 1, out);
    }
}
static const float PRESPLIT_AREA = [] { const char* e = getenv("ELEVEN_PRESPLIT_AREA"); return e ? (float)atof(e) : 2.0f; }();
static const float PRESPLIT_SLIVER = [] { const char* e = getenv("ELEVEN_PRESPLIT_SLIVER"); return e ? (float)atof(e) : 8.0f; }();
static const float PRESPLIT_TARGET = [] { const char* e = getenv("ELEVEN_PRESPLIT_TARGET"); return e ? (float)atof(e) : 0.5f; }();   // pieces until their box is below this x the mean (CPU walk: 1.0 / 0.5 / 0.25 -> 7.02 / 6.59 / 6.23 camera-ray tests at 14 119 / 14 334 / 14 491 nodes)
static const int PRESPLIT_DEPTH = [] { const char* e = getenv("ELEVEN_PRESPLIT_DEPTH"); return e ? atoi(e) : 5; }();

} // namespace

// Pieces of the triangles the rule above splits: out[k] = (box, triangle).  A split triangle contributes >= 2 pieces, an unsplit one none.
void presplitTriangles(const ElevenTri* tris, uint32_t n, std::vector<PresplitPiece>& out) {
    out.clear();
    if (n == 0 || PRESPLIT_DEPTH <= 0) return;
    // fixed chunks, combined in chunk order: the result does not depend on the number of threads (both builders and every GPU of a job
    // must see the same pieces)
    const uint32_t CHUNK = 65536, chunks = (n + CHUNK - 1) / CHUNK;
    const unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), chunks));
    auto parallel = [&](const std::function<void(uint32_t)>& body) {
        std::atomic<uint32_t> next(0);
        auto run = [&]() { for (;;) { const uint32_t c = next.fetch_add(1); if (c >= chunks) break; body(c); } };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nth; t++) th.emplace_back(run);
        run();
        for (auto& t : th) t.join();
    };
    std::vector<float> area(n);
    std::vector<double> chunkSum(chunks, 0.0);
    parallel([&](uint32_t c) {
        double sum = 0;
        for (uint32_t i = c * CHUNK, e = std::min(n, (c + 1) * CHUNK); i < e; i++) { Box b; b.reset(); for (int k = 0; k < 3; k++) b.grow(tris[i].vertices[k]); area[i] = b.area(); sum += area[i]; }
        chunkSum[c] = sum;
    });
    double sum = 0; for (double v : chunkSum) sum += v;
    const float mean = (float)(sum / n);
    std::vector<std::vector<PresplitPiece>> chunkOut(chunks);
    parallel([&](uint32_t c) {
        std::vector<Ref> pieces; std::vector<double> poly(9);
        for (uint32_t i = c * CHUNK, e = std::min(n, (c + 1) * CHUNK); i < e; i++) {
            if (!(area[i] > PRESPLIT_AREA * mean)) continue;
            const float (*v)[3] = tris[i].vertices;
            double e1[3], e2[3]; for (int a = 0; a < 3; a++) { e1[a] = (double)v[1][a] - v[0][a]; e2[a] = (double)v[2][a] - v[0][a]; }
            const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
            const double twiceTri = std::sqrt(cx * cx + cy * cy + cz * cz);          // 2 x the triangle's area = the area of a box that fits it flat
            if (!((double)area[i] > (double)PRESPLIT_SLIVER * twiceTri)) continue;
            Box b; b.reset(); for (int k = 0; k < 3; k++) b.grow(v[k]);
            for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) poly[3 * k + a] = v[k][a];
            pieces.clear();
            splitRec(poly, b, (int)i, PRESPLIT_TARGET * mean, PRESPLIT_DEPTH, pieces);
            if (pieces.size() < 2) continue;
            for (const Ref& r : pieces) { PresplitPiece p; p.tri = (uint32_t)r.tri; for (int a = 0; a < 3; a++) { p.lo[a] = r.box.lo[a]; p.hi[a] = r.box.hi[a]; } chunkOut[c].push_back(p); }
        }
    });
    for (auto& v : chunkOut) out.insert(out.end(), v.begin(), v.end());
}

void presplitParams(float& areaFactor, float& sliverFactor, int& depth) { areaFactor = PRESPLIT_AREA; sliverFactor = PRESPLIT_SLIVER; depth = PRESPLIT_DEPTH; }

// The clipping half on its own: `cand` (ascending triangle indices) were selected by the rule above — on the device, by the device builder.
void presplitCandidates(const ElevenTri* tris, const uint32_t* cand, size_t nc, float meanArea, std::vector<PresplitPiece>& out) {
    out.clear();
    std::vector<Ref> pieces; std::vector<double> poly(9);
    for (size_t c = 0; c < nc; c++) {
        const uint32_t i = cand[c];
        const float (*v)[3] = tris[i].vertices;
        Box b; b.reset(); for (int k = 0; k < 3; k++) b.grow(v[k]);
        for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) poly[3 * k + a] = v[k][a];
        pieces.clear();
        splitRec(poly, b, (int)i, PRESPLIT_TARGET * meanArea, PRESPLIT_DEPTH, pieces);
        if (pieces.size() < 2) continue;
        for (const Ref& r : pieces) { PresplitPiece p; p.tri = (uint32_t)r.tri; for (int a = 0; a < 3; a++) { p.lo[a] = r.box.lo[a]; p.hi[a] = r.box.hi[a]; } out.push_back(p); }
    }
}

void buildBvh8(const ElevenTri* tris, uint32_t n, const int32_t* triMaterial, Bvh8& out, int threads, bool presplit) {
    auto t0 = std::chrono::steady_clock::now();
    out.nodes.clear(); out.slots.clear(); out.keySlack = 0; out.maxDepth = 0;
    for (int a = 0; a < 3; a++) { out.boundsLo[a] = 0; out.boundsHi[a] = 0; }
    if (threads < 1) threads = 1;

    // references: one per triangle, or the pieces of a pre-split triangle (the first piece takes the triangle's own place)
    std::vector<PresplitPiece> pieces;
    if (const char* e = getenv("ELEVEN_PRESPLIT")) presplit = atoi(e) != 0;     // A/B knob
    if (presplit) presplitTriangles(tris, n, pieces);
    std::vector<int> refTri(n);
    for (uint32_t i = 0; i < n; i++) refTri[i] = (int)i;
    const uint32_t nr = n + (uint32_t)pieces.size() - [&] { uint32_t c = 0; for (size_t k = 0; k < pieces.size(); k++) if (k == 0 || pieces[k].tri != pieces[k - 1].tri) c++; return c; }();
    Builder B; B.tris = tris; B.n = nr; B.threads = threads; B.liveTasks = 0;
    B.tbox.resize(nr); B.cen.resize(3 * (size_t)nr); B.idx.resize(nr);
    refTri.resize(nr);
    Box scene; scene.reset();
    double slack = 0;
    std::vector<float> triShift(n);
    for (uint32_t i = 0; i < n; i++) {
        Box b; b.reset();
        for (int k = 0; k < 3; k++) b.grow(tris[i].vertices[k]);
        B.tbox[i] = b; scene.grow(b);
        B.idx[i] = (int)i;
        // Bound on the shadow-terminator shift |shadingPosition - geomPosition| (S/Tri.hpp:81-89): shadingPosition is
        // a convex combination of the projections p_i = P - dot(P - v_i, n_i) n_i, and dot(P - v_i, n_i) is linear in
        // P over the triangle, so max |p_i - P| is attained at a vertex: max_{i,j} |dot(v_j - v_i, n_i)| * |n_i|.
        double mine = 0;
        for (int vi = 0; vi < 3; vi++) {
            const float* nn = tris[i].normals[vi];
            double nl = std::sqrt((double)nn[0] * nn[0] + (double)nn[1] * nn[1] + (double)nn[2] * nn[2]);
            for (int vj = 0; vj < 3; vj++) {
                if (vj == vi) continue;
                double d = 0;
                for (int a = 0; a < 3; a++) d += ((double)tris[i].vertices[vj][a] - tris[i].vertices[vi][a]) * nn[a];
                mine = std::max(mine, std::fabs(d) * nl);
            }
        }
        triShift[i] = (float)(mine * 1.0001);
        slack = std::max(slack, mine);
    }
    {
        uint32_t next = n;
        for (size_t k = 0; k < pieces.size(); k++) {
            const bool first = k == 0 || pieces[k].tri != pieces[k - 1].tri;
            const uint32_t r = first ? pieces[k].tri : next++;
            for (int a = 0; a < 3; a++) { B.tbox[r].lo[a] = pieces[k].lo[a]; B.tbox[r].hi[a] = pieces[k].hi[a]; }
            refTri[r] = (int)pieces[k].tri; B.idx[r] = (int)r;
        }
    }
    for (uint32_t r = 0; r < nr; r++) for (int a = 0; a < 3; a++) B.cen[3 * (size_t)r + a] = 0.5f * (B.tbox[r].lo[a] + B.tbox[r].hi[a]);
    if (n == 0) { scene.lo[0] = scene.lo[1] = scene.lo[2] = 0; scene.hi[0] = scene.hi[1] = scene.hi[2] = 0; }
    float ext = 0, mag = 0;
    for (int a = 0; a < 3; a++) {
        out.boundsLo[a] = scene.lo[a]; out.boundsHi[a] = scene.hi[a];
        ext = std::max(ext, scene.hi[a] - scene.lo[a]);
        mag = std::max(mag, std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])));
    }
    // pad triangle boxes so that float rounding in the slab test can never cull a triangle Moeller-Trumbore accepts
    float pad = 4e-6f * std::max(ext, mag) + 1e-30f;
    for (uint32_t i = 0; i < nr; i++) for (int a = 0; a < 3; a++) { B.tbox[i].lo[a] -= pad; B.tbox[i].hi[a] += pad; }
    out.keySlack = (float)(slack * 1.0001 + 1e-5 * std::max(ext, mag));

    // ---- binary binned-SAH build ----------------------------------------------------------------
    B.nodes.resize(std::max<size_t>(1, 2 * (size_t)nr));
    B.nodeCount = 1;
    if (n > 0) B.build(0, 0, (int)nr, 0);
    else { B.nodes[0].box = scene; B.nodes[0].left = B.nodes[0].right = -1; B.nodes[0].first = 0; B.nodes[0].count = 0; }

    // ---- optional: SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1) -----------------------------------------------
    // C(n, i) = lowest SAH cost of the subtree of binary node n when it may occupy i child slots of its wide parent:
    //   leaf:      C(n, i) = A_n * count * c_tri
    //   internal:  D(n, j) = min_{0<k<j} C(l, k) + C(r, j-k);  C(n, 1) = D(n, 8) + A_n * c_node;  C(n, i) = min(D(n, i), C(n, i-1))
    // Children are allocated after their parent, so a descending index sweep is a post-order.
    const char* optEnv = getenv("ELEVEN_BVH_COLLAPSE");
    const bool optimal = !(optEnv && !strcmp(optEnv, "greedy"));     // default; "greedy" = open the largest child (round-1 first version)
    const float cNode = getenv("ELEVEN_COST_NODE") ? (float)atof(getenv("ELEVEN_COST_NODE")) : 3.0f;
    struct Dp { float c[9]; uint8_t k[9]; };     // k[j] for j >= 2: split of D(n, j), or 0 = "take C(n, j-1)"; k[1] = split of D(n, 8)
    std::vector<Dp> dp;
    if (optimal && n > 0) {
        const int nn = B.nodeCount.load();
        dp.resize(nn);
        for (int i = nn - 1; i >= 0; i--) {
            const Node2& N = B.nodes[i]; Dp& d = dp[i];
            const float area = N.box.area();
            if (isLeaf2(N) || N.left < 0) { for (int j = 0; j < 9; j++) { d.c[j] = area * (float)N.count * COST_TRI; d.k[j] = 0; } continue; }
            const Dp &L = dp[N.left], &R = dp[N.right];
            float D[9]; uint8_t K[9];
            for (int j = 2; j <= 8; j++) {
                D[j] = INFINITY; K[j] = 1;
                for (int k = 1; k < j; k++) { const float v = L.c[std::min(k, 7)] + R.c[std::min(j - k, 7)]; if (v < D[j]) { D[j] = v; K[j] = (uint8_t)k; } }
            }
            d.c[0] = INFINITY; d.k[0] = 0;
            d.c[1] = D[8] + area * cNode; d.k[1] = K[8];
            for (int j = 2; j <= 7; j++) { if (D[j] < d.c[j - 1]) { d.c[j] = D[j]; d.k[j] = K[j]; } else { d.c[j] = d.c[j - 1]; d.k[j] = 0; } }
            d.c[8] = d.c[7]; d.k[8] = 0;
        }
    }
    // children of the wide node rooted at binary node m when it may use j slots
    std::function<void(int, int, int*, int&)> expand = [&](int m, int j, int* ch, int& nc) {
        const Node2& M = B.nodes[m];
        if (isLeaf2(M) || M.left < 0 || j == 1) { ch[nc++] = m; return; }
        while (j > 1 && dp[m].k[j] == 0) j--;
        if (j == 1) { ch[nc++] = m; return; }
        const int k = dp[m].k[j];
        expand(M.left, k, ch, nc); expand(M.right, j - k, ch, nc);
    };

    // ---- collapse to 8-wide, breadth-first so that a node's internal children are contiguous ---------
    struct Item { int n2; uint32_t n8; uint32_t depth; };
    std::vector<Item> queue; queue.reserve(nr / 4 + 16);
    out.nodes.reserve(nr / 3 + 16); out.slots.reserve(nr); out.nodeSlack.clear(); out.nodeSlack.reserve(n / 3 + 16);
    out.nodes.emplace_back(); memset(&out.nodes[0], 0, sizeof(Node8));
    queue.push_back({0, 0u, 1u});
    for (size_t qi = 0; qi < queue.size(); qi++) {
        Item it = queue[qi];
        out.maxDepth = std::max(out.maxDepth, it.depth);
        const Node2& root = B.nodes[it.n2];
        int ch[8]; int nc = 0;
        if (isLeaf2(root) || root.left < 0) { if (root.count > 0) ch[nc++] = it.n2; }
        else if (optimal) { const int k = dp[it.n2].k[1]; expand(root.left, k, ch, nc); expand(root.right, 8 - k, ch, nc); }
        else { ch[nc++] = root.left; ch[nc++] = root.right; }
        while (!optimal && nc < 8) {                       // greedy: open the internal child with the largest area
            int best = -1; float bestA = -1;
            for (int i = 0; i < nc; i++) { const Node2& c = B.nodes[ch[i]]; if (!isLeaf2(c)) { float a = c.box.area(); if (a > bestA) { bestA = a; best = i; } } }
            if (best < 0) break;
            const Node2& c = B.nodes[ch[best]];
            ch[best] = c.left; ch[nc++] = c.right;
        }
        // octant-ordered slots: slot s "lies" in direction ((s&4)?+:-, (s&2)?+:-, (s&1)?+:-) from the node centre
        Box nb = root.box;
        float cx[3] = {0.5f * (nb.lo[0] + nb.hi[0]), 0.5f * (nb.lo[1] + nb.hi[1]), 0.5f * (nb.lo[2] + nb.hi[2])};
        float cost[8][8]; int slotOf[8]; bool cUsed[8] = {false}, sUsed[8] = {false};
        for (int c = 0; c < nc; c++) {
            const Box& cb = B.nodes[ch[c]].box;
            float d[3] = {0.5f * (cb.lo[0] + cb.hi[0]) - cx[0], 0.5f * (cb.lo[1] + cb.hi[1]) - cx[1], 0.5f * (cb.lo[2] + cb.hi[2]) - cx[2]};
            for (int s = 0; s < 8; s++) cost[c][s] = ((s & 4) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 1) ? d[2] : -d[2]);
        }
        for (int k = 0; k < nc; k++) {
            int bc = -1, bs = -1; float bv = -INFINITY;
            for (int c = 0; c < nc; c++) if (!cUsed[c]) for (int s = 0; s < 8; s++) if (!sUsed[s] && cost[c][s] > bv) { bv = cost[c][s]; bc = c; bs = s; }
            cUsed[bc] = true; sUsed[bs] = true; slotOf[bc] = bs;
        }
        int childAt[8]; for (int s = 0; s < 8; s++) childAt[s] = -1;
        for (int c = 0; c < nc; c++) childAt[slotOf[c]] = ch[c];

        Node8 N; memset(&N, 0, sizeof N);
        N.px = nb.lo[0]; N.py = nb.lo[1]; N.pz = nb.lo[2];
        N.ex = quantExp(nb.hi[0] - nb.lo[0]); N.ey = quantExp(nb.hi[1] - nb.lo[1]); N.ez = quantExp(nb.hi[2] - nb.lo[2]);
        N.childBase = (uint32_t)out.nodes.size(); N.triBase = (uint32_t)out.slots.size();
        const uint8_t ebits[3] = {N.ex, N.ey, N.ez};
        uint8_t* qlo[3] = {N.qlox, N.qloy, N.qloz}; uint8_t* qhi[3] = {N.qhix, N.qhiy, N.qhiz};
        uint32_t triOff = 0;
        for (int s = 0; s < 8; s++) {
            int c2 = childAt[s];
            if (c2 < 0) continue;
            const Node2& c = B.nodes[c2];
            for (int a = 0; a < 3; a++) {
                if (ebits[a] == 0) { qlo[a][s] = 0; qhi[a][s] = 0; continue; }
                double sc = std::ldexp(1.0, (int)ebits[a] - 127), p = (&N.px)[a];
                int lo = (int)std::floor(((double)c.box.lo[a] - p) / sc), hi = (int)std::ceil(((double)c.box.hi[a] - p) / sc);
                lo = std::max(0, std::min(255, lo)); hi = std::max(0, std::min(255, hi));
                // make sure the float decode the kernel performs is conservative
                float scf = (float)sc, pf = (&N.px)[a];
                while (lo > 0 && pf + (float)lo * scf > c.box.lo[a]) lo--;
                while (hi < 255 && pf + (float)hi * scf < c.box.hi[a]) hi++;
                qlo[a][s] = (uint8_t)lo; qhi[a][s] = (uint8_t)hi;
            }
            if (isLeaf2(c)) {
                uint32_t cnt = (uint32_t)c.count;               // <= EL_MAX_LEAF (the layout allows 3)
                N.triMask |= ((1u << cnt) - 1u) << (3 * s);
                for (uint32_t k = 0; k < cnt; k++) {
                    int t = refTri[B.idx[c.first + k]];
                    const ElevenTri& T = tris[t];
                    TriSlot S;
                    S.v0x = T.vertices[0][0]; S.v0y = T.vertices[0][1]; S.v0z = T.vertices[0][2];
                    S.e1x = T.vertices[1][0] - T.vertices[0][0]; S.e1y = T.vertices[1][1] - T.vertices[0][1]; S.e1z = T.vertices[1][2] - T.vertices[0][2];
                    S.e2x = T.vertices[2][0] - T.vertices[0][0]; S.e2y = T.vertices[2][1] - T.vertices[0][1]; S.e2z = T.vertices[2][2] - T.vertices[0][2];
                    S.tri = t; S.material = triMaterial ? triMaterial[t] : 0; S.shiftBound = triShift[t];
                    out.slots.push_back(S);
                }
                triOff += cnt;
            } else {
                N.imask |= (uint8_t)(1u << s);
            }
        }
        // internal children: contiguous, ascending slot order (rank = popcount of imask below the slot)
        for (int s = 0; s < 8; s++) {
            int c2 = childAt[s];
            if (c2 < 0 || isLeaf2(B.nodes[c2])) continue;
            uint32_t id = (uint32_t)out.nodes.size();
            out.nodes.emplace_back(); memset(&out.nodes.back(), 0, sizeof(Node8));
            queue.push_back({c2, id, it.depth + 1});
        }
        out.nodes[it.n8] = N;
        {   // largest shift bound of any triangle below this node: lets the traversal cull with a LOCAL slack
            float ms = 0.f;
            for (int k = root.first; k < root.first + root.span; k++) ms = std::max(ms, triShift[refTri[B.idx[k]]]);
            if (out.nodeSlack.size() <= it.n8) out.nodeSlack.resize(it.n8 + 1, 0.f);
            out.nodeSlack[it.n8] = ms;
            out.nodes[it.n8].slack = ms;
        }
    }
    out.nodeSlack.resize(out.nodes.size(), 0.f);
    out.buildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace eleven
