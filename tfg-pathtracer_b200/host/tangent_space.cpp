/*
 * tangent_space.cpp — per-corner tangents + handedness for one mesh object, computed the way the reference does.
 *
 * The reference runs Mikkelsen's tangent-space algorithm ("Simulation of Wrinkled Surfaces Revisited", 2008) over every `o`
 * block of scene.obj through its vendored third-party mikktspace (S/ObjLoader.hpp:167-168, S/mikktspaceCallback.hpp:27-32:
 * genTangSpaceDefault = angular threshold 180 degrees; callbacks: 3 corners per face, SMOOTH_SHADING normals; the result lands
 * in Tri::tangents[corner] and — because every corner overwrites it — Tri::tangentsSign = the sign of corner 2).  The tangents only
 * reach the image through normal maps and the TANGENT / BITANGENT passes (S/Tri.hpp:92,153, S/kernel.cu:90-100), but there they
 * are the difference between the reference's picture and another one, so the product loader follows the published algorithm
 * step by step (this is a re-statement written for this loader: index-based groups, std containers, no callbacks):
 *
 *   1. weld corners whose position, normal and texture coordinate are equal;
 *   2. set degenerate triangles (two equal positions) aside;
 *   3. per triangle: first-order derivatives dP/ds, dP/dt from the UV mapping (normalised, sign = mapping orientation) and their
 *      magnitudes; a triangle with a degenerate mapping may "group with anything";
 *   4. neighbours across edges shared with opposite winding;
 *   5. groups: around every welded vertex, the fan of triangles connected through such edges AND of equal mapping orientation;
 *   6. per group (and per angular sub-group, which at 180 degrees only splits exactly opposite derivatives): the angle-weighted
 *      sum of the members' derivatives projected into the vertex' tangent plane, normalised;
 *   7. degenerate triangles copy the frame of a good triangle that shares the welded vertex.
 *
 * tests/test_host_cli.py pins the result against the reference's own loader on the parity scenes (scene dump of
 * oracle/_ref/ref_host_vectors, fixture tests/golden/tangents_*.npz).
 */
#include "scene_loader.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace eleven_host {
namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 scale(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(V3 v) { return sqrtf(dot(v, v)); }
inline bool nonZero(float x) { return fabsf(x) > FLT_MIN; }
inline bool nonZero(V3 v) { return nonZero(v.x) || nonZero(v.y) || nonZero(v.z); }
inline V3 unit(V3 v) { return scale(1 / length(v), v); }
inline V3 unitIfNonZero(V3 v) { return nonZero(v) ? unit(v) : v; }
inline V3 inPlane(V3 n, V3 v) { return sub(v, scale(dot(n, v), n)); }           // v minus its component along n
inline bool same(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct Frame { V3 os{1.f, 0.f, 0.f}; bool orient = false; };                      // what an unassigned corner keeps

struct TriInfo {
    int face;                       // index into the caller's triangle array
    int v[3];                       // welded vertex ids of the corners
    int neighbour[3] = {-1, -1, -1};
    int group[3] = {-1, -1, -1};
    V3 os{0, 0, 0}, ot{0, 0, 0};
    float magS = 0, magT = 0;
    bool orientPreserving = false, groupWithAny = true;
};
struct Group { int vertex; bool orientPreserving; std::vector<int> tris; };

struct CornerKey {
    uint32_t w[8];
    bool operator==(const CornerKey& o) const { return memcmp(w, o.w, sizeof w) == 0; }
};
struct CornerHash {
    size_t operator()(const CornerKey& k) const { uint64_t h = 1469598103934665603ull; for (uint32_t x : k.w) { h ^= x; h *= 1099511628211ull; } return (size_t)h; }
};

struct Mesh {
    ElevenTri* tris; size_t n;
    V3 pos(int corner) const { const float* p = tris[corner / 3].vertices[corner % 3]; return {p[0], p[1], p[2]}; }
    V3 nrm(int corner) const { const float* p = tris[corner / 3].normals[corner % 3]; return {p[0], p[1], p[2]}; }
    V3 tex(int corner) const { const float* p = tris[corner / 3].uv[corner % 3]; return {p[0], p[1], 1.0f}; }
};

// step 5: the fan of `tri` around the group's vertex joins the group when its mapping orientation agrees
bool joinGroup(std::vector<TriInfo>& T, std::vector<Group>& groups, int tri, int g) {
    TriInfo& me = T[tri];
    Group& G = groups[g];
    int i = me.v[0] == G.vertex ? 0 : me.v[1] == G.vertex ? 1 : 2;
    if (me.group[i] == g) return true;
    if (me.group[i] != -1) return false;
    if (me.groupWithAny && me.group[0] == -1 && me.group[1] == -1 && me.group[2] == -1)
        me.orientPreserving = G.orientPreserving;      // the first group to reach a badly mapped triangle decides its orientation
    if (me.orientPreserving != G.orientPreserving) return false;
    G.tris.push_back(tri);
    me.group[i] = g;
    const int left = me.neighbour[i], right = me.neighbour[i > 0 ? i - 1 : 2];
    if (left >= 0) joinGroup(T, groups, left, g);
    if (right >= 0) joinGroup(T, groups, right, g);
    return true;
}

} // namespace

void computeTangentSpace(ElevenTri* tris, size_t n) {
    if (n == 0) return;
    const Mesh M{tris, n};
    const int nCorners = (int)(3 * n);

    // 1. weld: equal position, normal and texture coordinate (float equality: -0 == +0)
    std::vector<int> vid(nCorners);
    {
        std::unordered_map<CornerKey, int, CornerHash> seen;
        seen.reserve(nCorners);
        for (int c = 0; c < nCorners; c++) {
            const V3 p = M.pos(c), nn = M.nrm(c), t = M.tex(c);
            const float f[8] = {p.x, p.y, p.z, nn.x, nn.y, nn.z, t.x, t.y};
            CornerKey k;
            for (int j = 0; j < 8; j++) { const float v = f[j] == 0.f ? 0.f : f[j]; memcpy(&k.w[j], &v, 4); }
            vid[c] = seen.emplace(k, c).first->second;
        }
    }

    // 2. good triangles first (in order), degenerate ones aside
    std::vector<TriInfo> T; T.reserve(n);
    std::vector<int> degenerate;
    for (size_t f = 0; f < n; f++) {
        const V3 p0 = M.pos(3 * f), p1 = M.pos(3 * f + 1), p2 = M.pos(3 * f + 2);
        if (same(p0, p1) || same(p0, p2) || same(p1, p2)) { degenerate.push_back((int)f); continue; }
        TriInfo t; t.face = (int)f;
        for (int i = 0; i < 3; i++) t.v[i] = vid[3 * f + i];
        T.push_back(t);
    }
    const int nGood = (int)T.size();

    // 3. first-order derivatives of the position with respect to (s, t)
    for (TriInfo& t : T) {
        const int c = 3 * t.face;
        const V3 v1 = M.pos(c), v2 = M.pos(c + 1), v3 = M.pos(c + 2), t1 = M.tex(c), t2 = M.tex(c + 1), t3 = M.tex(c + 2);
        const float t21x = t2.x - t1.x, t21y = t2.y - t1.y, t31x = t3.x - t1.x, t31y = t3.y - t1.y;
        const V3 d1 = sub(v2, v1), d2 = sub(v3, v1);
        const float area2 = t21x * t31y - t21y * t31x;
        const V3 os = sub(scale(t31y, d1), scale(t21y, d2));
        const V3 ot = add(scale(-t31x, d1), scale(t21x, d2));
        t.orientPreserving = area2 > 0;
        if (nonZero(area2)) {
            const float absArea = fabsf(area2), lenOs = length(os), lenOt = length(ot);
            const float sgn = t.orientPreserving ? 1.0f : -1.0f;
            if (nonZero(lenOs)) t.os = scale(sgn / lenOs, os);
            if (nonZero(lenOt)) t.ot = scale(sgn / lenOt, ot);
            t.magS = lenOs / absArea; t.magT = lenOt / absArea;
            if (nonZero(t.magS) && nonZero(t.magT)) t.groupWithAny = false;
        }
    }

    // 4. neighbours: same welded edge, opposite winding, first free partner in triangle order
    {
        struct Edge { int a, b, tri; };
        std::vector<Edge> E; E.reserve(3 * (size_t)nGood);
        for (int f = 0; f < nGood; f++) for (int i = 0; i < 3; i++) {
            const int i0 = T[f].v[i], i1 = T[f].v[i < 2 ? i + 1 : 0];
            E.push_back({std::min(i0, i1), std::max(i0, i1), f});
        }
        std::sort(E.begin(), E.end(), [](const Edge& x, const Edge& y) { return x.a != y.a ? x.a < y.a : x.b != y.b ? x.b < y.b : x.tri < y.tri; });
        // directed edge of triangle `tri` that joins the welded vertices a, b: its number and its (from, to)
        auto directed = [&](int tri, int a, int b, int& from, int& to) -> int {
            const int* v = T[tri].v;
            if (v[0] == a || v[0] == b) {
                if (v[1] == a || v[1] == b) { from = v[0]; to = v[1]; return 0; }
                from = v[2]; to = v[0]; return 2;
            }
            from = v[1]; to = v[2]; return 1;
        };
        for (size_t i = 0; i < E.size(); i++) {
            int fromA, toA; const int eA = directed(E[i].tri, E[i].a, E[i].b, fromA, toA);
            if (T[E[i].tri].neighbour[eA] != -1) continue;
            for (size_t j = i + 1; j < E.size() && E[j].a == E[i].a && E[j].b == E[i].b; j++) {
                int fromB, toB; const int eB = directed(E[j].tri, E[j].a, E[j].b, fromB, toB);
                if (fromA == toB && toA == fromB && T[E[j].tri].neighbour[eB] == -1) {
                    T[E[i].tri].neighbour[eA] = E[j].tri; T[E[j].tri].neighbour[eB] = E[i].tri;
                    break;
                }
            }
        }
    }

    // 5. groups
    std::vector<Group> groups; groups.reserve(3 * (size_t)nGood);
    for (int f = 0; f < nGood; f++) for (int i = 0; i < 3; i++) {
        if (T[f].groupWithAny || T[f].group[i] != -1) continue;
        const int g = (int)groups.size();
        groups.push_back(Group{T[f].v[i], T[f].orientPreserving, {}});
        T[f].group[i] = g; groups[g].tris.push_back(f);
        const int left = T[f].neighbour[i], right = T[f].neighbour[i > 0 ? i - 1 : 2];
        if (left >= 0) joinGroup(T, groups, left, g);
        if (right >= 0) joinGroup(T, groups, right, g);
    }

    // 6. one frame per (sub-)group
    std::vector<Frame> frames(nCorners);
    const float thresholdCos = (float)cos((180.0f * (float)M_PI) / 180.0f);
    auto evaluate = [&](const std::vector<int>& members, int vertex) -> V3 {
        V3 sum{0, 0, 0};
        for (int f : members) {
            const TriInfo& t = T[f];
            if (t.groupWithAny) continue;
            const int i = t.v[0] == vertex ? 0 : t.v[1] == vertex ? 1 : 2;
            const V3 nn = M.nrm(t.v[i]);
            const V3 os = unitIfNonZero(inPlane(nn, t.os));
            const V3 p0 = M.pos(t.v[i > 0 ? i - 1 : 2]), p1 = M.pos(t.v[i]), p2 = M.pos(t.v[i < 2 ? i + 1 : 0]);
            const V3 e1 = unitIfNonZero(inPlane(nn, sub(p0, p1))), e2 = unitIfNonZero(inPlane(nn, sub(p2, p1)));
            float c = dot(e1, e2); c = c > 1 ? 1 : (c < -1 ? -1 : c);
            const float angle = (float)acos(c);
            sum = add(sum, scale(angle, os));
        }
        return unitIfNonZero(sum);
    };
    std::vector<int> members;
    std::vector<std::pair<std::vector<int>, V3>> subGroups;
    for (size_t g = 0; g < groups.size(); g++) {
        const Group& G = groups[g];
        subGroups.clear();
        for (int f : G.tris) {
            const TriInfo& tf = T[f];
            const int i = tf.group[0] == (int)g ? 0 : tf.group[1] == (int)g ? 1 : 2;
            const V3 nn = M.nrm(tf.v[i]);
            const V3 os = unitIfNonZero(inPlane(nn, tf.os)), ot = unitIfNonZero(inPlane(nn, tf.ot));
            members.clear();
            for (int t : G.tris) {
                const TriInfo& tt = T[t];
                const V3 os2 = unitIfNonZero(inPlane(nn, tt.os)), ot2 = unitIfNonZero(inPlane(nn, tt.ot));
                const bool any = tf.groupWithAny || tt.groupWithAny;
                if (any || tf.face == tt.face || (dot(os, os2) > thresholdCos && dot(ot, ot2) > thresholdCos)) members.push_back(t);
            }
            std::sort(members.begin(), members.end());
            size_t s = 0;
            while (s < subGroups.size() && subGroups[s].first != members) s++;
            if (s == subGroups.size()) subGroups.emplace_back(members, evaluate(members, G.vertex));
            Frame& out = frames[3 * (size_t)tf.face + i];
            out.os = subGroups[s].second; out.orient = G.orientPreserving;
        }
    }

    // 7. degenerate triangles borrow the frame of the first good corner with the same welded vertex
    if (!degenerate.empty()) {
        std::unordered_map<int, int> firstGoodCorner;
        for (int j = 0; j < nGood; j++) for (int k = 0; k < 3; k++) firstGoodCorner.emplace(T[j].v[k], 3 * T[j].face + k);
        for (int f : degenerate) for (int i = 0; i < 3; i++) {
            auto it = firstGoodCorner.find(vid[3 * (size_t)f + i]);
            if (it != firstGoodCorner.end()) frames[3 * (size_t)f + i] = frames[it->second];
        }
    }

    for (size_t f = 0; f < n; f++) {
        for (int i = 0; i < 3; i++) { const Frame& fr = frames[3 * f + i]; tris[f].tangents[i][0] = fr.os.x; tris[f].tangents[i][1] = fr.os.y; tris[f].tangents[i][2] = fr.os.z; }
        tris[f].tangentsSign = frames[3 * f + 2].orient ? 1.0f : -1.0f;      // set_tspace_basic overwrites the sign per corner: corner 2 stays
    }
}

} // namespace eleven_host
