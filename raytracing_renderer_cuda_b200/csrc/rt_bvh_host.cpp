// rt_bvh_host.cpp — binned SAH BVH build on the host (small / medium scenes).
#include "rt_bvh_host.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>

namespace rth {

namespace {

struct Builder {
    const std::vector<Box>& boxes;
    std::vector<NodeHost>& nodes;
    std::vector<uint32_t> prims;
    std::vector<float> cen; // 3 per prim
    uint32_t max_depth = 0;
    double sah = 0.0;

    static void grow(Box& b, const Box& o) {
        for (int a = 0; a < 3; ++a) {
            b.lo[a] = std::min(b.lo[a], o.lo[a]);
            b.hi[a] = std::max(b.hi[a], o.hi[a]);
        }
    }
    static Box empty() { return Box{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; }
    static double area(const Box& b) {
        double dx = double(b.hi[0]) - b.lo[0], dy = double(b.hi[1]) - b.lo[1], dz = double(b.hi[2]) - b.lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }

    // builds the subtree over prims[b,e); returns the child reference and its bounds
    int32_t build(uint32_t b, uint32_t e, uint32_t depth, Box& bounds) {
        if (e - b == 1) {
            bounds = boxes[prims[b]];
            max_depth = std::max(max_depth, depth);
            return ~int32_t(prims[b]);
        }
        const int32_t me = int32_t(nodes.size());
        nodes.emplace_back();

        float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t i = b; i < e; ++i)
            for (int a = 0; a < 3; ++a) {
                float c = cen[3 * prims[i] + a];
                cmin[a] = std::min(cmin[a], c);
                cmax[a] = std::max(cmax[a], c);
            }
        constexpr int NB = 16;
        int best_axis = -1, best_bin = -1;
        double best_cost = DBL_MAX;
        const uint32_t n = e - b;
        if (n > 2) {
            for (int a = 0; a < 3; ++a) {
                float ext = cmax[a] - cmin[a];
                if (!(ext > 0.f)) continue;
                Box bb[NB];
                uint32_t cnt[NB] = {0};
                for (int k = 0; k < NB; ++k) bb[k] = empty();
                float scale = float(NB) / ext;
                for (uint32_t i = b; i < e; ++i) {
                    int k = std::min(NB - 1, std::max(0, int((cen[3 * prims[i] + a] - cmin[a]) * scale)));
                    cnt[k]++;
                    grow(bb[k], boxes[prims[i]]);
                }
                double right_area[NB];
                uint32_t right_cnt[NB];
                Box acc = empty();
                uint32_t c = 0;
                for (int k = NB - 1; k > 0; --k) {
                    grow(acc, bb[k]);
                    c += cnt[k];
                    right_area[k] = area(acc);
                    right_cnt[k] = c;
                }
                acc = empty();
                c = 0;
                for (int k = 0; k < NB - 1; ++k) {
                    grow(acc, bb[k]);
                    c += cnt[k];
                    if (c == 0 || right_cnt[k + 1] == 0) continue;
                    double cost = area(acc) * c + right_area[k + 1] * right_cnt[k + 1];
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = a;
                        best_bin = k;
                    }
                }
            }
        }
        uint32_t mid;
        if (best_axis >= 0) {
            float ext = cmax[best_axis] - cmin[best_axis];
            float scale = float(NB) / ext;
            float lo = cmin[best_axis];
            int a = best_axis, kb = best_bin;
            auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](uint32_t p) {
                int k = std::min(NB - 1, std::max(0, int((cen[3 * p + a] - lo) * scale)));
                return k <= kb;
            });
            mid = uint32_t(it - prims.begin());
        } else {
            mid = b; // force the fallback below
        }
        if (mid == b || mid == e) { // degenerate (coincident centroids / n == 2): median split on the widest axis
            int a = 0;
            if (cmax[1] - cmin[1] > cmax[a] - cmin[a]) a = 1;
            if (cmax[2] - cmin[2] > cmax[a] - cmin[a]) a = 2;
            mid = b + n / 2;
            std::nth_element(prims.begin() + b, prims.begin() + mid, prims.begin() + e, [&](uint32_t x, uint32_t y) {
                float cx = cen[3 * x + a], cy = cen[3 * y + a];
                return cx < cy || (cx == cy && x < y);
            });
        }
        Box lb, rb;
        int32_t l = build(b, mid, depth + 1, lb);
        int32_t r = build(mid, e, depth + 1, rb);
        NodeHost& nd = nodes[size_t(me)];
        for (int a = 0; a < 3; ++a) {
            nd.lmin[a] = lb.lo[a];
            nd.lmax[a] = lb.hi[a];
            nd.rmin[a] = rb.lo[a];
            nd.rmax[a] = rb.hi[a];
        }
        nd.left = l;
        nd.right = r;
        nd.pad0 = nd.pad1 = 0.f;
        bounds = lb;
        grow(bounds, rb);
        sah += area(lb) + area(rb);
        return me;
    }
};

} // namespace

void build_bvh_sah(const std::vector<Box>& boxes, std::vector<NodeHost>& nodes, BvhStats* stats) {
    nodes.clear();
    if (boxes.size() < 2) return;
    nodes.reserve(boxes.size() - 1);
    Builder bld{boxes, nodes, {}, {}};
    bld.prims.resize(boxes.size());
    std::iota(bld.prims.begin(), bld.prims.end(), 0u);
    bld.cen.resize(3 * boxes.size());
    for (size_t i = 0; i < boxes.size(); ++i)
        for (int a = 0; a < 3; ++a) bld.cen[3 * i + a] = 0.5f * boxes[i].lo[a] + 0.5f * boxes[i].hi[a];
    Box root;
    bld.build(0, uint32_t(boxes.size()), 0, root);
    if (stats) {
        stats->depth = bld.max_depth;
        double ra = Builder::area(root);
        stats->sah_cost = ra > 0 ? float(bld.sah / ra) : 0.f;
    }
}

Box sphere_box(const float c0[3], const float c1[3], float r) {
    Box b;
    float ar = std::fabs(r);
    for (int a = 0; a < 3; ++a) {
        float lo = std::min(c0[a], c1[a]) - ar;
        float hi = std::max(c0[a], c1[a]) + ar;
        // pad: relative to the radius (rounding of the quadratic at grazing incidence) plus a
        // few ulp of the coordinates themselves
        float pad = ar * 6.1035156e-5f /* 2^-14 */ + (std::fabs(lo) + std::fabs(hi)) * 3.8146973e-6f /* 2^-18 */;
        b.lo[a] = lo - pad;
        b.hi[a] = hi + pad;
    }
    return b;
}

} // namespace rth
