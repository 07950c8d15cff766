// ptb_bvh.hpp -- host-side builder of the bounding-volume hierarchy over a scene's ordinary-sized
// spheres (the reference author's first TODO, README.md:8).
//
// The reference's closest-hit query is a linear scan (src/main.cpp:30-42): 10 001 sphere tests per
// ray on BASELINE config 5.  The hierarchy keeps the query's result -- the same sphere test decides
// every hit (ptb_path_f32.cuh: key_small), the closest root wins and equal roots go to the lower
// list position -- and only skips spheres whose (padded) box the ray misses or reaches later than
// the best root so far.
//
// Layout (device, ptb_scene.cuh: GeoLists): binary tree, 64-byte nodes that hold the boxes of BOTH
// children so that one node fetch (4 x 16 bytes) decides where to go:
//   n0 = (c0.mid.x, c0.half.x, c0.mid.y, c0.half.y)    n1 = the same for child 1
//   n2 = (c0.mid.z, c0.half.z, c1.mid.z, c1.half.z)    n3 = (child0, child1, -, -) as int bits
// each slab as (centre, half width) -- the half width rounded up, so the pair still encloses the padded box: entry and
// exit distances are then two FFMAs off the centre's distance, with no per-axis min / max (ptb_path_f32.cuh: bvh_closest_hit).
// child >= 0: index of an inner node; child < 0: leaf, ~child = first * 8 + (count - 1) into the
// leaf-ordered sphere arrays (count <= 8).  Built top-down with a 16-bin surface-area heuristic.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <vector>

namespace ptb {

struct BvhSphere // input: one sphere in the shifted FP32 frame
{
    float cx, cy, cz, r;
};

struct BvhNode64
{
    float n0[4], n1[4], n2[4];
    int32_t child[4];
};
static_assert(sizeof(BvhNode64) == 64, "node layout");

struct BvhTree
{
    std::vector<BvhNode64> nodes; // nodes[0] = root (absent when the whole set fits one leaf)
    std::vector<int> leaf_order;  // leaf slot -> index of the input sphere
    int32_t root = 0;             // child code of the root (>= 0 inner node 0, < 0 a single leaf)
    int max_depth = 0;
};

namespace bvh_detail {

constexpr int kLeafMax = 1;  // split until a node holds at most this many spheres (measured on config 5: 1 -> 112.9 ms, 2 -> 114.0,
                             // 4 -> 117.1, 8 -> 117.0: leaf tests run at ~6 of 32 lanes, box tests at ~21) ...
constexpr int kLeafHard = 8; // ... unless no split separates them (coincident centres): then up to this many per leaf
constexpr int kBins = 16;

struct Box
{
    float lo[3] = { std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max() };
    float hi[3] = { -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max() };
    void grow(Box const& b)
    {
        for(int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    [[nodiscard]] double area() const
    {
        double const dx = static_cast<double>(hi[0]) - lo[0], dy = static_cast<double>(hi[1]) - lo[1], dz = static_cast<double>(hi[2]) - lo[2];
        return dx < 0 ? 0.0 : 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

// Box of one sphere, padded: the FP32 sphere test may report a root whose point lies a few ulp outside
// the exact sphere, and the slab test rounds too; 1e-4 relative + 1e-5 absolute is orders above both
// and costs nothing measurable in culling.
inline Box sphere_box(BvhSphere const& s)
{
    Box b;
    float const c[3] = { s.cx, s.cy, s.cz };
    for(int a = 0; a < 3; ++a) {
        float const pad = s.r * 1.0001f + 1e-5f + 1e-6f * std::fabs(c[a]);
        b.lo[a] = c[a] - pad;
        b.hi[a] = c[a] + pad;
    }
    return b;
}

struct Builder
{
    std::vector<BvhSphere> const& sph;
    std::vector<Box> boxes;
    std::vector<int> idx; // permutation being partitioned
    BvhTree tree;
    int leaf_max = kLeafMax;

    explicit Builder(std::vector<BvhSphere> const& s) : sph(s)
    {
        boxes.reserve(s.size());
        for(auto const& x : s) {
            boxes.push_back(sphere_box(x));
        }
        idx.resize(s.size());
        for(size_t i = 0; i < s.size(); ++i) {
            idx[i] = static_cast<int>(i);
        }
    }

    Box range_box(int b, int e) const
    {
        Box r;
        for(int i = b; i < e; ++i) {
            r.grow(boxes[static_cast<size_t>(idx[static_cast<size_t>(i)])]);
        }
        return r;
    }

    int32_t make_leaf(int b, int e)
    {
        int const first = static_cast<int>(tree.leaf_order.size());
        for(int i = b; i < e; ++i) {
            tree.leaf_order.push_back(idx[static_cast<size_t>(i)]);
        }
        return ~static_cast<int32_t>(first * 8 + (e - b - 1));
    }

    // Returns the child code of the subtree over idx[b, e); `box` is its bounding box.
    int32_t build(int b, int e, Box const& box, int depth)
    {
        tree.max_depth = std::max(tree.max_depth, depth);
        int const n = e - b;
        if(n <= leaf_max) {
            return make_leaf(b, e);
        }
        // binned SAH over the centroids
        float clo[3], chi[3];
        for(int a = 0; a < 3; ++a) {
            clo[a] = std::numeric_limits<float>::max();
            chi[a] = -std::numeric_limits<float>::max();
        }
        for(int i = b; i < e; ++i) {
            BvhSphere const& s = sph[static_cast<size_t>(idx[static_cast<size_t>(i)])];
            float const c[3] = { s.cx, s.cy, s.cz };
            for(int a = 0; a < 3; ++a) {
                clo[a] = std::min(clo[a], c[a]);
                chi[a] = std::max(chi[a], c[a]);
            }
        }
        double best_cost = std::numeric_limits<double>::max();
        int best_axis = -1, best_bin = -1;
        for(int a = 0; a < 3; ++a) {
            float const ext = chi[a] - clo[a];
            if(!(ext > 0.0f)) {
                continue;
            }
            std::array<Box, kBins> bb;
            std::array<int, kBins> cnt{};
            float const scale = static_cast<float>(kBins) / ext;
            for(int i = b; i < e; ++i) {
                int const id = idx[static_cast<size_t>(i)];
                BvhSphere const& s = sph[static_cast<size_t>(id)];
                float const c = a == 0 ? s.cx : (a == 1 ? s.cy : s.cz);
                int const k = std::min(kBins - 1, static_cast<int>((c - clo[a]) * scale));
                bb[static_cast<size_t>(k)].grow(boxes[static_cast<size_t>(id)]);
                cnt[static_cast<size_t>(k)]++;
            }
            std::array<double, kBins> right_area{};
            std::array<int, kBins> right_cnt{};
            Box acc;
            int c = 0;
            for(int k = kBins - 1; k > 0; --k) {
                acc.grow(bb[static_cast<size_t>(k)]);
                c += cnt[static_cast<size_t>(k)];
                right_area[static_cast<size_t>(k)] = acc.area();
                right_cnt[static_cast<size_t>(k)] = c;
            }
            Box left;
            int lc = 0;
            for(int k = 0; k < kBins - 1; ++k) {
                left.grow(bb[static_cast<size_t>(k)]);
                lc += cnt[static_cast<size_t>(k)];
                int const rc = right_cnt[static_cast<size_t>(k + 1)];
                if(lc == 0 || rc == 0) {
                    continue;
                }
                double const cost = left.area() * lc + right_area[static_cast<size_t>(k + 1)] * rc;
                if(cost < best_cost) {
                    best_cost = cost;
                    best_axis = a;
                    best_bin = k;
                }
            }
        }
        int mid = -1;
        if(best_axis >= 0) {
            float const ext = chi[best_axis] - clo[best_axis];
            float const scale = static_cast<float>(kBins) / ext;
            auto const it = std::partition(idx.begin() + b, idx.begin() + e, [&](int id) {
                BvhSphere const& s = sph[static_cast<size_t>(id)];
                float const c = best_axis == 0 ? s.cx : (best_axis == 1 ? s.cy : s.cz);
                return std::min(kBins - 1, static_cast<int>((c - clo[best_axis]) * scale)) <= best_bin;
            });
            mid = static_cast<int>(it - idx.begin());
        }
        if(mid <= b || mid >= e) {
            // all centroids coincide (or the bins could not separate them)
            if(n <= kLeafHard) {
                return make_leaf(b, e);
            }
            mid = b + n / 2; // arbitrary halves: still correct, boxes overlap
        }
        Box const lb = range_box(b, mid), rb = range_box(mid, e);
        int32_t const me = static_cast<int32_t>(tree.nodes.size());
        tree.nodes.emplace_back();
        int32_t const c0 = build(b, mid, lb, depth + 1);
        int32_t const c1 = build(mid, e, rb, depth + 1);
        BvhNode64& nd = tree.nodes[static_cast<size_t>(me)];
        // each slab as (centre, half width), the half width rounded UP so that the float pair still encloses [lo, hi]
        auto const slab = [](float lo, float hi, float& c, float& h) {
            c = static_cast<float>(0.5 * (static_cast<double>(lo) + static_cast<double>(hi)));
            double const need = std::max(static_cast<double>(hi) - static_cast<double>(c), static_cast<double>(c) - static_cast<double>(lo));
            h = static_cast<float>(need);
            if(static_cast<double>(h) < need) {
                h = std::nextafter(h, std::numeric_limits<float>::infinity());
            }
        };
        slab(lb.lo[0], lb.hi[0], nd.n0[0], nd.n0[1]);
        slab(lb.lo[1], lb.hi[1], nd.n0[2], nd.n0[3]);
        slab(rb.lo[0], rb.hi[0], nd.n1[0], nd.n1[1]);
        slab(rb.lo[1], rb.hi[1], nd.n1[2], nd.n1[3]);
        slab(lb.lo[2], lb.hi[2], nd.n2[0], nd.n2[1]);
        slab(rb.lo[2], rb.hi[2], nd.n2[2], nd.n2[3]);
        nd.child[0] = c0;
        nd.child[1] = c1;
        nd.child[2] = nd.child[3] = 0;
        (void)box;
        return me;
    }
};

} // namespace bvh_detail

inline BvhTree build_bvh(std::vector<BvhSphere> const& spheres)
{
    bvh_detail::Builder b(spheres);
    if(char const* e = std::getenv("PTB_BVH_LEAF")) { // experiments (dev/): spheres per leaf, 1..8
        int const v = std::atoi(e);
        b.leaf_max = v >= 1 && v <= bvh_detail::kLeafHard ? v : b.leaf_max;
    }
    if(spheres.empty()) {
        b.tree.root = ~0; // empty leaf code is never traversed: callers test the sphere count first
        return std::move(b.tree);
    }
    b.tree.root = b.build(0, static_cast<int>(spheres.size()), b.range_box(0, static_cast<int>(spheres.size())), 1);
    return std::move(b.tree);
}

} // namespace ptb
