// Host-side check of the bounding-volume hierarchy builder (cpu-path-tracing_b200/csrc/ptb_bvh.hpp):
//   1. structure: every sphere sits in exactly one leaf, every child box (stored as centre + half width per slab) encloses its spheres;
//   2. queries: a CPU restatement of the device traversal (ptb_path_f32.cuh: bvh_closest_hit, binary32, same
//      operation order) returns exactly the (root, sphere) pair of a linear scan with the same sphere test,
//      for rays from outside, from inside the cloud and from sphere surfaces.
// Built and run by tests/test_bvh_host.py (g++, no GPU).  argv[1] = number of random spheres, argv[2] = rays.
#include "../../cpu-path-tracing_b200/csrc/ptb_bvh.hpp"
#include "../../cpu-path-tracing_b200/host/pt.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

namespace {

using ptb::BvhSphere;
using ptb::BvhTree;

uint32_t bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
float from_bits(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
constexpr uint32_t kNoHit = 0x7F800000u;

struct Ray
{
    float ox, oy, oz, dx, dy, dz, eps;
};

// key_small<kBoth = true>, non-FMA-contracted restatement is NOT needed: both scan and traversal call this very function
uint32_t sphere_key(BvhSphere const& s, Ray const& r)
{
    float const cx = s.cx - r.ox, cy = s.cy - r.oy, cz = s.cz - r.oz;
    float const nb = std::fmaf(cx, r.dx, std::fmaf(cy, r.dy, cz * r.dz));
    float const cc = std::fmaf(cx, cx, std::fmaf(cy, cy, std::fmaf(cz, cz, -s.r * s.r)));
    float const sq = std::sqrt(std::fmaf(nb, nb, -cc)); // NaN when negative: orders above +inf as a key
    float const h = nb - r.eps;
    return std::min(bits(h - sq), bits(h + sq));
}

void scan(std::vector<BvhSphere> const& s, Ray const& r, uint32_t& best, int& id)
{
    best = kNoHit;
    id = -1;
    for(size_t i = 0; i < s.size(); ++i) {
        uint32_t const k = sphere_key(s[i], r);
        if(k < best) {
            best = k;
            id = static_cast<int>(i);
        }
    }
}

long g_nodes = 0, g_tests = 0;

void traverse(std::vector<BvhSphere> const& s, BvhTree const& t, Ray const& r, uint32_t& best, int& id)
{
    best = kNoHit;
    id = -1;
    auto const safe_rcp = [](float d) { return 1.0f / (std::fabs(d) > 1e-20f ? d : std::copysign(1e-20f, d)); };
    float const ix = safe_rcp(r.dx), iy = safe_rcp(r.dy), iz = safe_rcp(r.dz);
    float const nx = -r.ox * ix, ny = -r.oy * iy, nz = -r.oz * iz; // t_mid = mid * (1/d) - o * (1/d), as on the device
    float const jx = std::fabs(ix), jy = std::fabs(iy), jz = std::fabs(iz); // entry / exit = t_mid -+ half * |1/d|
    float tbest = 3.0e38f;
    int stack[64];
    int sp = 0;
    int node = t.root;
    for(;;) {
        if(node >= 0) {
            ++g_nodes;
            ptb::BvhNode64 const& n = t.nodes[static_cast<size_t>(node)];
            float const acx = std::fmaf(n.n0[0], ix, nx), acy = std::fmaf(n.n0[2], iy, ny), acz = std::fmaf(n.n2[0], iz, nz);
            float const bcx = std::fmaf(n.n1[0], ix, nx), bcy = std::fmaf(n.n1[2], iy, ny), bcz = std::fmaf(n.n2[2], iz, nz);
            float const amin = std::fmax(std::fmax(std::fmaf(-n.n0[1], jx, acx), std::fmaf(-n.n0[3], jy, acy)), std::fmax(std::fmaf(-n.n2[1], jz, acz), 0.0f));
            float const amax = std::fmin(std::fmin(std::fmaf(n.n0[1], jx, acx), std::fmaf(n.n0[3], jy, acy)), std::fmin(std::fmaf(n.n2[1], jz, acz), tbest));
            float const bmin = std::fmax(std::fmax(std::fmaf(-n.n1[1], jx, bcx), std::fmaf(-n.n1[3], jy, bcy)), std::fmax(std::fmaf(-n.n2[3], jz, bcz), 0.0f));
            float const bmax = std::fmin(std::fmin(std::fmaf(n.n1[1], jx, bcx), std::fmaf(n.n1[3], jy, bcy)), std::fmin(std::fmaf(n.n2[3], jz, bcz), tbest));
            bool const ha = amin <= amax, hb = bmin <= bmax;
            int const ca = n.child[0], cb = n.child[1];
            if(ha && hb) {
                bool const a_first = amin <= bmin;
                stack[sp++] = a_first ? cb : ca;
                node = a_first ? ca : cb;
                continue;
            }
            if(ha || hb) {
                node = ha ? ca : cb;
                continue;
            }
        }
        else {
            int const code = ~node;
            int const first = code >> 3, count = (code & 7) + 1;
            for(int k = 0; k < count; ++k) {
                ++g_tests;
                int const pos = t.leaf_order[static_cast<size_t>(first + k)];
                uint32_t const key = sphere_key(s[static_cast<size_t>(pos)], r);
                if(key < best || (key == best && pos < id)) {
                    best = key;
                    id = pos;
                    tbest = key < kNoHit ? from_bits(key) + r.eps : tbest;
                }
            }
        }
        if(sp == 0) {
            break;
        }
        node = stack[--sp];
    }
}

int check(std::vector<BvhSphere> const& s, int nrays, unsigned seed, char const* what)
{
    BvhTree const t = ptb::build_bvh(s);
    // 1. structure
    std::vector<int> seen(s.size(), 0);
    for(int i : t.leaf_order) {
        seen[static_cast<size_t>(i)]++;
    }
    for(size_t i = 0; i < s.size(); ++i) {
        if(seen[i] != 1) {
            std::printf("FAIL %s: sphere %zu appears %d times in the leaves\n", what, i, seen[i]);
            return 1;
        }
    }
    // every child's slabs (centre -+ half width, taken exactly) enclose every sphere below that child
    struct Bounds
    {
        double lo[3], hi[3];
    };
    bool enclosed = true;
    auto const below = [&](auto const& self, int code) -> Bounds {
        Bounds b{ { 1e300, 1e300, 1e300 }, { -1e300, -1e300, -1e300 } };
        auto const grow = [&](Bounds const& o) {
            for(int a = 0; a < 3; ++a) {
                b.lo[a] = std::min(b.lo[a], o.lo[a]);
                b.hi[a] = std::max(b.hi[a], o.hi[a]);
            }
        };
        if(code < 0) {
            int const first = (~code) >> 3, count = ((~code) & 7) + 1;
            for(int k = 0; k < count; ++k) {
                BvhSphere const& sp = s[static_cast<size_t>(t.leaf_order[static_cast<size_t>(first + k)])];
                double const c[3] = { sp.cx, sp.cy, sp.cz }, r = std::fabs(static_cast<double>(sp.r));
                grow(Bounds{ { c[0] - r, c[1] - r, c[2] - r }, { c[0] + r, c[1] + r, c[2] + r } });
            }
            return b;
        }
        ptb::BvhNode64 const& n = t.nodes[static_cast<size_t>(code)];
        double const mid[2][3] = { { n.n0[0], n.n0[2], n.n2[0] }, { n.n1[0], n.n1[2], n.n2[2] } };
        double const half[2][3] = { { n.n0[1], n.n0[3], n.n2[1] }, { n.n1[1], n.n1[3], n.n2[3] } };
        for(int c = 0; c < 2; ++c) {
            Bounds const sub = self(self, n.child[c]);
            for(int a = 0; a < 3; ++a) {
                enclosed = enclosed && mid[c][a] - half[c][a] <= sub.lo[a] && sub.hi[a] <= mid[c][a] + half[c][a];
            }
            grow(sub);
        }
        return b;
    };
    below(below, t.root);
    if(!enclosed) {
        std::printf("FAIL %s: a child's slabs do not enclose its spheres\n", what);
        return 1;
    }
    if(t.max_depth > 64) {
        std::printf("FAIL %s: depth %d\n", what, t.max_depth);
        return 1;
    }
    // 2. queries
    std::mt19937 g(seed);
    std::uniform_real_distribution<float> u(-1.0f, 1.0f);
    std::uniform_int_distribution<size_t> pick(0, s.size() - 1);
    long hits = 0;
    g_nodes = g_tests = 0;
    for(int k = 0; k < nrays; ++k) {
        Ray r{};
        float dx, dy, dz, l;
        do {
            dx = u(g);
            dy = u(g);
            dz = u(g);
            l = dx * dx + dy * dy + dz * dz;
        } while(l > 1.0f || l < 1e-4f);
        l = 1.0f / std::sqrt(l);
        r.dx = dx * l;
        r.dy = dy * l;
        r.dz = dz * l;
        if(k % 7 == 0) { // axis-parallel: zero components exercise the inf / NaN handling of the slab test
            r.dx = k % 3 == 0 ? 1.0f : 0.0f;
            r.dy = k % 3 == 1 ? -1.0f : 0.0f;
            r.dz = k % 3 == 2 ? 1.0f : 0.0f;
        }
        BvhSphere const& c = s[pick(g)];
        if(k % 3 == 0) { // from a sphere surface (what a bounce does)
            r.ox = c.cx + c.r * r.dx;
            r.oy = c.cy + c.r * r.dy;
            r.oz = c.cz + c.r * r.dz;
            if(k % 2 == 0) {
                r.dx = -r.dx, r.dy = -r.dy, r.dz = -r.dz;
            }
        }
        else { // from somewhere around a sphere, up to 30 radii away
            float const d = 30.0f * std::fabs(u(g)) * c.r;
            r.ox = c.cx + d * u(g);
            r.oy = c.cy + d * u(g);
            r.oz = c.cz + d * u(g);
        }
        r.eps = 1e-4f;
        uint32_t b0, b1;
        int i0, i1;
        scan(s, r, b0, i0);
        traverse(s, t, r, b1, i1);
        bool const hit0 = b0 < kNoHit, hit1 = b1 < kNoHit;
        if(hit0 != hit1 || (hit0 && (b0 != b1 || i0 != i1))) {
            std::printf("FAIL %s ray %d: scan (%08x, %d) traversal (%08x, %d)\n", what, k, b0, i0, b1, i1);
            return 1;
        }
        hits += hit0;
    }
    std::printf("ok %s: %zu spheres, %zu nodes, depth %d, %d rays (%ld hit), %.1f node visits and %.1f sphere tests per ray\n", what,
                s.size(), t.nodes.size(), t.max_depth, nrays, hits, double(g_nodes) / nrays, double(g_tests) / nrays);
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    int const n = argc > 1 ? std::atoi(argv[1]) : 2000;
    int const nrays = argc > 2 ? std::atoi(argv[2]) : 20000;
    int rc = 0;
    {
        // BASELINE config 5, as the host mirror builds it (small spheres only: the ground stays outside the tree)
        pt::scene const scn = pt::spheres10k_scene(1920, 1080);
        std::vector<BvhSphere> s;
        for(auto const& sp : scn.spheres) {
            if(sp.radius <= 32.0) {
                s.push_back(BvhSphere{ static_cast<float>(sp.position.x), static_cast<float>(sp.position.y),
                                       static_cast<float>(sp.position.z), static_cast<float>(sp.radius) });
            }
        }
        rc |= check(s, nrays, 1, "spheres10k");
    }
    std::mt19937 g(7);
    std::uniform_real_distribution<float> u(0.0f, 1.0f);
    {
        std::vector<BvhSphere> s; // uniform cloud, mixed radii, heavy overlap
        for(int i = 0; i < n; ++i) {
            s.push_back(BvhSphere{ 20.0f * u(g) - 10.0f, 20.0f * u(g) - 10.0f, 20.0f * u(g) - 10.0f, 0.05f + 0.6f * u(g) * u(g) });
        }
        rc |= check(s, nrays, 2, "random cloud");
    }
    {
        std::vector<BvhSphere> s; // coincident centres and duplicates: the splitter cannot separate them
        for(int i = 0; i < 40; ++i) {
            s.push_back(BvhSphere{ 1.0f, 2.0f, 3.0f, 0.1f + 0.01f * static_cast<float>(i % 5) });
        }
        for(int i = 0; i < 9; ++i) {
            s.push_back(BvhSphere{ static_cast<float>(i), 0.0f, 0.0f, 0.5f });
        }
        rc |= check(s, nrays / 4, 3, "coincident centres");
    }
    {
        std::vector<BvhSphere> s{ BvhSphere{ 0.0f, 0.0f, 0.0f, 1.0f } }; // a single leaf as the root
        rc |= check(s, 1000, 4, "single sphere");
        s.push_back(BvhSphere{ 3.0f, 0.0f, 0.0f, 1.0f });
        s.push_back(BvhSphere{ 0.0f, 3.0f, 0.0f, 1.0f });
        s.push_back(BvhSphere{ 0.0f, 0.0f, 3.0f, 1.0f });
        s.push_back(BvhSphere{ 3.0f, 3.0f, 0.0f, 1.0f });
        rc |= check(s, 2000, 5, "five spheres");
    }
    return rc;
}
