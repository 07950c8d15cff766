// pt.hpp -- host-side mirror of the reference's scene layer, header-only C++17.
//
// A program written against the reference's headers (pt::vec3, pt::sphere,
// pt::reflection_type, pt::camera_config, pt::camera::with_config, pt::scene,
// pt::simple_scene / pt::box_scene, pt::clamp, pt::color_to_int) compiles against
// this header unchanged and hands `scene.spheres.data()` and `&camera` to the C ABI
// of include/ptb200.h.  Same names, same member order, same byte layout (checked by
// the static_asserts below and, value for value, by tests/test_host_scenes.py against
// the reference's own builders).
//
// Mirrors (reference file:line):
//   vec3            src/vec.hpp:7-32, src/vec.cpp:8-69
//   ray             src/ray.hpp:9-15
//   reflection_type src/reflection.hpp:7-12
//   sphere          src/sphere.hpp:10-22     (intersect() is NOT here: that is the GPU's job)
//   camera_config   src/camera.hpp:11-21
//   camera          src/camera.hpp:23-43, with_config src/camera.cpp:3-17
//   scene           src/scene.hpp:12-16
//   scene builders  src/simple_scene.hpp:14-52, src/box_scene.hpp:14-72,
//                   src/box_mirror_scene.hpp:14-72 (named box_mirror_scene here: the
//                   reference gives both box headers the same function name, so only one
//                   of them can be included per translation unit -- SURVEY.md section 5)
//   clamp / color_to_int  src/utils.cpp:6-16
#ifndef PTB200_HOST_PT_HPP
#define PTB200_HOST_PT_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace pt {

inline constexpr double epsilon = 1e-4;
inline constexpr double pi = 3.14159265358979323846;
inline constexpr double inf = 1e20;
inline constexpr int depth_limit = 100;

struct vec3
{
    double x{ 0.0 };
    double y{ 0.0 };
    double z{ 0.0 };

    vec3() noexcept = delete;
    constexpr vec3(double const x_, double const y_, double const z_) noexcept
        : x{ x_ }
        , y{ y_ }
        , z{ z_ }
    {
    }

    [[nodiscard]] constexpr auto operator+(vec3 const& b) const noexcept -> vec3
    {
        return { x + b.x, y + b.y, z + b.z };
    }
    [[nodiscard]] constexpr auto operator-(vec3 const& b) const noexcept -> vec3
    {
        return { x - b.x, y - b.y, z - b.z };
    }
    [[nodiscard]] constexpr auto operator*(double const b) const noexcept -> vec3
    {
        return { x * b, y * b, z * b };
    }
    [[nodiscard]] constexpr auto blend(vec3 const& b) const noexcept -> vec3
    {
        return { x * b.x, y * b.y, z * b.z };
    }
    [[nodiscard]] constexpr auto dot(vec3 const& b) const noexcept -> double
    {
        return x * b.x + y * b.y + z * b.z;
    }
    [[nodiscard]] constexpr auto cross(vec3 const& b) const noexcept -> vec3
    {
        return { y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x };
    }
    // in place, like the reference (vec.cpp:35-38)
    auto norm() noexcept -> vec3&
    {
        double const inv = 1 / std::sqrt(x * x + y * y + z * z);
        x *= inv;
        y *= inv;
        z *= inv;
        return *this;
    }
    [[nodiscard]] auto length() const noexcept -> double
    {
        return std::hypot(x, y, z);
    }
    [[nodiscard]] auto operator[](int const index) const noexcept -> double const&
    {
        return index == 1 ? y : (index == 2 ? z : x);
    }
};

struct ray
{
    vec3 origin{ 0, 0, 0 };
    vec3 direction{ 0, 0, 0 };

    [[nodiscard]] auto at(double const t) const noexcept -> vec3
    {
        return origin + direction * t;
    }
};

enum class reflection_type
{
    diffuse,
    specular,
    dielectric
};

struct sphere
{
    double radius{ 0.0 };
    vec3 position{ 0, 0, 0 };
    vec3 emission{ 0, 0, 0 };
    vec3 color{ 0, 0, 0 };
    reflection_type reflection{ reflection_type::diffuse };
};
static_assert(sizeof(sphere) == 88, "must match the reference's pt::sphere (and PTB_SPHERE_BYTES)");

struct camera_config
{
    vec3 position{ 0, 0, 0 };
    vec3 direction{ 0, 0, 0 }; // a look-at POINT (camera.cpp:8)
    vec3 up{ 0, 1, 0 };
    double aspect_ratio{ 16.0 / 9.0 };
    double vertical_fov_radians{ 0.785398163 };
    double focal_length{ 1.0 };
    double aperture{ 0.0 };
    double focus_distance{ 0.0 };
};
static_assert(sizeof(camera_config) == 112, "must match PTB_CAMERA_CONFIG_BYTES");

struct camera
{
    vec3 position{ 0, 0, 0 };
    vec3 lower_left_corner{ 0, 0, 0 };
    vec3 cam_x_axis{ 0, 0, 0 };
    vec3 cam_y_axis{ 0, 0, 0 };
    vec3 u{ 0, 0, 0 };
    vec3 v{ 0, 0, 0 };
    vec3 w{ 0, 0, 0 };
    double lens_radius{ 0.0 };

    [[nodiscard]] static auto with_config(camera_config const& cfg) noexcept -> camera
    {
        double const view_h = 2.0 * std::tan(0.5 * cfg.vertical_fov_radians);
        double const view_w = cfg.aspect_ratio * view_h;

        vec3 back = cfg.position - cfg.direction;
        back.norm();
        vec3 right = cfg.up.cross(back);
        right.norm();
        vec3 const upward = back.cross(right);

        vec3 const x_axis = right * view_w * cfg.focus_distance;
        vec3 const y_axis = upward * view_h * cfg.focus_distance;
        vec3 const corner = cfg.position - x_axis * 0.5 - y_axis * 0.5 - back * cfg.focus_distance;

        return camera{ cfg.position, corner, x_axis, y_axis, right, upward, back, cfg.aperture / 2.0 };
    }
};
static_assert(sizeof(camera) == 176, "must match the reference's pt::camera (and PTB_CAMERA_BYTES)");

struct scene
{
    std::vector<sphere> spheres{};
    camera_config camera_parameters{};
};

[[nodiscard]] inline auto clamp(double const x) noexcept -> double
{
    return std::clamp(x, 0.0, 1.0);
}

[[nodiscard]] inline auto color_to_int(double const x) noexcept -> int
{
    return static_cast<int>(std::round(std::pow(clamp(x), 1.0 / 2.2) * 255.0));
}

namespace detail {

inline auto aim(scene& s, vec3 const from, vec3 const at, int const w, int const h, double const vfov,
                double const aperture) -> void
{
    s.camera_parameters.position = from;
    s.camera_parameters.direction = at;
    s.camera_parameters.aspect_ratio = (w * 1.0) / (h * 1.0);
    s.camera_parameters.vertical_fov_radians = vfov;
    s.camera_parameters.aperture = aperture;
    s.camera_parameters.focus_distance = (from - at).length();
}

inline constexpr auto D = reflection_type::diffuse;
inline constexpr auto S = reflection_type::specular;
inline constexpr auto G = reflection_type::dielectric;

// the Cornell-style box shared by box_scene / box_mirror_scene: five R = 1e6 wall
// spheres (left, right, back, top, bottom; no front wall), a light, a mirror ball and
// a glass ball.
inline auto box_like(int const w, int const h, reflection_type const wall, vec3 const light_at, vec3 const light_e,
                     vec3 const light_c, double const ball_z, double const vfov) -> scene
{
    constexpr double R = 1E6;
    constexpr double o = 0.4;
    constexpr double z = -1.0;
    vec3 const none{ 0.0, 0.0, 0.0 };
    vec3 const white{ 1.0, 1.0, 1.0 };

    scene s{};
    s.spheres = {
        { R, { -R - o, 0.0, z }, none, { 0.9, 0.1, 0.2 }, wall },
        { R, { R + o, 0.0, z }, none, { 0.3, 0.1, 0.9 }, wall },
        { R, { 0.0, 0.0, z - R }, none, { 0.1, 0.7, 0.2 }, wall },
        { R, { 0.0, R + o, z }, none, { 0.3, 0.7, 0.2 }, wall },
        { R, { 0.0, -R - o, z }, none, { 0.9, 0.9, 0.9 }, wall },
        { o / 2.0, light_at, light_e, light_c, D },
        { o / 2.0, { o / 2.0, -o / 2.0, ball_z }, none, white, S },
        { o / 2.0, { -o / 2.0, -o / 2.0, ball_z }, none, white, G },
    };
    aim(s, { 0.0, 0.0, 2.0 }, { 0.0, 0.0, z + o * 1.5 }, w, h, vfov, 0.2);
    return s;
}

} // namespace detail

// src/simple_scene.hpp:14-52
[[nodiscard]] inline auto simple_scene(int const w, int const h) -> scene
{
    using namespace detail;
    scene s{};
    s.spheres = {
        { 100.0, { 0.0, -100.5, -1.0 }, { 0.0, 0.0, 0.0 }, { 0.8, 0.8, 0.0 }, D },   // ground
        { 0.5, { 1.0, 0.0, -1.0 }, { 0.0, 0.0, 0.0 }, { 0.999, 0.999, 0.999 }, S },  // right, mirror
        { 0.5, { -1.0, 0.0, -1.0 }, { 0.0, 0.0, 0.0 }, { 0.999, 0.999, 0.999 }, G }, // left, glass
        { 0.5, { 0.0, 0.0, -1.0 }, { 0.1, 0.1, 0.9 }, { 0.0, 0.7, 0.1 }, D },        // centre, glowing
        { 1.0, { 1.0, 3.1, -1.0 }, { 30.0, 30.0, 30.0 }, { 0.0, 0.0, 0.0 }, D },     // light
    };
    aim(s, { -2.0, 2.0, 1.0 }, { 0.0, 0.0, -1.0 }, w, h, 1.2, 0.2);
    return s;
}

// src/box_scene.hpp:14-72 (diffuse walls, E = 9 light above the back of the box)
[[nodiscard]] inline auto box_scene(int const w, int const h) -> scene
{
    constexpr double o = 0.4;
    constexpr double z = -1.0;
    return detail::box_like(w, h, detail::D, { 0.0, 0.0 + o / 4.0, z - o / 2.5 }, { 9.0, 9.0, 9.0 }, { 1.8, 1.8, 1.8 },
                            z + o * 1.5, 0.5);
}

// src/box_mirror_scene.hpp:14-72 (mirror walls: the "disco sphere" the reference ships with, main.cpp:25)
[[nodiscard]] inline auto box_mirror_scene(int const w, int const h) -> scene
{
    constexpr double o = 0.4;
    constexpr double z = -1.0;
    return detail::box_like(w, h, detail::S, { 0.0, 0.0 + o / 4.0, z + o * 1.5 }, { 1.92, 1.91, 1.9 },
                            { 1.92, 1.91, 1.9 }, z + o, 0.75);
}

// BASELINE.json config 4: simple_scene geometry with the right
// sphere switched to glass and a wider aperture -- depth of field through two dielectrics.
[[nodiscard]] inline auto dof_glass_scene(int const w, int const h) -> scene
{
    scene s = simple_scene(w, h);
    s.spheres[1].reflection = reflection_type::dielectric;
    s.camera_parameters.aperture = 0.4;
    return s;
}

// BASELINE.json config 5: 10 001 spheres on a ground sphere,
// mixed materials, placed by SplitMix64(0x5EED).
[[nodiscard]] inline auto spheres10k_scene(int const w, int const h) -> scene
{
    std::uint64_t state = 0x5EEDull;
    auto const U = [&state]() noexcept -> double {
        state += 0x9E3779B97F4A7C15ull;
        std::uint64_t zz = state;
        zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ull;
        zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBull;
        zz ^= zz >> 31;
        return static_cast<double>(zz >> 11) * (1.0 / 9007199254740992.0);
    };

    scene s{};
    s.spheres.reserve(10001);
    s.spheres.push_back({ 1000.0, { 0.0, -1000.0, 0.0 }, { 0.0, 0.0, 0.0 }, { 0.5, 0.5, 0.5 }, detail::D });
    for(int a = -50; a < 50; ++a) {
        for(int b = -50; b < 50; ++b) {
            double const m = U();
            double const cx = a + 0.9 * U();
            double const cz = b + 0.9 * U();
            sphere sp{ 0.2, { cx, 0.2, cz }, { 0.0, 0.0, 0.0 }, { 1.0, 1.0, 1.0 }, detail::G };
            if(m < 0.70) {
                double const r1 = U(), r2 = U(), g1 = U(), g2 = U(), b1 = U(), b2 = U();
                sp.color = { r1 * r2, g1 * g2, b1 * b2 };
                sp.reflection = detail::D;
            }
            else if(m < 0.90) {
                double const r = U(), g = U(), bb = U();
                sp.color = { 0.5 + 0.5 * r, 0.5 + 0.5 * g, 0.5 + 0.5 * bb };
                sp.reflection = detail::S;
            }
            if(m < 0.02) {
                sp.emission = { 4.0, 4.0, 4.0 };
            }
            s.spheres.push_back(sp);
        }
    }
    s.camera_parameters.position = { 13.0, 2.0, 3.0 };
    s.camera_parameters.direction = { 0.0, 0.0, 0.0 };
    s.camera_parameters.aspect_ratio = (w * 1.0) / (h * 1.0);
    s.camera_parameters.vertical_fov_radians = 0.35;
    s.camera_parameters.aperture = 0.1;
    s.camera_parameters.focus_distance = 10.0;
    return s;
}

// ---- sandbox/main.cpp (the stand-alone smallpt fork) ------------------------------------------------------
// Its Sphere (sandbox/main.cpp:62-66) has pt::sphere's layout and material order, so the scene is a
// std::vector<pt::sphere>; its camera is four constants in main (sandbox/main.cpp:235,260).
struct smallpt_camera
{
    vec3 position{ 50, 52, 295.6 };
    vec3 direction{ 0, -0.042612, -1 };
    double fov_factor{ .5135 };
    double push{ 140 };
};
static_assert(sizeof(smallpt_camera) == 64, "cam8 of ptb_set_smallpt_camera");

// sandbox/main.cpp:94-122: six R = 1e5 wall spheres seen from inside, two mirrors, a light, a glass ball
[[nodiscard]] inline auto smallpt_scene() -> std::vector<sphere>
{
    using namespace detail;
    vec3 const none{ 0, 0, 0 };
    vec3 const grey{ .75, .75, .75 };
    vec3 const shiny = vec3{ 1, 1, 1 } * .999;
    return {
        { 1e5, { 1e5 + 1, 40.8, 81.6 }, none, { .75, .25, .25 }, D },   // left
        { 1e5, { -1e5 + 99, 40.8, 81.6 }, none, { .25, .25, .75 }, D }, // right
        { 1e5, { 50, 40.8, 1e5 }, none, grey, D },                      // back
        { 1e5, { 50, 40.8, -1e5 + 170 }, none, none, D },               // front
        { 1e5, { 50, 1e5, 81.6 }, none, grey, D },                      // bottom
        { 1e5, { 50, -1e5 + 81.6, 81.6 }, none, { .25, .75, .15 }, D }, // top
        { 16.5, { 27, 16.5, 47 }, none, shiny, S },                     // mirror
        { 16.5, { 65, 16.5, 37 }, none, { 0.6, 0.1, 0.6 }, S },         // purple mirror
        { 16.5, { 45, 46.5, 50 }, { 22, 22, 22 }, none, D },            // light
        { 16.5, { 73, 16.5, 78 }, none, shiny, G },                     // glass
    };
}

// toInt of sandbox/main.cpp:130-133 (rounds by adding .5, unlike pt::color_to_int's std::round)
[[nodiscard]] inline auto smallpt_to_int(double const x) noexcept -> int
{
    double const c = x < 0 ? 0 : x > 1 ? 1 : x;
    return int(std::pow(c, 1 / 2.2) * 255 + .5);
}

} // namespace pt

#endif // PTB200_HOST_PT_HPP
