// include/rt/scenes.hpp — the benchmark scenes of BASELINE.json, written against the
// façade exactly the way the reference writes populate_scene_balls (main.cu:188-356).
//   C1 earth_emitter  : restatement of main.cu:192-349 (8 primitives, camera :331-349)
//   hdr_sphere        : restatement of the disabled populate_scene_hdr (main.cu:136-182)
//   C2 book1_final    : Shirley book-1 cover scene, ~484 spheres
//   C3 perlin_motion  : every texture kind + moving spheres + emitters, 145 primitives
//   C4 random_spheres : N random spheres + ground (GPU LBVH case)
// Scene RNG: splitmix64, rnd = (next >> 40) * 2^-24; every draw is assigned to a named
// temporary in the documented order (C++ argument evaluation order is unspecified).
#pragma once

#include "scene.hpp"

namespace rt {
namespace scenes {

struct splitmix64 {
    uint64_t s;
    explicit splitmix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    float rnd() { return float(next() >> 40) * (1.0f / 16777216.0f); }
};

struct built {
    std::vector<hitable_object*> objects; // the hitable_object* array handed to hitable_list
    hitable_list* list = nullptr;
    camera* cam = nullptr;
};

// C1 — line-for-line restatement of populate_scene_balls (main.cu:188-356).
// `earth` is the stbi_loadf result for textures/earth.jpg (1200x600x3 floats); the
// reference passes the RENDER size as the texture size (main.cu:237) — both are 1200x600.
inline built earth_emitter(arena& A, const float* earth, int earth_w, int earth_h, uint32_t bvh_mode = RT_BVH_AUTO) {
    built b;
    b.objects.resize(9);
    hitable_object** objects = b.objects.data();

    // sphere 1
    objects[0] = A.obj<sphere>(vec3(0, 0, -1), 0.5f, A.mat<lambertian>(A.tex<constant_texture>(vec3(0.6, 0.1, 0.1))));
    objects[0]->set_id(0);

    // sphere 2: the marble ground
    text* noise1 = A.tex<noise_texture>(noise_type::MARBLE, 1.f);
    objects[1] = A.obj<sphere>(vec3(0, -1000.5, 1), 1000.f, A.mat<lambertian>(noise1));
    objects[1]->set_id(1);

    // sphere 3: earth-textured emitter, intensity 2, inside=true (no effect)
    text* im_text = A.tex<image_texture>(earth, earth_w, earth_h);
    objects[2] = A.obj<sphere>(vec3(1, 0, -1), 0.5f, A.mat<emitter>(im_text, 2.f), true);
    objects[2]->set_id(2);

    // sphere 4: mirror
    objects[3] = A.obj<sphere>(vec3(-1, 0, -2), 0.5f, A.mat<metal>(vec3(1.f), 0.f));
    objects[3]->set_id(3);

    // sphere 5: rough metal
    objects[4] = A.obj<sphere>(vec3(0, 0, -2), 0.5f, A.mat<metal>(vec3(0.8, 0.8, 0.8), 0.5f));
    objects[4]->set_id(4);

    objects[5] = A.obj<sphere>(vec3(1, 0, -2), 0.5f, A.mat<dielectric>(1.5f, vec3(1, 1, 1)));
    objects[5]->set_id(5);

    objects[6] = A.obj<sphere>(vec3(-1, 0, -1), 0.5f, A.mat<emitter>(A.tex<constant_texture>(vec3(0.5, 1, 0.5))));
    objects[6]->set_id(6);

    objects[7] = A.obj<moving_sphere>(vec3(-1, 1, -1), vec3(-2, 1, -1), 0.f, 1.f, 0.2f,
                                      A.mat<lambertian>(A.tex<constant_texture>(vec3(0.6, 0.1, 0.1))));
    objects[7]->set_id(7);

    objects[8] = A.obj<bvh_node>(objects, 8, 0.f, 1.f, nullptr, 0, bvh_mode);
    objects[8]->set_id(8);

    b.list = A.obj<hitable_list>(objects, static_cast<bvh_node*>(objects[8]), 8u);
    b.list->set_id(9);

    vec3 lookfrom = vec3(-1, 1, 5);
    vec3 lookat = vec3(0, 0, -1);
    float dist_to_focus = (lookfrom - lookat).length();
    float aperture = .25f;
    b.cam = A.cam(lookfrom, lookat, vec3(0, 1, 0), 20.f, float(1200) / float(600), aperture, dist_to_focus, 0.f, 0.2f);
    return b;
}

// The reference's second hard-coded scene, populate_scene_hdr (main.cu:136-182; compiled out there by SCENE_BALLS,
// main.cu:17-18, and its textures/hdr.jpg is not shipped): a metal and a lambertian ball inside an r = 10 sphere that
// emits an environment image; brute-force list (bvh == nullptr).  `env` is the stbi_loadf result the reference would
// upload (it passes WIDTH*2 x HEIGHT*2 as the image size, main.cu:147: the dimensions of its hdr.jpg).
inline built hdr_sphere(arena& A, const float* env, int env_w, int env_h) {
    built b;
    b.objects.resize(3);
    hitable_object** objects = b.objects.data();

    objects[0] = A.obj<sphere>(vec3(1., 0, -1), 1.f, A.mat<metal>(vec3(0.8, 0.2, 0.5), 0.05f));
    objects[0]->set_id(0);

    text* hdr_texture = A.tex<image_texture>(env, env_w, env_h);
    // sphere 2
    objects[1] = A.obj<sphere>(vec3(0, 0, 0), 10.f, A.mat<emitter>(hdr_texture));
    objects[1]->set_id(1);

    objects[2] = A.obj<sphere>(vec3(-1., 0, -1), 1.f, A.mat<lambertian>(A.tex<constant_texture>(vec3(0.6, 0.1, 0.1))));
    objects[2]->set_id(2);

    b.list = A.obj<hitable_list>(objects, nullptr, 3u);
    b.list->set_id(3);

    vec3 lookfrom = vec3(-1, 2, 9);
    vec3 lookat = vec3(0, 0, -1);
    float dist_to_focus = (lookfrom - lookat).length();
    float aperture = .25f;
    b.cam = A.cam(lookfrom, lookat, vec3(0, 1, 0), 20.f, float(1200) / float(600), aperture, dist_to_focus, 0.f, 0.2f);
    return b;
}

// C2 — book-1 final scene (SURVEY.md §8d "C2 inputs"), seed 1000
inline built book1_final(arena& A, uint32_t bvh_mode = RT_BVH_AUTO) {
    built b;
    splitmix64 rng(1000);
    std::vector<hitable_object*>& objs = b.objects;
    objs.push_back(A.obj<sphere>(vec3(0, -1000, 0), 1000.f, A.mat<lambertian>(A.tex<constant_texture>(vec3(0.5, 0.5, 0.5)))));
    for (int a = -11; a < 11; ++a) {
        for (int bb = -11; bb < 11; ++bb) {
            float m = rng.rnd();
            float x = rng.rnd();
            float z = rng.rnd();
            vec3 center(a + 0.9f * x, 0.2f, bb + 0.9f * z);
            if ((center - vec3(4, 0.2, 0)).length() > 0.9f) {
                if (m < 0.8f) {
                    float r1 = rng.rnd(), r2 = rng.rnd(), r3 = rng.rnd(), r4 = rng.rnd(), r5 = rng.rnd(), r6 = rng.rnd();
                    objs.push_back(A.obj<sphere>(center, 0.2f,
                                                 A.mat<lambertian>(A.tex<constant_texture>(vec3(r1 * r2, r3 * r4, r5 * r6)))));
                } else if (m < 0.95f) {
                    float r1 = rng.rnd(), r2 = rng.rnd(), r3 = rng.rnd(), r4 = rng.rnd();
                    objs.push_back(A.obj<sphere>(
                        center, 0.2f, A.mat<metal>(vec3(0.5f * (1 + r1), 0.5f * (1 + r2), 0.5f * (1 + r3)), 0.5f * r4)));
                } else {
                    objs.push_back(A.obj<sphere>(center, 0.2f, A.mat<dielectric>(1.5f, vec3(1, 1, 1))));
                }
            }
        }
    }
    objs.push_back(A.obj<sphere>(vec3(0, 1, 0), 1.0f, A.mat<dielectric>(1.5f, vec3(1, 1, 1))));
    objs.push_back(A.obj<sphere>(vec3(-4, 1, 0), 1.0f, A.mat<lambertian>(A.tex<constant_texture>(vec3(0.4, 0.2, 0.1)))));
    objs.push_back(A.obj<sphere>(vec3(4, 1, 0), 1.0f, A.mat<metal>(vec3(0.7, 0.6, 0.5), 0.0f)));
    const uint32_t n = uint32_t(objs.size());
    for (uint32_t i = 0; i < n; ++i) objs[i]->set_id(i);
    objs.push_back(nullptr);
    objs[n] = A.obj<bvh_node>(objs.data(), int(n), 0.f, 0.f, nullptr, 0, bvh_mode);
    objs[n]->set_id(n);
    b.list = A.obj<hitable_list>(objs.data(), static_cast<bvh_node*>(objs[n]), n);
    b.list->set_id(n + 1);
    b.cam = A.cam(vec3(13, 2, 3), vec3(0, 0, 0), vec3(0, 1, 0), 20.f, 16.f / 9.f, 0.1f, 10.f, 0.f, 0.f);
    return b;
}

// C3 — Perlin marble/wood/turbulence + checker textures, moving spheres, emitters; seed 1001
inline built perlin_motion(arena& A, uint32_t bvh_mode = RT_BVH_AUTO) {
    built b;
    splitmix64 rng(1001);
    std::vector<hitable_object*>& objs = b.objects;
    text* white = A.tex<constant_texture>(vec3(0.9, 0.9, 0.9));
    objs.push_back(A.obj<sphere>(vec3(0, -1000, 0), 1000.f,
                                 A.mat<lambertian>(A.tex<checker_texture>(A.tex<noise_texture>(noise_type::MARBLE, 4.f), white))));
    int slot7 = 0;
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) {
            vec3 c(1.2f * i - 6.6f, 0.4f, 1.2f * j - 6.6f);
            int k = (12 * i + j) % 8;
            const float r = 0.4f;
            switch (k) {
            case 0: objs.push_back(A.obj<sphere>(c, r, A.mat<lambertian>(A.tex<noise_texture>(noise_type::PERLIN, 4.f)))); break;
            case 1: objs.push_back(A.obj<sphere>(c, r, A.mat<lambertian>(A.tex<noise_texture>(noise_type::TURBULANCE, 2.f)))); break;
            case 2: objs.push_back(A.obj<sphere>(c, r, A.mat<lambertian>(A.tex<noise_texture>(noise_type::MARBLE, 5.f)))); break;
            case 3:
                objs.push_back(A.obj<sphere>(
                    c, r, A.mat<lambertian>(A.tex<wood_texture>(vec3(0.792, 0.643, 0.447), vec3(0.412, 0.349, 0.306), 10.f))));
                break;
            case 4: {
                float dy = 0.5f * rng.rnd();
                float cr = rng.rnd(), cg = rng.rnd(), cb = rng.rnd();
                text* chk = A.tex<checker_texture>(A.tex<constant_texture>(vec3(cr, cg, cb)), white);
                objs.push_back(A.obj<moving_sphere>(c, c + vec3(0.f, dy, 0.f), 0.f, 1.f, r, A.mat<lambertian>(chk)));
                break;
            }
            case 5: {
                float cr = rng.rnd(), cg = rng.rnd(), cb = rng.rnd(), ro = 0.3f * rng.rnd();
                objs.push_back(A.obj<sphere>(c, r, A.mat<metal>(vec3(cr, cg, cb), ro)));
                break;
            }
            case 6: objs.push_back(A.obj<sphere>(c, r, A.mat<dielectric>(1.5f, vec3(1, 1, 1)))); break;
            default:
                if (slot7++ % 3 == 0) {
                    objs.push_back(A.obj<sphere>(c, r, A.mat<emitter>(A.tex<constant_texture>(vec3(4, 4, 4)))));
                } else {
                    text* chk = A.tex<checker_texture>(A.tex<noise_texture>(noise_type::TURBULANCE, 4.f),
                                                       A.tex<constant_texture>(vec3(0.8, 0.8, 0.8)));
                    objs.push_back(A.obj<sphere>(c, r, A.mat<lambertian>(chk)));
                }
            }
        }
    }
    const uint32_t n = uint32_t(objs.size());
    for (uint32_t i = 0; i < n; ++i) objs[i]->set_id(i);
    objs.push_back(nullptr);
    objs[n] = A.obj<bvh_node>(objs.data(), int(n), 0.f, 1.f, nullptr, 0, bvh_mode);
    objs[n]->set_id(n);
    b.list = A.obj<hitable_list>(objs.data(), static_cast<bvh_node*>(objs[n]), n);
    b.list->set_id(n + 1);
    vec3 lookfrom(10, 4, 10), lookat(0, 0.4, 0);
    b.cam = A.cam(lookfrom, lookat, vec3(0, 1, 0), 30.f, 2.f, 0.1f, (lookfrom - lookat).length(), 0.f, 1.f);
    return b;
}

// C4 — N random spheres + ground; seed 1002.  Materials come from a 512-entry palette
// (70 % lambertian, 20 % metal, 10 % dielectric) so the material table stays small.
inline built random_spheres(arena& A, uint32_t count, uint32_t bvh_mode = RT_BVH_AUTO) {
    built b;
    splitmix64 rng(1002);
    std::vector<material*> palette;
    for (int i = 0; i < 512; ++i) {
        float m = rng.rnd();
        float cr = rng.rnd(), cg = rng.rnd(), cb = rng.rnd(), ro = 0.5f * rng.rnd();
        if (m < 0.7f)
            palette.push_back(A.mat<lambertian>(A.tex<constant_texture>(vec3(cr, cg, cb))));
        else if (m < 0.9f)
            palette.push_back(A.mat<metal>(vec3(cr, cg, cb), ro));
        else
            palette.push_back(A.mat<dielectric>(1.5f, vec3(1, 1, 1)));
    }
    std::vector<hitable_object*>& objs = b.objects;
    objs.reserve(size_t(count) + 2);
    objs.push_back(A.obj<sphere>(vec3(0, -1000, 0), 1000.f, A.mat<lambertian>(A.tex<constant_texture>(vec3(0.5, 0.5, 0.5)))));
    for (uint32_t i = 0; i < count; ++i) {
        float x = -200.f + 400.f * rng.rnd();
        float y = 0.1f + 39.9f * rng.rnd();
        float z = -200.f + 400.f * rng.rnd();
        float r = 0.05f + 0.3f * rng.rnd();
        float pick = rng.rnd();
        float mv = rng.rnd();
        material* mat = palette[size_t(pick * 512.f) & 511u];
        vec3 c0(x, y, z);
        if (mv < 0.05f) {
            float dx = rng.rnd() - 0.5f, dy = rng.rnd() - 0.5f, dz = rng.rnd() - 0.5f;
            objs.push_back(A.obj<moving_sphere>(c0, c0 + vec3(dx, dy, dz), 0.f, 1.f, r, mat));
        } else {
            objs.push_back(A.obj<sphere>(c0, r, mat));
        }
    }
    const uint32_t n = uint32_t(objs.size());
    for (uint32_t i = 0; i < n; ++i) objs[i]->set_id(i);
    objs.push_back(nullptr);
    objs[n] = A.obj<bvh_node>(objs.data(), int(n), 0.f, 1.f, nullptr, 0, bvh_mode);
    objs[n]->set_id(n);
    b.list = A.obj<hitable_list>(objs.data(), static_cast<bvh_node*>(objs[n]), n);
    b.list->set_id(n + 1);
    b.cam = A.cam(vec3(0, 60, -320), vec3(0, 10, 0), vec3(0, 1, 0), 40.f, 16.f / 9.f, 0.f, 10.f, 0.f, 1.f);
    return b;
}

} // namespace scenes
} // namespace rt
