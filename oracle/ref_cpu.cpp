// oracle/ref_cpu.cpp — TEST INFRASTRUCTURE (oracle O3): the reference's own headers, unchanged
// and included from where they lie under /root/reference/src, compiled for the host through
// oracle/shim.  Builds oracle/_ref/libref_cpu.so.  The reference ships no CPU renderer and its
// classes are __device__-only (ray.h:11-12, camera.h:7,33, sphere.h:8-11); the shim is what
// makes a host build possible, so this is "the reference headers on CPU cores", used
//   (1) to pin the restatement in oracle/rt_oracle.cpp draw for draw and bit for bit, and
//   (2) as the host-core baseline of bench.py (`cpu_baseline.kind = "reference"`).
// Everything that computes is the reference's code; this file only builds the object graph from
// the flat scene (same constructor calls as populate_scene_balls, main.cu:188-356), loops over
// pixels the way render does (main.cu:109-127) and walks color()'s recurrence (main.cu:35-74)
// by calling the reference's hit / emit / scatter.
#include "common.h" // reference (pulls oracle/shim/curand_kernel.h)
#include "vec3.h"
#include "ray.h"
#include "sphere.h"
#include "hitable_list.h"
#include "camera.h"
#include "bvh.h"
#include "texture.h"

#include <atomic>
#include <thread>
#include <vector>

#include "rt_api.h"

namespace {

struct RefScene {
    std::vector<text*> textures;
    std::vector<material*> materials; // one per sphere: the reference's sphere owns its material
    std::vector<hitable_object*> objects;
    std::vector<uint32_t> ids;
    bvh_node* bvh = nullptr;
    hitable_list* list = nullptr;
    camera* cam = nullptr;
    std::vector<std::vector<float>> images;
};

// color() (main.cu:35-74): the recurrence A <- (emit + 0.1) + att * A, walked with the
// reference's own hit / emit / scatter.  `rays` counts scene.hit() queries.
vec3 color_walk(const ray& r, hitable_list* scene, curandState* rstate, unsigned long long* rays) {
    ray cur = r;
    vec3 acc(1.f, .8f, .7f);
    for (int bounce = 0; bounce < RAY_BOUNCES; ++bounce) {
        hit_record rec;
        ++*rays;
        if (!scene->hit(cur, 0.00001f, FLT_MAX, rec)) return acc;
        ray next;
        vec3 att;
        vec3 e = rec.m()->emit(rec) + vec3(0.1, 0.1, 0.1);
        if (!rec.m()->scatter(cur, next, rec, att, rstate)) return e;
        acc = e + att * acc;
        cur = next;
    }
    return vec3();
}

} // namespace

extern "C" {

void* refcpu_scene_create(const rt_scene_desc* d, int use_bvh) {
    RefScene* s = new RefScene();
    for (uint32_t i = 0; i < d->n_images; ++i) {
        const rt_image& im = d->images[i];
        s->images.emplace_back(im.rgb, im.rgb + size_t(im.width) * im.height * 3);
    }
    s->textures.assign(d->n_textures, nullptr);
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const rt_texture& t = d->textures[i];
        vec3 c1(t.color1[0], t.color1[1], t.color1[2]), c2(t.color2[0], t.color2[1], t.color2[2]);
        switch (t.kind) {
        case RT_TEX_CONSTANT: s->textures[i] = new constant_texture(c1); break;
        case RT_TEX_NOISE_PERLIN: s->textures[i] = new noise_texture(noise_type::PERLIN, t.density); break;
        case RT_TEX_NOISE_TURBULANCE: s->textures[i] = new noise_texture(noise_type::TURBULANCE, t.density); break;
        case RT_TEX_NOISE_MARBLE: s->textures[i] = new noise_texture(noise_type::MARBLE, t.density); break;
        case RT_TEX_WOOD: s->textures[i] = new wood_texture(c1, c2, t.density, t.hardness); break;
        case RT_TEX_IMAGE:
            s->textures[i] = new image_texture(s->images[t.image].data(), d->images[t.image].width, d->images[t.image].height);
            break;
        default: break;
        }
    }
    for (uint32_t pass = 0; pass < d->n_textures; ++pass)
        for (uint32_t i = 0; i < d->n_textures; ++i) {
            const rt_texture& t = d->textures[i];
            if (s->textures[i] || t.kind != RT_TEX_CHECKER) continue;
            if (s->textures[t.even] && s->textures[t.odd]) s->textures[i] = new checker_texture(s->textures[t.even], s->textures[t.odd]);
        }
    const uint32_t n = d->n_spheres;
    s->objects.assign(size_t(n) + 1, nullptr);
    for (uint32_t i = 0; i < n; ++i) {
        const rt_sphere& sp = d->spheres[i];
        const rt_material& m = d->materials[sp.material];
        vec3 alb(m.albedo[0], m.albedo[1], m.albedo[2]);
        material* mat = nullptr;
        switch (m.kind) {
        case RT_MAT_LAMBERTIAN: mat = new lambertian(s->textures[m.texture]); break;
        case RT_MAT_METAL: mat = new metal(alb, m.param); break;
        case RT_MAT_DIELECTRIC: mat = new dielectric(m.param, alb); break;
        default: mat = new emitter(s->textures[m.texture], m.param); break;
        }
        s->materials.push_back(mat);
        vec3 c0(sp.center0[0], sp.center0[1], sp.center0[2]), c1(sp.center1[0], sp.center1[1], sp.center1[2]);
        if (sp.flags & RT_SPHERE_MOVING) s->objects[i] = new moving_sphere(c0, c1, sp.time0, sp.time1, sp.radius, mat);
        else s->objects[i] = new sphere(c0, sp.radius, mat, (sp.flags & RT_SPHERE_INSIDE) != 0);
        s->objects[i]->set_id(sp.id);
        s->ids.push_back(sp.id);
    }
    if (use_bvh && n > 0) {
        curandState rs;
        curand_init(SEED, 0, 0, &rs);
        // the ctor sorts the pointer array in place; hand it a copy so ordinals stay stable
        std::vector<hitable_object*>* sorted = new std::vector<hitable_object*>(s->objects.begin(), s->objects.begin() + n);
        s->bvh = new bvh_node(sorted->data(), int(n), d->camera.time0, d->camera.time1, &rs, 0);
    }
    s->list = new hitable_list(s->objects.data(), s->bvh, n);
    const rt_camera& c = d->camera;
    s->cam = new camera(vec3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]), vec3(c.lookat[0], c.lookat[1], c.lookat[2]),
                        vec3(c.up[0], c.up[1], c.up[2]), c.vfov, c.aspect, c.aperture, c.focus_dist, c.time0, c.time1);
    return s;
}

// objects are deliberately not destroyed: the reference's destructors printf and double-free
// shared state (sphere.h:148-155, hitable_list.h:104-123); the process owns the leak.
void refcpu_scene_destroy(void*) {}

void refcpu_trace(void* h, const rt_ray* rays, size_t n, float tmin, rt_hit* hits) {
    RefScene* s = static_cast<RefScene*>(h);
    for (size_t i = 0; i < n; ++i) {
        const rt_ray& in = rays[i];
        ray r(vec3(in.origin[0], in.origin[1], in.origin[2]), vec3(in.direction[0], in.direction[1], in.direction[2]), in.time);
        hit_record rec;
        rt_hit out;
        memset(&out, 0, sizeof out);
        if (s->list->hit(r, tmin, FLT_MAX, rec)) {
            size_t k = 0;
            while (k < s->materials.size() && s->materials[k] != rec.m()) ++k;
            out.t = rec.t();
            out.id = k < s->ids.size() ? s->ids[k] : RT_INVALID_ID;
            out.p[0] = rec.p().x(); out.p[1] = rec.p().y(); out.p[2] = rec.p().z();
            out.n[0] = rec.n().x(); out.n[1] = rec.n().y(); out.n[2] = rec.n().z();
            out.u = rec.u();
            out.v = rec.v();
        } else {
            out.t = FLT_MAX;
            out.id = RT_INVALID_ID;
        }
        hits[i] = out;
    }
}

// render (main.cu:97-132) with runtime sizes.  mean: W*H*3 floats = col / spp before saturate
// and gamma, index j*W+i with j = 0 the bottom row.  fb (may be NULL): the finished pixel.
void refcpu_render(void* h, int width, int height, int spp, unsigned seed, int nthreads, float* mean, float* fb,
                   unsigned long long* rays_out) {
    RefScene* s = static_cast<RefScene*>(h);
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next_row{0};
    std::atomic<unsigned long long> total{0};
    auto work = [&]() {
        unsigned long long rays = 0;
        for (;;) {
            int j = next_row.fetch_add(1);
            if (j >= height) break;
            for (int i = 0; i < width; ++i) {
                size_t index = size_t(j) * size_t(width) + size_t(i);
                curandState rstate;
                curand_init(seed + index, 0, 0, &rstate); // init_rand_state, main.cu:91
                vec3 col;
                for (int k = 0; k < spp; ++k) {
                    float u = float(i + curand_uniform(&rstate)) / float(width);
                    float v = float(j + curand_uniform(&rstate)) / float(height);
                    ray r = s->cam->get_ray(u, v, &rstate);
                    col += color_walk(r, s->list, &rstate, &rays);
                }
                col /= float(spp);
                mean[index * 3 + 0] = col.x();
                mean[index * 3 + 1] = col.y();
                mean[index * 3 + 2] = col.z();
                if (fb) {
                    vec3 out = col.saturate().gamma_correct();
                    fb[index * 3 + 0] = out.x();
                    fb[index * 3 + 1] = out.y();
                    fb[index * 3 + 2] = out.z();
                }
            }
        }
        total += rays;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (rays_out) *rays_out = total.load();
}

// ---- unit hooks for the golden vectors ----
float refcpu_perlin_noise(const float p[3]) {
    static perlin_noise pn;
    return pn.noise(vec3(p[0], p[1], p[2]));
}
float refcpu_turbulence(const float p[3]) {
    static perlin_noise pn;
    return pn.turbulance_noise(vec3(p[0], p[1], p[2]));
}
void refcpu_texture_value(void* h, int tex, float u, float v, const float p[3], float out[3]) {
    RefScene* s = static_cast<RefScene*>(h);
    vec3 c = s->textures[size_t(tex)]->value(u, v, vec3(p[0], p[1], p[2]));
    out[0] = c.x(); out[1] = c.y(); out[2] = c.z();
}
void refcpu_reflect(const float v[3], const float n[3], float out[3]) {
    vec3 r = utils::reflect(vec3(v[0], v[1], v[2]), vec3(n[0], n[1], n[2]));
    out[0] = r.x(); out[1] = r.y(); out[2] = r.z();
}
int refcpu_refract(const float v[3], const float n[3], float mu, float out[3]) {
    vec3 r;
    bool ok = utils::refract(vec3(v[0], v[1], v[2]), vec3(n[0], n[1], n[2]), mu, r);
    out[0] = r.x(); out[1] = r.y(); out[2] = r.z();
    return ok ? 1 : 0;
}
float refcpu_shlick(float cosine, float ri) { return utils::shlick(cosine, ri); }
// camera::get_ray (camera.h:33-38) with a caller-seeded generator
void refcpu_camera_ray(void* h, float s_, float t_, unsigned long long seed, rt_ray* out) {
    RefScene* s = static_cast<RefScene*>(h);
    curandState rs;
    curand_init(seed, 0, 0, &rs);
    ray r = s->cam->get_ray(s_, t_, &rs);
    out->origin[0] = r.origin().x(); out->origin[1] = r.origin().y(); out->origin[2] = r.origin().z();
    out->direction[0] = r.direction().x(); out->direction[1] = r.direction().y(); out->direction[2] = r.direction().z();
    out->time = r.t();
}

} // extern "C"
