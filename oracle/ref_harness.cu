// oracle/ref_harness.cu — TEST INFRASTRUCTURE (oracle O2).  Not part of the
// product: nothing under raytracing_renderer_cuda_b200/ links, loads or calls this.
//
// Runs the REFERENCE's own device code on scenes other than its hard-coded one.  The whole
// reference translation unit is pulled in unchanged, from where it lies under
// /root/reference/src (the Makefile passes -I to it; no reference source is copied into this
// repo), with only its `main` renamed so this file can supply its own.  Every arithmetic
// operation of hit / scatter / emit / value / get_ray / color is therefore the reference's.
// What this file adds is plumbing:
//   * populate_from_pod : builds the reference's device object graph (its constructors,
//                         device `new`) from the flat scene file the product also reads
//                         (rt_scene_desc_save, include/rt_api.h)
//   * trace             : (*scene)->hit(r, tmin, FLT_MAX, rec) for caller-supplied rays
//                         (hitable_list.h:60-79) -> id, t, p, n, u, v per ray
//   * render            : per-pixel sample loop around the reference's color() and
//                         camera::get_ray with runtime width/height/spp and a 64-bit-safe
//                         pixel index (the reference bakes 1200x600x100 in: common.h:13-14,
//                         main.cu:15, utils.h:8-15); also an instrumented pass that counts rays
//   * earth             : decodes a JPEG with the reference's vendored stb exactly as
//                         main.cu:376-380 does and dumps the floats (RGB, row 0 = top)
//
// Build: oracle/Makefile -> oracle/_ref/ref_harness (git-ignored, shipped to the GPU box).
#define main reference_main_not_used
#include "main.cu" // resolved through -I/root/reference/src
#undef main

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../include/rt_api.h"

#define HCHECK(x)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (x);                                                                     \
        if (e__ != cudaSuccess) {                                                                  \
            fprintf(stderr, "ref_harness: %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
            exit(3);                                                                               \
        }                                                                                          \
    } while (0)

#define MAT_STRIDE 32 // >= sizeof of every reference material class (16..24 B)

struct DevScene {
    hitable_object** objects; // n + 1 (bvh root last)
    hitable_list** scene;
    camera** cam;
    unsigned char* matbuf; // n * MAT_STRIDE, material i belongs to sphere i
    uint32_t n;
};

// One thread, like populate_scene_balls<<<1,1>>> (main.cu:188-356, :423).
__global__ void populate_from_pod(const rt_sphere* sph, uint32_t n, const rt_material* mats, const rt_texture* texs,
                                  uint32_t n_tex, float** images, const int32_t* image_wh, rt_camera c, int use_bvh,
                                  DevScene ds, text** tex_out, int* status) {
    if (threadIdx.x || blockIdx.x) return;
    // textures: leaves first, then checkers once both children exist (children may be shared)
    for (uint32_t i = 0; i < n_tex; ++i) tex_out[i] = nullptr;
    for (uint32_t i = 0; i < n_tex; ++i) {
        const rt_texture& t = texs[i];
        vec3 c1(t.color1[0], t.color1[1], t.color1[2]), c2(t.color2[0], t.color2[1], t.color2[2]);
        switch (t.kind) {
        case RT_TEX_CONSTANT: tex_out[i] = new constant_texture(c1); break;
        case RT_TEX_NOISE_PERLIN: tex_out[i] = new noise_texture(noise_type::PERLIN, t.density); break;
        case RT_TEX_NOISE_TURBULANCE: tex_out[i] = new noise_texture(noise_type::TURBULANCE, t.density); break;
        case RT_TEX_NOISE_MARBLE: tex_out[i] = new noise_texture(noise_type::MARBLE, t.density); break;
        case RT_TEX_WOOD: tex_out[i] = new wood_texture(c1, c2, t.density, t.hardness); break;
        case RT_TEX_IMAGE:
            tex_out[i] = new image_texture(images[t.image], image_wh[2 * t.image], image_wh[2 * t.image + 1]);
            break;
        default: break;
        }
    }
    for (uint32_t pass = 0; pass < n_tex; ++pass) {
        bool pending = false;
        for (uint32_t i = 0; i < n_tex; ++i) {
            if (tex_out[i] || texs[i].kind != RT_TEX_CHECKER) continue;
            text* e = tex_out[texs[i].even];
            text* o = tex_out[texs[i].odd];
            if (e && o) tex_out[i] = new checker_texture(e, o);
            else pending = true;
        }
        if (!pending) break;
    }
    for (uint32_t i = 0; i < n_tex; ++i)
        if (!tex_out[i]) { *status = 1; return; }

    for (uint32_t i = 0; i < n; ++i) {
        const rt_sphere& s = sph[i];
        const rt_material& m = mats[s.material];
        void* slot = ds.matbuf + size_t(i) * MAT_STRIDE;
        material* mat = nullptr;
        vec3 alb(m.albedo[0], m.albedo[1], m.albedo[2]);
        switch (m.kind) {
        case RT_MAT_LAMBERTIAN: mat = new (slot) lambertian(tex_out[m.texture]); break;
        case RT_MAT_METAL: mat = new (slot) metal(alb, m.param); break;
        case RT_MAT_DIELECTRIC: mat = new (slot) dielectric(m.param, alb); break;
        case RT_MAT_EMITTER: mat = new (slot) emitter(tex_out[m.texture], m.param); break;
        default: *status = 2; return;
        }
        vec3 c0(s.center0[0], s.center0[1], s.center0[2]), c1(s.center1[0], s.center1[1], s.center1[2]);
        if (s.flags & RT_SPHERE_MOVING) ds.objects[i] = new moving_sphere(c0, c1, s.time0, s.time1, s.radius, mat);
        else ds.objects[i] = new sphere(c0, s.radius, mat, (s.flags & RT_SPHERE_INSIDE) != 0);
        if (!ds.objects[i]) { *status = 3; return; }
        ds.objects[i]->set_id(s.id);
    }
    bvh_node* bvh = nullptr;
    if (use_bvh) {
        curandState rs;
        curand_init(SEED, 0, 0, &rs); // the reference passes an uninitialised state here (main.cu:317 vs :438)
        bvh = new bvh_node(ds.objects, int(n), c.time0, c.time1, &rs, 0);
        if (!bvh) { *status = 4; return; }
        ds.objects[n] = bvh;
        bvh->set_id(n);
    }
    *ds.scene = new hitable_list(ds.objects, bvh, n);
    (*ds.scene)->set_id(n + 1);
    *ds.cam = new camera(vec3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]), vec3(c.lookat[0], c.lookat[1], c.lookat[2]),
                         vec3(c.up[0], c.up[1], c.up[2]), c.vfov, c.aspect, c.aperture, c.focus_dist, c.time0, c.time1);
    *status = (*ds.scene && *ds.cam) ? 0 : 5;
}

__global__ void trace_kernel(const rt_ray* rays, size_t n, float tmin, DevScene ds, const rt_sphere* sph, rt_hit* out) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rt_ray in = rays[i];
    ray r(vec3(in.origin[0], in.origin[1], in.origin[2]), vec3(in.direction[0], in.direction[1], in.direction[2]), in.time);
    hit_record rec;
    rt_hit h;
    memset(&h, 0, sizeof h);
    if ((*ds.scene)->hit(r, tmin, FLT_MAX, rec)) {
        size_t ordinal = size_t(reinterpret_cast<const unsigned char*>(rec.m()) - ds.matbuf) / MAT_STRIDE;
        h.t = rec.t();
        h.id = sph[ordinal].id;
        h.p[0] = rec.p().x(); h.p[1] = rec.p().y(); h.p[2] = rec.p().z();
        h.n[0] = rec.n().x(); h.n[1] = rec.n().y(); h.n[2] = rec.n().z();
        h.u = rec.u();
        h.v = rec.v();
    } else {
        h.t = FLT_MAX;
        h.id = RT_INVALID_ID;
    }
    out[i] = h;
}

// Same per-thread RNG as init_rand_state (main.cu:91) but with an exact integer index.
__global__ void init_rand_state_rt(curandState* st, int width, int height) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    int j = threadIdx.y + blockIdx.y * blockDim.y;
    if (i >= width || j >= height) return;
    size_t index = size_t(j) * size_t(width) + size_t(i);
    curand_init(SEED + index, 0, 0, &st[index]);
}

// The sample loop of render (main.cu:109-127) around the reference's own color()/get_ray with
// runtime sizes.  `sum` receives the per-pixel mean BEFORE saturate/gamma; `fb` the finished
// pixel exactly as the reference stores it.
__global__ void render_rt(vec3* fb, vec3* mean, int width, int height, int spp, hitable_list** scene, camera** cam,
                          curandState* st) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    int j = threadIdx.y + blockIdx.y * blockDim.y;
    if (i >= width || j >= height) return;
    size_t index = size_t(j) * size_t(width) + size_t(i);
    curandState rstate = st[index];
    vec3 col;
    for (int s = 0; s < spp; ++s) {
        float u = float(i + curand_uniform(&rstate)) / float(width);
        float v = float(j + curand_uniform(&rstate)) / float(height);
        ray r = (*cam)->get_ray(u, v, &rstate);
        col += color(r, scene, &rstate);
    }
    col /= float(spp);
    if (mean) mean[index] = col;
    fb[index] = col.saturate().gamma_correct();
}

// Instrumented pass: same RNG stream, same calls, counts scene.hit() queries.  The bounce loop
// has the semantics of color() (main.cu:42-70) and exists only to count; its colours are unused.
__global__ void count_rays_rt(unsigned long long* total, int width, int height, int spp, hitable_list** scene,
                              camera** cam, curandState* st) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    int j = threadIdx.y + blockIdx.y * blockDim.y;
    unsigned long long mine = 0;
    if (i < width && j < height) {
        size_t index = size_t(j) * size_t(width) + size_t(i);
        curandState rstate = st[index];
        for (int s = 0; s < spp; ++s) {
            float u = float(i + curand_uniform(&rstate)) / float(width);
            float v = float(j + curand_uniform(&rstate)) / float(height);
            ray cur = (*cam)->get_ray(u, v, &rstate);
            for (int b = 0; b < RAY_BOUNCES; ++b) {
                hit_record rec;
                ++mine;
                if (!(*scene)->hit(cur, 0.00001f, FLT_MAX, rec)) break;
                ray next;
                vec3 att;
                if (!rec.m()->scatter(cur, next, rec, att, &rstate)) break;
                cur = next;
            }
        }
    }
    atomicAdd(total, mine);
}

// ------------------------------------------------------------------ host side ----
struct HostScene {
    uint32_t hdr[7];
    rt_camera cam;
    std::vector<rt_sphere> spheres;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    std::vector<std::vector<float>> images;
    std::vector<int32_t> image_wh;
};

static bool load_scene(const char* path, HostScene& s) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    bool ok = fread(s.hdr, sizeof s.hdr, 1, f) == 1 && s.hdr[0] == 0x43535452u && s.hdr[1] == 1u;
    ok = ok && fread(&s.cam, sizeof s.cam, 1, f) == 1;
    if (ok) {
        s.spheres.resize(s.hdr[2]);
        s.materials.resize(s.hdr[3]);
        s.textures.resize(s.hdr[4]);
        ok = (!s.hdr[2] || fread(s.spheres.data(), sizeof(rt_sphere), s.hdr[2], f) == s.hdr[2]) &&
             (!s.hdr[3] || fread(s.materials.data(), sizeof(rt_material), s.hdr[3], f) == s.hdr[3]) &&
             (!s.hdr[4] || fread(s.textures.data(), sizeof(rt_texture), s.hdr[4], f) == s.hdr[4]);
    }
    for (uint32_t i = 0; ok && i < s.hdr[5]; ++i) {
        int32_t wh[2];
        ok = fread(wh, sizeof wh, 1, f) == 1;
        if (!ok) break;
        s.image_wh.push_back(wh[0]);
        s.image_wh.push_back(wh[1]);
        s.images.emplace_back(size_t(wh[0]) * wh[1] * 3);
        ok = fread(s.images.back().data(), sizeof(float), s.images.back().size(), f) == s.images.back().size();
    }
    fclose(f);
    return ok;
}

template <class T>
static T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    HCHECK(cudaMalloc(&d, (v.size() ? v.size() : 1) * sizeof(T)));
    if (!v.empty()) HCHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

struct Built {
    DevScene ds;
    rt_sphere* d_sph;
    float ms_build;
};

static Built build_scene(const HostScene& hs, int use_bvh) {
    const uint32_t n = uint32_t(hs.spheres.size());
    // device heap for `new`: objects (<=56 B), noise textures (~800 B), bvh nodes (56 B)
    size_t heap = (size_t(64) << 20) + size_t(n) * 256 + hs.textures.size() * 1024;
    HCHECK(cudaDeviceSetLimit(cudaLimitMallocHeapSize, heap));
    HCHECK(cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024)); // recursive bvh_node ctor + in-thread thrust::sort
    Built b{};
    b.d_sph = upload(hs.spheres);
    rt_material* d_mat = upload(hs.materials);
    rt_texture* d_tex = upload(hs.textures);
    std::vector<float*> img_ptrs;
    for (auto& im : hs.images) img_ptrs.push_back(upload(im));
    float** d_imgs = upload(img_ptrs);
    int32_t* d_wh = upload(hs.image_wh);
    text** d_texout = nullptr;
    HCHECK(cudaMalloc(&d_texout, (hs.textures.size() + 1) * sizeof(text*)));
    int* d_status = nullptr;
    HCHECK(cudaMalloc(&d_status, sizeof(int)));
    HCHECK(cudaMemset(d_status, 0xff, sizeof(int)));
    b.ds.n = n;
    HCHECK(cudaMalloc(&b.ds.objects, (size_t(n) + 1) * sizeof(hitable_object*)));
    HCHECK(cudaMalloc(&b.ds.scene, sizeof(hitable_list*)));
    HCHECK(cudaMalloc(&b.ds.cam, sizeof(camera*)));
    HCHECK(cudaMalloc(&b.ds.matbuf, (size_t(n) + 1) * MAT_STRIDE));
    cudaEvent_t e0, e1;
    HCHECK(cudaEventCreate(&e0));
    HCHECK(cudaEventCreate(&e1));
    HCHECK(cudaEventRecord(e0));
    populate_from_pod<<<1, 1>>>(b.d_sph, n, d_mat, d_tex, uint32_t(hs.textures.size()), d_imgs, d_wh, hs.cam, use_bvh, b.ds,
                                d_texout, d_status);
    HCHECK(cudaEventRecord(e1));
    HCHECK(cudaGetLastError());
    HCHECK(cudaDeviceSynchronize());
    HCHECK(cudaEventElapsedTime(&b.ms_build, e0, e1));
    int status = -1;
    HCHECK(cudaMemcpy(&status, d_status, sizeof status, cudaMemcpyDeviceToHost));
    if (status != 0) {
        fprintf(stderr, "ref_harness: populate_from_pod failed with status %d\n", status);
        exit(4);
    }
    HCHECK(cudaDeviceSetLimit(cudaLimitStackSize, 4 * 1024));
    return b;
}

static bool write_file(const char* path, const void* p, size_t bytes) {
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    bool ok = fwrite(p, 1, bytes, f) == bytes;
    return (fclose(f) == 0) && ok;
}

static int cmd_trace(int argc, char** argv) {
    if (argc < 5) return 2;
    HostScene hs;
    if (!load_scene(argv[2], hs)) { fprintf(stderr, "cannot read scene %s\n", argv[2]); return 2; }
    int use_bvh = argc > 5 ? atoi(argv[5]) : 1;
    float tmin = argc > 6 ? float(atof(argv[6])) : 0.00001f;
    FILE* f = fopen(argv[3], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    size_t bytes = size_t(ftell(f));
    fseek(f, 0, SEEK_SET);
    std::vector<rt_ray> rays(bytes / sizeof(rt_ray));
    if (fread(rays.data(), sizeof(rt_ray), rays.size(), f) != rays.size()) return 2;
    fclose(f);
    Built b = build_scene(hs, use_bvh);
    rt_ray* d_rays = upload(rays);
    rt_hit* d_hits = nullptr;
    HCHECK(cudaMalloc(&d_hits, (rays.size() + 1) * sizeof(rt_hit)));
    if (!rays.empty()) trace_kernel<<<unsigned((rays.size() + 127) / 128), 128>>>(d_rays, rays.size(), tmin, b.ds, b.d_sph, d_hits);
    HCHECK(cudaGetLastError());
    HCHECK(cudaDeviceSynchronize());
    std::vector<rt_hit> hits(rays.size());
    if (!hits.empty()) HCHECK(cudaMemcpy(hits.data(), d_hits, hits.size() * sizeof(rt_hit), cudaMemcpyDeviceToHost));
    if (!write_file(argv[4], hits.data(), hits.size() * sizeof(rt_hit))) return 2;
    printf("{\"cmd\": \"trace\", \"rays\": %zu, \"use_bvh\": %d, \"ms_build\": %.3f}\n", rays.size(), use_bvh, b.ms_build);
    return 0;
}

static int cmd_render(int argc, char** argv) {
    if (argc < 7) return 2;
    HostScene hs;
    if (!load_scene(argv[2], hs)) { fprintf(stderr, "cannot read scene %s\n", argv[2]); return 2; }
    int W = atoi(argv[3]), H = atoi(argv[4]), spp = atoi(argv[5]);
    const char* out = argv[6];
    int use_bvh = argc > 7 ? atoi(argv[7]) : 1;
    int count = argc > 8 ? atoi(argv[8]) : 1;
    int reps = argc > 9 ? atoi(argv[9]) : 1;
    Built b = build_scene(hs, use_bvh);
    size_t npix = size_t(W) * H;
    curandState* d_st = nullptr;
    vec3 *d_fb = nullptr, *d_mean = nullptr;
    HCHECK(cudaMalloc(&d_st, npix * sizeof(curandState)));
    HCHECK(cudaMalloc(&d_fb, npix * sizeof(vec3)));
    HCHECK(cudaMalloc(&d_mean, npix * sizeof(vec3)));
    dim3 blocks(W / THREAD_SIZE_X + 1, H / THREAD_SIZE_Y + 1), threads(THREAD_SIZE_X, THREAD_SIZE_Y); // main.cu:434-435
    cudaEvent_t e0, e1, e2;
    HCHECK(cudaEventCreate(&e0));
    HCHECK(cudaEventCreate(&e1));
    HCHECK(cudaEventCreate(&e2));
    float best_init = 1e30f, best_render = 1e30f;
    for (int r = 0; r < reps; ++r) {
        HCHECK(cudaEventRecord(e0));
        init_rand_state_rt<<<blocks, threads>>>(d_st, W, H);
        HCHECK(cudaEventRecord(e1));
        render_rt<<<blocks, threads>>>(d_fb, d_mean, W, H, spp, b.ds.scene, b.ds.cam, d_st);
        HCHECK(cudaEventRecord(e2));
        HCHECK(cudaGetLastError());
        HCHECK(cudaDeviceSynchronize());
        float a, c;
        HCHECK(cudaEventElapsedTime(&a, e0, e1));
        HCHECK(cudaEventElapsedTime(&c, e1, e2));
        if (c < best_render) { best_render = c; best_init = a; }
    }
    unsigned long long rays = 0;
    if (count) {
        unsigned long long* d_total = nullptr;
        HCHECK(cudaMalloc(&d_total, sizeof *d_total));
        HCHECK(cudaMemset(d_total, 0, sizeof *d_total));
        count_rays_rt<<<blocks, threads>>>(d_total, W, H, spp, b.ds.scene, b.ds.cam, d_st);
        HCHECK(cudaGetLastError());
        HCHECK(cudaDeviceSynchronize());
        HCHECK(cudaMemcpy(&rays, d_total, sizeof rays, cudaMemcpyDeviceToHost));
    }
    // output: float32 [2][H][W][3]: finished framebuffer, then the un-tonemapped mean
    std::vector<float> host(npix * 6);
    HCHECK(cudaMemcpy(host.data(), d_fb, npix * sizeof(vec3), cudaMemcpyDeviceToHost));
    HCHECK(cudaMemcpy(host.data() + npix * 3, d_mean, npix * sizeof(vec3), cudaMemcpyDeviceToHost));
    if (!write_file(out, host.data(), host.size() * sizeof(float))) return 2;
    printf("{\"cmd\": \"render\", \"width\": %d, \"height\": %d, \"spp\": %d, \"use_bvh\": %d, \"paths\": %llu, \"rays\": %llu, "
           "\"ms_init_rand\": %.3f, \"ms_render\": %.3f, \"ms_build\": %.3f, \"reps\": %d}\n",
           W, H, spp, use_bvh, (unsigned long long)npix * spp, rays, best_init, best_render, b.ms_build, reps);
    return 0;
}

// earth <in.jpg> <out.f32>: writes int32 w, int32 h, then w*h*3 floats as main.cu:376-380 loads them
static int cmd_earth(int argc, char** argv) {
    if (argc < 4) return 2;
    int w, h, ch;
    stbi_ldr_to_hdr_scale(1.0f);
    stbi_ldr_to_hdr_gamma(1.0f);
    float* img = stbi_loadf(argv[2], &w, &h, &ch, 0);
    if (!img || ch != 3) { fprintf(stderr, "cannot decode %s\n", argv[2]); return 2; }
    FILE* f = fopen(argv[3], "wb");
    if (!f) return 2;
    int32_t wh[2] = {w, h};
    fwrite(wh, sizeof wh, 1, f);
    fwrite(img, sizeof(float), size_t(w) * h * 3, f);
    fclose(f);
    stbi_image_free(img);
    printf("{\"cmd\": \"earth\", \"width\": %d, \"height\": %d}\n", w, h);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr,
                "usage: ref_harness trace  <scene.rtsc> <rays.bin> <hits.bin> [use_bvh=1] [tmin=1e-5]\n"
                "       ref_harness render <scene.rtsc> W H spp <out.f32> [use_bvh=1] [count_rays=1] [reps=1]\n"
                "       ref_harness earth  <in.jpg> <out.f32>\n");
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd == "trace") return cmd_trace(argc, argv);
    if (cmd == "render") return cmd_render(argc, argv);
    if (cmd == "earth") return cmd_earth(argc, argv);
    return 2;
}
