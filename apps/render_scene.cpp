// apps/render_scene.cpp — the C++ host program that replaces the reference's main() (main.cu:368-512)
// for the render path: builds a scene through the façade (include/rt/scenes.hpp keeps the reference's
// constructor calls), hands the flattened description to the C-ABI, renders, converts with the
// reference's writer loop and stores the frame.
//
//   render_scene [scene] [width height spp] [out.ppm] [earth.ppm]
//     scene: earth_emitter (default) | book1_final | perlin_motion | random_spheres:N
//
// The earth texture is read from a binary PPM of the stb-decoded JPEG (see tools/make_assets.py);
// JPEG decode/encode themselves are vendored stb in the reference and stay outside this library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/rt/scenes.hpp"

#define CHECK(x)                                                                   \
    do {                                                                           \
        rt_status st__ = (x);                                                      \
        if (st__ != RT_OK) {                                                       \
            fprintf(stderr, "%s failed (%d): %s\n", #x, int(st__), rt_last_error()); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

int main(int argc, char** argv) {
    std::string scene_name = argc > 1 ? argv[1] : "earth_emitter";
    rt_render_params p;
    rt_default_render_params(&p); // 1200x600x100, depth 50, seed 1000, tmin 1e-5 (common.h:13-20, main.cu:15,45)
    if (argc > 4) {
        p.width = atoi(argv[2]);
        p.height = atoi(argv[3]);
        p.spp = atoi(argv[4]);
    }
    const char* out_path = argc > 5 ? argv[5] : "render.ppm";
    const char* earth_path = argc > 6 ? argv[6] : "assets/earth_stb.ppm";

    rt::arena A;
    rt::scenes::built b;
    float* earth = nullptr;
    if (scene_name == "earth_emitter") {
        int32_t ew = 0, eh = 0;
        CHECK(rt_read_ppm_f32(earth_path, &earth, &ew, &eh)); // stbi_loadf equivalent: byte/255.f (main.cu:376-380)
        b = rt::scenes::earth_emitter(A, earth, ew, eh);
    } else if (scene_name == "book1_final") {
        b = rt::scenes::book1_final(A);
    } else if (scene_name == "perlin_motion") {
        b = rt::scenes::perlin_motion(A);
    } else if (scene_name.rfind("random_spheres", 0) == 0) {
        uint32_t n = scene_name.size() > 15 ? uint32_t(atoll(scene_name.c_str() + 15)) : 100000u;
        b = rt::scenes::random_spheres(A, n);
    } else {
        fprintf(stderr, "unknown scene %s\n", scene_name.c_str());
        return 2;
    }
    rt::flat_scene fs = rt::flatten(*b.list, *b.cam);
    rt_scene_desc desc = fs.desc();

    rt_context* ctx = nullptr;
    rt_scene* scene = nullptr;
    CHECK(rt_context_create(0, &ctx));
    CHECK(rt_scene_create(ctx, &desc, &scene));
    rt_scene_info info;
    CHECK(rt_scene_get_info(scene, &info));
    printf("Rendering a %dx%d image (%d samples per pixel): %u spheres, %u BVH nodes (mode %u, build %.3f ms)\n", p.width,
           p.height, p.spp, info.n_spheres, info.n_nodes, info.bvh_mode, info.ms_build);

    std::vector<float> fb(size_t(p.width) * p.height * 3);
    rt_stats st;
    CHECK(rt_render(ctx, scene, &p, fb.data(), &st));
    printf("took %.0fus.  (%.1f Mpaths/s, %.1f Mrays/s, %u launches)\n", st.ms_total * 1e3, st.paths / st.ms_total / 1e3,
           st.rays / st.ms_total / 1e3, st.launches);

    std::vector<uint8_t> img(fb.size());
    CHECK(rt_quantize_rgb8(fb.data(), p.width, p.height, img.data())); // main.cu:475-488
    CHECK(rt_write_ppm(out_path, p.width, p.height, img.data()));
    rt_scene_destroy(scene);
    rt_context_destroy(ctx);
    rt_free(earth);
    return 0;
}
