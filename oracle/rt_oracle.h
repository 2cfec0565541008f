/* oracle/rt_oracle.h — TEST INFRASTRUCTURE: C interface of the CPU restatement (liboracle.so).
 * See rt_oracle.cpp for what `arith` and `sampler` select.  Scenes are the PODs of rt_api.h. */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include "../include/rt_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_scene orc_scene;
orc_scene* orc_scene_create(const rt_scene_desc* desc);
void orc_scene_destroy(orc_scene* s);
/* hitable_list::hit without a BVH (hitable_list.h:66-78) for caller-supplied rays */
void orc_trace(const orc_scene* s, const rt_ray* rays, size_t n, float tmin, int arith, rt_hit* hits);
/* render (main.cu:97-132): accum = width*height*4 floats (sum r,g,b, count), j = 0 bottom row */
void orc_render(const orc_scene* s, const rt_render_params* rp, int sampler, int arith, int nthreads, float* accum,
                unsigned long long* rays_out);
void orc_shade_probe(const orc_scene* s, const rt_ray* rays, size_t n, const rt_render_params* rp, int arith,
                     rt_shade_sample* out);
void orc_tonemap(const float* accum, int width, int height, float* out_rgb); /* main.cu:124-127 */
float orc_perlin_noise(const float p[3]);
float orc_turbulence(const float p[3]);
void orc_texture_value(const orc_scene* s, int tex, float u, float v, const float p[3], float out[3]);
void orc_reflect(const float v[3], const float n[3], float out[3]);
int orc_refract(const float v[3], const float n[3], float mu, float out[3]);
float orc_shlick(float cosine, float ri);
void orc_sphere_uv(const float n[3], float* u, float* v);
void orc_camera_ray(const orc_scene* s, float s_, float t_, unsigned long long seed, rt_ray* out);
/* RT_RENDER_EMITTER_SAMPLING: the shadow ray of a lambertian hit, returns p_ref / p_sel (the scene must hold an emitter) */
float orc_light_sample(const orc_scene* s, const float p[3], const float n[3], float time, uint32_t seed, uint32_t index,
                       float dir[3]);
/* jpeg_oracle.cpp: stbi_write_jpg(..., comp 3, quality) (main.cu:491) into memory; returns the file size
 * (out == NULL: size only), 0 if cap is too small */
size_t orc_jpeg_encode(const uint8_t* rgb8, int width, int height, int quality, uint8_t* out, size_t cap);
/* jpeg_decode_oracle.cpp: pixel stages of stb_image's JPEG reader; out = height*width*(3 or 1) bytes; 0 on success */
int orc_jpeg_pixels(const rt_jpeg_coefficients* c, uint8_t* out);
#ifdef __cplusplus
}
#endif
#endif
