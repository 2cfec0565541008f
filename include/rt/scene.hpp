// include/rt/scene.hpp — header-only host façade that keeps the reference's
// scene-construction API (class names and constructor signatures) and flattens
// the object graph into the POD arrays of rt_api.h.
//
// Reference (all under /root/reference/src): the same constructors are called
// from device code inside populate_scene_balls<<<1,1>>> (main.cu:188-356):
//   camera          camera.h:7-10        sphere / moving_sphere  sphere.h:9, :33-36
//   hitable_list    hitable_list.h:10    bvh_node                bvh.h:12
//   lambertian/metal/dielectric/emitter  material.h:61,74,94,40
//   constant/checker/noise/wood/image textures  texture.h:21,32,52,88-93,116
// Here they run on the host, own nothing on the device, and only describe the
// scene; rt_scene_create() does the upload and builds the acceleration structure.
// bvh_node(...) is therefore a *request* for an acceleration structure: its
// curandState*/level arguments are accepted and ignored (the reference reads an
// uninitialised curandState there, main.cu:317 vs :438).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../rt_api.h"

namespace rt {

// ---- vec3: host arithmetic of vec3.h's non-intrinsic branch (vec3.h:73-151) ----
class vec3 {
public:
    vec3() : _v{0.f, 0.f, 0.f} {}
    vec3(float v) : _v{v, v, v} {}
    // the reference mixes int / float / double literals (vec3(0.6, 0.1, 0.1), vec3(0, 1, 0)); all
    // of them convert to float exactly as its vec3(float, float, float) does
    template <class A, class B, class C> vec3(A a, B b, C c) : _v{float(a), float(b), float(c)} {}
    float x() const { return _v[0]; }
    float y() const { return _v[1]; }
    float z() const { return _v[2]; }
    float r() const { return _v[0]; }
    float g() const { return _v[1]; }
    float b() const { return _v[2]; }
    float operator[](int i) const { return _v[i]; }
    float& operator[](int i) { return _v[i]; }
    vec3 operator-() const { return vec3(-_v[0], -_v[1], -_v[2]); }
    vec3& operator+=(const vec3& o) { _v[0] += o._v[0]; _v[1] += o._v[1]; _v[2] += o._v[2]; return *this; }
    vec3& operator-=(const vec3& o) { _v[0] -= o._v[0]; _v[1] -= o._v[1]; _v[2] -= o._v[2]; return *this; }
    vec3& operator*=(float f) { _v[0] *= f; _v[1] *= f; _v[2] *= f; return *this; }
    float sq_length() const { return _v[0] * _v[0] + _v[1] * _v[1] + _v[2] * _v[2]; }
    float length() const { return std::sqrt(sq_length()); }
    bool is_null() const { return _v[0] == 0.f && _v[1] == 0.f && _v[2] == 0.f; }
    static float dot(const vec3& a, const vec3& b) { return a._v[0] * b._v[0] + a._v[1] * b._v[1] + a._v[2] * b._v[2]; }
    static vec3 cross(const vec3& a, const vec3& b) {
        return vec3(a._v[1] * b._v[2] - a._v[2] * b._v[1], -(a._v[0] * b._v[2] - a._v[2] * b._v[0]),
                    a._v[0] * b._v[1] - a._v[1] * b._v[0]);
    }
    static vec3 normalize(vec3 v) {
        if (v.is_null()) return v;
        float l = v.length();
        return vec3(v._v[0] / l, v._v[1] / l, v._v[2] / l);
    }
    friend vec3 operator+(const vec3& a, const vec3& b) { return vec3(a._v[0] + b._v[0], a._v[1] + b._v[1], a._v[2] + b._v[2]); }
    friend vec3 operator-(const vec3& a, const vec3& b) { return vec3(a._v[0] - b._v[0], a._v[1] - b._v[1], a._v[2] - b._v[2]); }
    friend vec3 operator*(const vec3& a, const vec3& b) { return vec3(a._v[0] * b._v[0], a._v[1] * b._v[1], a._v[2] * b._v[2]); }
    friend vec3 operator*(float t, const vec3& v) { return vec3(t * v._v[0], t * v._v[1], t * v._v[2]); }
    friend vec3 operator*(const vec3& v, float t) { return vec3(t * v._v[0], t * v._v[1], t * v._v[2]); }
    friend vec3 operator/(const vec3& v, float t) { return vec3(v._v[0] / t, v._v[1] / t, v._v[2] / t); }
    void store(float* out) const { out[0] = _v[0]; out[1] = _v[1]; out[2] = _v[2]; }

private:
    float _v[3];
};

// ---- textures (texture.h) -------------------------------------------------
enum class noise_type : uint8_t { PERLIN, TURBULANCE, MARBLE, UNKNOWN };

class text {
public:
    virtual ~text() = default;
    virtual rt_texture describe() const = 0;
    virtual const text* child(int) const { return nullptr; }
    virtual const rt_image* image() const { return nullptr; }
};

class constant_texture : public text {
public:
    constant_texture() {}
    constant_texture(vec3 col) : _col(col) {}
    rt_texture describe() const override {
        rt_texture t{};
        t.kind = RT_TEX_CONSTANT;
        t.even = t.odd = t.image = -1;
        _col.store(t.color1);
        return t;
    }

private:
    vec3 _col;
};

class checker_texture : public text {
public:
    checker_texture(const text* t0, const text* t1) : _even(t0), _odd(t1) {
        if (!t0 || !t1) throw std::invalid_argument("checker_texture: null child texture");
    }
    rt_texture describe() const override {
        rt_texture t{};
        t.kind = RT_TEX_CHECKER;
        t.even = t.odd = t.image = -1; // children resolved by the flattener
        return t;
    }
    const text* child(int i) const override { return i == 0 ? _even : _odd; }

private:
    const text* _even;
    const text* _odd;
};

class noise_texture : public text {
public:
    noise_texture(noise_type ntype = noise_type::PERLIN, float density = 4.f) : _density(density), _ntype(ntype) {
        if (_density <= 0.f) _density = 4.f; // texture.h:54-56
    }
    rt_texture describe() const override {
        rt_texture t{};
        t.even = t.odd = t.image = -1;
        t.density = _density;
        switch (_ntype) {
        case noise_type::PERLIN: t.kind = RT_TEX_NOISE_PERLIN; break;
        case noise_type::TURBULANCE: t.kind = RT_TEX_NOISE_TURBULANCE; break;
        case noise_type::MARBLE: t.kind = RT_TEX_NOISE_MARBLE; break;
        default: // texture.h:77-79: unknown type evaluates to white
            t.kind = RT_TEX_CONSTANT;
            t.color1[0] = t.color1[1] = t.color1[2] = 1.f;
        }
        return t;
    }

private:
    float _density;
    noise_type _ntype;
};

class wood_texture : public text {
public:
    wood_texture(const vec3& color1, const vec3& color2, float density = 4.f, float hardness = 50.f)
        : _density(density), _hardness(hardness), _color1(color1), _color2(color2) {
        if (_density <= 0.f) _density = 4.f; // texture.h:95-97
    }
    rt_texture describe() const override {
        rt_texture t{};
        t.kind = RT_TEX_WOOD;
        t.even = t.odd = t.image = -1;
        _color1.store(t.color1);
        _color2.store(t.color2);
        t.density = _density;
        t.hardness = _hardness;
        return t;
    }

private:
    float _density, _hardness;
    vec3 _color1, _color2;
};

class image_texture : public text {
public:
    // buffer: width*height*3 floats, row 0 = top (stbi_loadf layout, main.cu:378-380)
    image_texture(const float* buffer, int width, int height) : _img{buffer, width, height} {
        if (!buffer || width <= 0 || height <= 0) throw std::invalid_argument("image_texture: empty image");
    }
    rt_texture describe() const override {
        rt_texture t{};
        t.kind = RT_TEX_IMAGE;
        t.even = t.odd = -1;
        t.image = -1; // resolved by the flattener
        return t;
    }
    const rt_image* image() const override { return &_img; }

private:
    rt_image _img;
};

// ---- materials (material.h) -----------------------------------------------
class material {
public:
    virtual ~material() = default;
    virtual rt_material describe() const = 0;
    virtual const text* texture() const { return nullptr; }
};

class lambertian : public material {
public:
    lambertian(const text* tex) : _albedo(tex) {
        if (!tex) throw std::invalid_argument("lambertian: null texture");
    }
    rt_material describe() const override {
        rt_material m{};
        m.kind = RT_MAT_LAMBERTIAN;
        m.texture = -1;
        return m;
    }
    const text* texture() const override { return _albedo; }

private:
    const text* _albedo;
};

class metal : public material {
public:
    metal(const vec3& a, float r) : _albedo(a), _roughness(r < 1.f ? r : 1.f) {} // material.h:74-81
    rt_material describe() const override {
        rt_material m{};
        m.kind = RT_MAT_METAL;
        m.texture = -1;
        _albedo.store(m.albedo);
        m.param = _roughness;
        return m;
    }

private:
    vec3 _albedo;
    float _roughness;
};

class dielectric : public material {
public:
    dielectric(float ri, const vec3& tint) : _ri(ri), _tint(tint) {}
    rt_material describe() const override {
        rt_material m{};
        m.kind = RT_MAT_DIELECTRIC;
        m.texture = -1;
        _tint.store(m.albedo);
        m.param = _ri;
        return m;
    }

private:
    float _ri;
    vec3 _tint;
};

class emitter : public material {
public:
    emitter(const text* tex, float intensity = 1.f) : _texture(tex), _intensity(intensity) {
        if (!tex) throw std::invalid_argument("emitter: null texture");
    }
    rt_material describe() const override {
        rt_material m{};
        m.kind = RT_MAT_EMITTER;
        m.texture = -1;
        m.param = _intensity;
        return m;
    }
    const text* texture() const override { return _texture; }

private:
    const text* _texture;
    float _intensity;
};
using diffuse_light = emitter; // the north-star's name for the same class

// ---- hitables (hitable_object.h, sphere.h, bvh.h, hitable_list.h) -----------
enum class object_type { SPHERE, MOVING_SPHERE, HITABLE_LIST, BOUNDING_VOLUME_HIERARCHY, UNKNOWN };

class hitable_object {
public:
    virtual ~hitable_object() = default;
    virtual object_type get_object_type() const { return object_type::UNKNOWN; }
    bool is_leaf() const { return get_object_type() != object_type::BOUNDING_VOLUME_HIERARCHY; }
    uint32_t get_id() const { return _id; }
    void set_id(uint32_t id) { _id = id; }

private:
    uint32_t _id = 0;
};
using hitable = hitable_object;

class sphere : public hitable_object {
public:
    sphere(vec3 center, float radius, const material* mat, bool inside = false)
        : _c(center), _r(radius), _m(mat), _inside(inside) {
        if (!mat) throw std::invalid_argument("sphere: null material");
    }
    object_type get_object_type() const override { return object_type::SPHERE; }
    vec3 get_center() const { return _c; }
    float radius() const { return _r; }
    const material* mat() const { return _m; }
    bool inside() const { return _inside; }

private:
    vec3 _c;
    float _r;
    const material* _m;
    bool _inside;
};

class moving_sphere : public hitable_object {
public:
    moving_sphere(vec3 center0, vec3 center1, float time0, float time1, float radius, const material* mat)
        : _c0(center0), _c1(center1), _t0(time0), _t1(time1), _r(radius), _m(mat) {
        if (!mat) throw std::invalid_argument("moving_sphere: null material");
    }
    object_type get_object_type() const override { return object_type::MOVING_SPHERE; }
    vec3 center0() const { return _c0; }
    vec3 center1() const { return _c1; }
    float time0() const { return _t0; }
    float time1() const { return _t1; }
    float radius() const { return _r; }
    const material* mat() const { return _m; }

private:
    vec3 _c0, _c1;
    float _t0, _t1, _r;
    const material* _m;
};

class bvh_node : public hitable_object {
public:
    // (hlist, n, time0, time1, curandState*, level) — bvh.h:12. A request only.
    bvh_node(hitable_object** hlist, int n, float time0, float time1, void* /*rstate*/ = nullptr, int /*level*/ = 0,
             uint32_t mode = RT_BVH_AUTO)
        : _hlist(hlist), _n(n), _t0(time0), _t1(time1), _mode(mode) {}
    object_type get_object_type() const override { return object_type::BOUNDING_VOLUME_HIERARCHY; }
    uint32_t mode() const { return _mode; }
    int size() const { return _n; }

private:
    hitable_object** _hlist;
    int _n;
    float _t0, _t1;
    uint32_t _mode;
};

class hitable_list : public hitable_object {
public:
    // bvh == nullptr  =>  brute-force closest-hit loop (hitable_list.h:66-78)
    hitable_list(hitable_object** objs, bvh_node* bvh, uint32_t size) : _objs(objs), _bvh(bvh), _size(size) {}
    object_type get_object_type() const override { return object_type::HITABLE_LIST; }
    hitable_object* get_object(uint32_t id) const {
        for (uint32_t i = 0; i < _size; ++i)
            if (_objs[i]->get_id() == id) return _objs[i];
        return nullptr;
    }
    hitable_object** objects() const { return _objs; }
    const bvh_node* bvh() const { return _bvh; }
    uint32_t size() const { return _size; }

private:
    hitable_object** _objs;
    bvh_node* _bvh;
    uint32_t _size;
};

// ---- camera (camera.h:7-31) ---------------------------------------------------
class camera {
public:
    camera(vec3 lookfrom, vec3 lookat, vec3 up, float vfov, float aspect, float aperture, float focus_dist,
           float time0 = 0.f, float time1 = 0.f) {
        lookfrom.store(_c.lookfrom);
        lookat.store(_c.lookat);
        up.store(_c.up);
        _c.vfov = vfov;
        _c.aspect = aspect;
        _c.aperture = aperture;
        _c.focus_dist = focus_dist;
        _c.time0 = time0;
        _c.time1 = time1;
    }
    const rt_camera& describe() const { return _c; }

private:
    rt_camera _c;
};

// ---- arena: owns façade objects created while describing a scene ------------------
// (the reference leaks its textures and BVH nodes and frees the rest in device
// destructors, hitable_list.h:104-123 / sphere.h:148-155; here one arena frees all)
class arena {
public:
    arena() = default;
    arena(const arena&) = delete;
    arena& operator=(const arena&) = delete;
    ~arena() {
        for (auto it = _objs.rbegin(); it != _objs.rend(); ++it) delete *it;
        for (auto it = _mats.rbegin(); it != _mats.rend(); ++it) delete *it;
        for (auto it = _texs.rbegin(); it != _texs.rend(); ++it) delete *it;
        for (auto* c : _cams) delete c;
    }
    template <class T, class... A> T* tex(A&&... a) { T* p = new T(std::forward<A>(a)...); _texs.push_back(p); return p; }
    template <class T, class... A> T* mat(A&&... a) { T* p = new T(std::forward<A>(a)...); _mats.push_back(p); return p; }
    template <class T, class... A> T* obj(A&&... a) { T* p = new T(std::forward<A>(a)...); _objs.push_back(p); return p; }
    template <class... A> camera* cam(A&&... a) { camera* p = new camera(std::forward<A>(a)...); _cams.push_back(p); return p; }

private:
    std::vector<text*> _texs;
    std::vector<material*> _mats;
    std::vector<hitable_object*> _objs;
    std::vector<camera*> _cams;
};

// ---- flattening ------------------------------------------------------------------
struct flat_scene {
    std::vector<rt_sphere> spheres;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    std::vector<rt_image> images;
    rt_camera cam{};
    uint32_t bvh_mode = RT_BVH_AUTO;

    rt_scene_desc desc() const {
        rt_scene_desc d{};
        d.spheres = spheres.data();
        d.n_spheres = uint32_t(spheres.size());
        d.materials = materials.data();
        d.n_materials = uint32_t(materials.size());
        d.textures = textures.data();
        d.n_textures = uint32_t(textures.size());
        d.images = images.data();
        d.n_images = uint32_t(images.size());
        d.camera = cam;
        d.bvh_mode = bvh_mode;
        return d;
    }
};

namespace detail {
struct flattener {
    flat_scene& out;
    std::map<const material*, uint32_t> mat_ix;
    std::map<const text*, int32_t> tex_ix;
    std::map<const float*, int32_t> img_ix;

    int32_t add_texture(const text* t) {
        auto it = tex_ix.find(t);
        if (it != tex_ix.end()) return it->second;
        int32_t ix = int32_t(out.textures.size());
        tex_ix[t] = ix;
        out.textures.push_back(t->describe());
        if (out.textures[ix].kind == RT_TEX_CHECKER) {
            int32_t e = add_texture(t->child(0));
            int32_t o = add_texture(t->child(1));
            out.textures[ix].even = e;
            out.textures[ix].odd = o;
        } else if (out.textures[ix].kind == RT_TEX_IMAGE) {
            const rt_image* im = t->image();
            auto ii = img_ix.find(im->rgb);
            int32_t k;
            if (ii == img_ix.end()) {
                k = int32_t(out.images.size());
                out.images.push_back(*im);
                img_ix[im->rgb] = k;
            } else {
                k = ii->second;
            }
            out.textures[ix].image = k;
        }
        return ix;
    }
    uint32_t add_material(const material* m) {
        auto it = mat_ix.find(m);
        if (it != mat_ix.end()) return it->second;
        rt_material d = m->describe();
        if (m->texture()) d.texture = add_texture(m->texture());
        uint32_t ix = uint32_t(out.materials.size());
        out.materials.push_back(d);
        mat_ix[m] = ix;
        return ix;
    }
};
} // namespace detail

// Walks a hitable_list the way hitable_list::hit does (objects [0,size)), de-duplicating
// shared materials/textures/images, and records the acceleration-structure request.
inline flat_scene flatten(const hitable_list& list, const camera& cam) {
    flat_scene fs;
    detail::flattener f{fs, {}, {}, {}};
    for (uint32_t i = 0; i < list.size(); ++i) {
        const hitable_object* o = list.objects()[i];
        rt_sphere s{};
        if (o->get_object_type() == object_type::SPHERE) {
            const sphere* sp = static_cast<const sphere*>(o);
            sp->get_center().store(s.center0);
            sp->get_center().store(s.center1);
            s.radius = sp->radius();
            s.time0 = 0.f;
            s.time1 = 1.f;
            s.material = f.add_material(sp->mat());
            s.flags = sp->inside() ? RT_SPHERE_INSIDE : 0u;
        } else if (o->get_object_type() == object_type::MOVING_SPHERE) {
            const moving_sphere* ms = static_cast<const moving_sphere*>(o);
            ms->center0().store(s.center0);
            ms->center1().store(s.center1);
            s.radius = ms->radius();
            s.time0 = ms->time0();
            s.time1 = ms->time1();
            s.material = f.add_material(ms->mat());
            s.flags = RT_SPHERE_MOVING;
        } else {
            throw std::invalid_argument("flatten: hitable_list may only contain sphere / moving_sphere leaves");
        }
        s.id = o->get_id();
        fs.spheres.push_back(s);
    }
    fs.cam = cam.describe();
    fs.bvh_mode = list.bvh() ? list.bvh()->mode() : uint32_t(RT_BVH_NONE);
    return fs;
}

} // namespace rt
