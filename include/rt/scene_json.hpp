// include/rt/scene_json.hpp — runtime scene description (the reference's README roadmap item
// "scene description in JSON", README.md:10, and its compile-time WIDTH/HEIGHT/SAMPLES_PER_PIXEL/SEED macros,
// common.h:13-20, main.cu:15).  A JSON document is turned into the SAME façade objects a C++ caller would create
// (include/rt/scene.hpp: the reference's class names and constructor signatures), so everything downstream —
// flattening, BVH choice, the C-ABI — is shared with the hard-coded scenes of scenes.hpp.
//
//   { "camera":    { "lookfrom":[x,y,z], "lookat":[..], "up":[0,1,0], "vfov":20, "aspect":2, "aperture":0.25,
//                    "focus_dist": <number> | "auto"  (auto = |lookfrom - lookat|, main.cu:333), "time0":0, "time1":0.2 },
//     "textures":  { "<name>": {"type":"constant","color":[r,g,b]} | {"type":"checker","even":"<name>","odd":"<name>"}
//                    | {"type":"noise","noise":"PERLIN|TURBULANCE|MARBLE","density":4}
//                    | {"type":"wood","color1":[..],"color2":[..],"density":4,"hardness":50}
//                    | {"type":"image","file":"earth.ppm"} },
//     "materials": { "<name>": {"type":"lambertian","texture":"<name>"} | {"type":"metal","albedo":[..],"roughness":0.5}
//                    | {"type":"dielectric","ri":1.5,"tint":[1,1,1]} | {"type":"emitter","texture":"<name>","intensity":1} },
//     "objects":   [ {"type":"sphere","center":[..],"radius":0.5,"material":"<name>","id":0,"inside":false}
//                    | {"type":"moving_sphere","center0":[..],"center1":[..],"time0":0,"time1":1,"radius":0.2,"material":..} ],
//     "bvh": "auto" | "none" | "sah" | "lbvh",          (none = hitable_list without a bvh_node, hitable_list.h:66-78)
//     "render": { "width":1200, "height":600, "spp":100, "max_depth":50, "seed":1000, "tmin":1e-5,
//                 "world":[1,0.8,0.7], "bloom":0.1, "emitter_sampling":0 } }
// Textures and materials are created in document order; objects get ids in array order unless "id" is given.
#pragma once

#include <cmath>
#include <cstdio>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "scenes.hpp"

namespace rt {
namespace json {

struct value {
    enum kind_t { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<value> arr;
    std::vector<std::pair<std::string, value>> obj; // document order is kept

    const value* find(const std::string& k) const {
        for (const auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

class parser {
public:
    explicit parser(const std::string& s) : _s(s) {}
    value parse() {
        value v = val();
        ws();
        if (_i != _s.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& _s;
    size_t _i = 0;
    int _depth = 0; // containers open at _i: scene documents nest 4 deep, and val() recurses once per level
    static constexpr int kMaxDepth = 64;
    struct nested {
        parser& p;
        explicit nested(parser& q) : p(q) {
            if (++p._depth > kMaxDepth) p.fail("nesting too deep");
        }
        ~nested() { --p._depth; }
    };
    [[noreturn]] void fail(const char* what) const {
        throw std::invalid_argument("JSON: " + std::string(what) + " at offset " + std::to_string(_i));
    }
    void ws() {
        while (_i < _s.size() && (_s[_i] == ' ' || _s[_i] == '\n' || _s[_i] == '\t' || _s[_i] == '\r')) ++_i;
    }
    bool eat(char c) {
        ws();
        if (_i < _s.size() && _s[_i] == c) {
            ++_i;
            return true;
        }
        return false;
    }
    value val() {
        ws();
        if (_i >= _s.size()) fail("unexpected end");
        const char c = _s[_i];
        value v;
        if (c == '{') {
            const nested level(*this);
            ++_i;
            v.kind = value::OBJ;
            if (eat('}')) return v;
            do {
                ws();
                value k = str();
                if (!eat(':')) fail("expected ':'");
                v.obj.emplace_back(k.str, val());
            } while (eat(','));
            if (!eat('}')) fail("expected '}'");
        } else if (c == '[') {
            const nested level(*this);
            ++_i;
            v.kind = value::ARR;
            if (eat(']')) return v;
            do v.arr.push_back(val());
            while (eat(','));
            if (!eat(']')) fail("expected ']'");
        } else if (c == '"') {
            v = str();
        } else if (_s.compare(_i, 4, "true") == 0) {
            _i += 4;
            v.kind = value::BOOL;
            v.b = true;
        } else if (_s.compare(_i, 5, "false") == 0) {
            _i += 5;
            v.kind = value::BOOL;
        } else if (_s.compare(_i, 4, "null") == 0) {
            _i += 4;
        } else {
            size_t n = 0;
            try {
                v.num = std::stod(_s.substr(_i, 64), &n);
            } catch (const std::exception&) {
                fail("bad number");
            }
            v.kind = value::NUM;
            _i += n;
        }
        return v;
    }
    value str() {
        if (_i >= _s.size() || _s[_i] != '"') fail("expected string");
        ++_i;
        value v;
        v.kind = value::STR;
        while (_i < _s.size() && _s[_i] != '"') {
            char c = _s[_i++];
            if (c == '\\') {
                if (_i >= _s.size()) fail("bad escape");
                const char e = _s[_i++];
                c = e == 'n' ? '\n' : e == 't' ? '\t' : e; // \" \\ \/ and friends; \uXXXX is not needed for scene names
            }
            v.str.push_back(c);
        }
        if (_i >= _s.size()) fail("unterminated string");
        ++_i;
        return v;
    }
};

} // namespace json

namespace scenes {

// image files referenced by a document; the loader callback returns width*height*3 floats (stbi_loadf layout)
using image_loader = std::function<bool(const std::string& path, std::vector<float>& rgb, int& w, int& h)>;

struct json_scene {
    built b;
    std::vector<std::vector<float>> images; // storage behind image_texture (which borrows, texture.h:145)
};

namespace detail {
inline float num(const json::value& o, const char* key, double dflt, bool required = false) {
    const json::value* v = o.find(key);
    if (!v) {
        if (required) throw std::invalid_argument(std::string("scene JSON: missing \"") + key + "\"");
        return float(dflt);
    }
    if (v->kind != json::value::NUM) throw std::invalid_argument(std::string("scene JSON: \"") + key + "\" is not a number");
    if (!std::isfinite(v->num) || std::fabs(v->num) > 3.0e38) throw std::invalid_argument(std::string("scene JSON: \"") + key + "\" is not a finite float");
    return float(v->num);
}
// integer fields (seed, object ids, frame size, sample counts): taken from the document's double — exact up to 2^53 —
// with a range check, never through float (which rounds above 2^24) and never by casting an out-of-range value
inline long long integer(const json::value& o, const char* key, long long dflt, long long lo, long long hi) {
    const json::value* v = o.find(key);
    if (!v) return dflt;
    if (v->kind != json::value::NUM || !(v->num >= double(lo) && v->num <= double(hi)) || v->num != std::floor(v->num))
        throw std::invalid_argument(std::string("scene JSON: \"") + key + "\" must be an integer in [" + std::to_string(lo) + ", " +
                                    std::to_string(hi) + "]");
    return (long long)v->num;
}
inline vec3 vec(const json::value& o, const char* key, vec3 dflt, bool required = false) {
    const json::value* v = o.find(key);
    if (!v) {
        if (required) throw std::invalid_argument(std::string("scene JSON: missing \"") + key + "\"");
        return dflt;
    }
    if (v->kind != json::value::ARR || v->arr.size() != 3) throw std::invalid_argument(std::string("scene JSON: \"") + key + "\" is not [x,y,z]");
    for (const auto& c : v->arr)
        if (c.kind != json::value::NUM) throw std::invalid_argument(std::string("scene JSON: \"") + key + "\" is not [x,y,z]");
    return vec3(v->arr[0].num, v->arr[1].num, v->arr[2].num);
}
inline std::string str(const json::value& o, const char* key, const char* dflt = nullptr) {
    const json::value* v = o.find(key);
    if (!v) {
        if (!dflt) throw std::invalid_argument(std::string("scene JSON: missing \"") + key + "\"");
        return dflt;
    }
    if (v->kind != json::value::STR) throw std::invalid_argument(std::string("scene JSON: \"") + key + "\" is not a string");
    return v->str;
}
} // namespace detail

// Builds the façade objects of a JSON scene.  `render` (may be null) receives the "render" block on top of its
// current contents.  Throws std::invalid_argument on malformed input.
inline void scene_from_json(const std::string& doc_text, const std::string& base_dir, const image_loader& load_image, arena& A,
                            json_scene& out, rt_render_params* render) {
    using namespace detail;
    const json::value doc = json::parser(doc_text).parse();
    if (doc.kind != json::value::OBJ) throw std::invalid_argument("scene JSON: the document is not an object");

    std::map<std::string, text*> texs;
    if (const json::value* ts = doc.find("textures")) {
        for (const auto& kv : ts->obj) {
            const json::value& t = kv.second;
            const std::string type = str(t, "type");
            text* made = nullptr;
            if (type == "constant") {
                made = A.tex<constant_texture>(vec(t, "color", vec3(0, 0, 0), true));
            } else if (type == "checker") {
                auto e = texs.find(str(t, "even")), o = texs.find(str(t, "odd"));
                if (e == texs.end() || o == texs.end()) throw std::invalid_argument("scene JSON: checker child texture \"" + kv.first + "\" refers to an undefined texture");
                made = A.tex<checker_texture>(e->second, o->second);
            } else if (type == "noise") {
                const std::string n = str(t, "noise", "PERLIN");
                const noise_type nt = n == "PERLIN" ? noise_type::PERLIN : n == "TURBULANCE" ? noise_type::TURBULANCE
                                      : n == "MARBLE" ? noise_type::MARBLE : noise_type::UNKNOWN;
                if (nt == noise_type::UNKNOWN) throw std::invalid_argument("scene JSON: unknown noise kind " + n);
                made = A.tex<noise_texture>(nt, num(t, "density", 4.0));
            } else if (type == "wood") {
                made = A.tex<wood_texture>(vec(t, "color1", vec3(0, 0, 0), true), vec(t, "color2", vec3(0, 0, 0), true),
                                           num(t, "density", 4.0), num(t, "hardness", 50.0));
            } else if (type == "image") {
                std::string file = str(t, "file");
                if (!file.empty() && file[0] != '/' && !base_dir.empty()) file = base_dir + "/" + file;
                out.images.emplace_back();
                int w = 0, h = 0;
                if (!load_image || !load_image(file, out.images.back(), w, h))
                    throw std::invalid_argument("scene JSON: cannot load image " + file);
                made = A.tex<image_texture>(out.images.back().data(), w, h);
            } else {
                throw std::invalid_argument("scene JSON: unknown texture type " + type);
            }
            texs[kv.first] = made;
        }
    }
    auto tex_ref = [&](const json::value& o) -> text* {
        auto it = texs.find(str(o, "texture"));
        if (it == texs.end()) throw std::invalid_argument("scene JSON: undefined texture \"" + str(o, "texture") + "\"");
        return it->second;
    };

    std::map<std::string, material*> mats;
    if (const json::value* ms = doc.find("materials")) {
        for (const auto& kv : ms->obj) {
            const json::value& m = kv.second;
            const std::string type = str(m, "type");
            material* made = nullptr;
            if (type == "lambertian") made = A.mat<lambertian>(tex_ref(m));
            else if (type == "metal") made = A.mat<metal>(vec(m, "albedo", vec3(1, 1, 1), true), num(m, "roughness", 0.0));
            else if (type == "dielectric") made = A.mat<dielectric>(num(m, "ri", 1.5), vec(m, "tint", vec3(1, 1, 1)));
            else if (type == "emitter" || type == "diffuse_light") made = A.mat<emitter>(tex_ref(m), num(m, "intensity", 1.0));
            else throw std::invalid_argument("scene JSON: unknown material type " + type);
            mats[kv.first] = made;
        }
    }

    const json::value* objs = doc.find("objects");
    if (!objs || objs->kind != json::value::ARR) throw std::invalid_argument("scene JSON: \"objects\" must be an array");
    built& b = out.b;
    float t_lo = 0.f, t_hi = 0.f;
    for (const json::value& o : objs->arr) {
        const std::string type = str(o, "type");
        auto mit = mats.find(str(o, "material"));
        if (mit == mats.end()) throw std::invalid_argument("scene JSON: undefined material \"" + str(o, "material") + "\"");
        hitable_object* made = nullptr;
        if (type == "sphere") {
            const json::value* in = o.find("inside");
            made = A.obj<sphere>(vec(o, "center", vec3(0, 0, 0), true), num(o, "radius", 0, true), mit->second,
                                 in && in->kind == json::value::BOOL && in->b);
        } else if (type == "moving_sphere") {
            const float t0 = num(o, "time0", 0.0), t1 = num(o, "time1", 1.0);
            made = A.obj<moving_sphere>(vec(o, "center0", vec3(0, 0, 0), true), vec(o, "center1", vec3(0, 0, 0), true), t0, t1,
                                        num(o, "radius", 0, true), mit->second);
            t_lo = std::fmin(t_lo, t0);
            t_hi = std::fmax(t_hi, t1);
        } else {
            throw std::invalid_argument("scene JSON: unknown object type " + type);
        }
        made->set_id(uint32_t(integer(o, "id", (long long)b.objects.size(), 0, 0xffffffffll)));
        b.objects.push_back(made);
    }
    const uint32_t n = uint32_t(b.objects.size());
    const std::string bvh = str(doc, "bvh", "auto");
    bvh_node* node = nullptr;
    if (bvh != "none") {
        const uint32_t mode = bvh == "sah" ? RT_BVH_HOST_SAH : bvh == "lbvh" ? RT_BVH_GPU_LBVH : RT_BVH_AUTO;
        if (bvh != "auto" && bvh != "sah" && bvh != "lbvh") throw std::invalid_argument("scene JSON: unknown bvh mode " + bvh);
        b.objects.push_back(nullptr); // the reference keeps the bvh_node in the same array (main.cu:317)
        node = A.obj<bvh_node>(b.objects.data(), int(n), t_lo, t_hi, nullptr, 0, mode);
        node->set_id(n);
        b.objects[n] = node;
    }
    b.list = A.obj<hitable_list>(b.objects.data(), node, n);
    b.list->set_id(n + 1);

    const json::value* c = doc.find("camera");
    if (!c) throw std::invalid_argument("scene JSON: missing \"camera\"");
    const vec3 from = vec(*c, "lookfrom", vec3(0, 0, 0), true), at = vec(*c, "lookat", vec3(0, 0, 0), true);
    const json::value* fd = c->find("focus_dist");
    const float focus = (!fd || fd->kind == json::value::STR) ? (from - at).length() : float(fd->num);
    b.cam = A.cam(from, at, vec(*c, "up", vec3(0, 1, 0)), num(*c, "vfov", 20.0), num(*c, "aspect", 2.0), num(*c, "aperture", 0.0), focus,
                  num(*c, "time0", 0.0), num(*c, "time1", 0.0));

    if (render) {
        if (const json::value* r = doc.find("render")) {
            render->width = int32_t(integer(*r, "width", render->width, 1, 0x7fffffffll));
            render->height = int32_t(integer(*r, "height", render->height, 1, 0x7fffffffll));
            render->spp = int32_t(integer(*r, "spp", render->spp, 0, 0x7fffffffll));
            render->max_depth = int32_t(integer(*r, "max_depth", render->max_depth, 0, (1 << 23) - 1));
            render->seed = uint32_t(integer(*r, "seed", render->seed, 0, 0xffffffffll));
            render->tmin = num(*r, "tmin", render->tmin);
            render->bloom = num(*r, "bloom", render->bloom);
            vec(*r, "world", vec3(render->world[0], render->world[1], render->world[2])).store(render->world);
            // "emitter_sampling": 1 -> RT_RENDER_EMITTER_SAMPLING (the reference README's roadmap item, README.md:27-28)
            if (num(*r, "emitter_sampling", (render->flags & RT_RENDER_EMITTER_SAMPLING) ? 1.0 : 0.0) != 0.0)
                render->flags |= RT_RENDER_EMITTER_SAMPLING;
            else render->flags &= ~RT_RENDER_EMITTER_SAMPLING;
        }
    }
}

} // namespace scenes
} // namespace rt
