// host_scene_import.cpp — Collada (.dae) scene import: the on-disk input format of the path.
//
// The reference reads scenes through assimp (RenderEngine/scene/Scene.cpp:73-176) and then applies its own rules: material
// classification (Scene.cpp:178-265), point / spot lights (:267-285), quad emitters from the first triangle of an emissive
// mesh (:287-311), the scene AABB (:147-160), the default camera (:469-485), one material per mesh (:496-503).  assimp is
// a third-party dependency that is not under /root/reference and not in this image, so this file restates the part of it the
// reference relies on for COLLADA 1.4 documents — <library_geometries> (triangles / polylist / polygons), effects with
// phong / lambert / blinn / constant techniques, images, lights, cameras and the <visual_scene> node tree — and the
// post-processing steps the reference requests (Scene.cpp:96-108): Triangulate (fan), JoinIdenticalVertices (by index tuple),
// PreTransformVertices (node matrices baked, lights and cameras included), GenSmoothNormals (area-weighted, per position) and
// CalcTangentSpace (per-face UV-gradient tangents averaged per vertex).  Parity of those steps against assimp itself is
// UNPINNED (assimp cannot be run here); everything downstream of them follows Scene.cpp line by line, including its quirks:
//   * emitter lights are appended twice (once in createFromFile :141, once more when the scene graph is built :508);
//   * the AABB uses the first vertex of every face three times (:151-153);
//   * the camera field of view is mHorizontalFOV * 365 / (2 pi) (:480-481).
// No CUDA here.
#include "../../include/opr.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <stdexcept>

namespace {

thread_local std::string g_importError;

// ---- a small XML reader ---------------------------------------------------------------------------------------------------
struct Xml {
    std::string name, text;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<Xml>> kids;
    const char* attr(const char* k) const { for (auto& a : attrs) if (a.first == k) return a.second.c_str(); return nullptr; }
    std::string attrs_or(const char* k, const char* d = "") const { const char* v = attr(k); return v ? v : d; }
    const Xml* child(const char* n) const { for (auto& c : kids) if (c->name == n) return c.get(); return nullptr; }
    std::vector<const Xml*> children(const char* n) const { std::vector<const Xml*> r; for (auto& c : kids) if (c->name == n) r.push_back(c.get()); return r; }
    const Xml* find(const char* n) const   // depth-first descendant
    {
        for (auto& c : kids) { if (c->name == n) return c.get(); if (const Xml* f = c->find(n)) return f; }
        return nullptr;
    }
};

std::string unescape(const std::string& s)
{
    if (s.find('&') == std::string::npos) return s;
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] != '&') { o += s[i]; continue; }
        static const char* ent[][2] = {{"&lt;", "<"}, {"&gt;", ">"}, {"&amp;", "&"}, {"&quot;", "\""}, {"&apos;", "'"}};
        bool done = false;
        for (auto& e : ent) { size_t n = strlen(e[0]); if (s.compare(i, n, e[0]) == 0) { o += e[1]; i += n - 1; done = true; break; } }
        if (!done) o += s[i];
    }
    return o;
}

struct XmlParser {
    const std::string& s; size_t p = 0;
    explicit XmlParser(const std::string& src) : s(src) {}
    void skipWs() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    bool starts(const char* t) const { return s.compare(p, strlen(t), t) == 0; }
    std::unique_ptr<Xml> parseDocument()
    {
        for (;;) {
            skipWs();
            if (p >= s.size()) throw std::runtime_error("empty XML document");
            if (starts("<?")) { p = s.find("?>", p); if (p == std::string::npos) throw std::runtime_error("unterminated <?"); p += 2; }
            else if (starts("<!--")) { p = s.find("-->", p); if (p == std::string::npos) throw std::runtime_error("unterminated comment"); p += 3; }
            else if (starts("<!")) { p = s.find('>', p); if (p == std::string::npos) throw std::runtime_error("unterminated <!"); p += 1; }
            else break;
        }
        return parseElement();
    }
    std::unique_ptr<Xml> parseElement()
    {
        if (s[p] != '<') throw std::runtime_error("XML: expected '<'");
        p++;
        std::unique_ptr<Xml> e(new Xml);
        while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/') e->name += s[p++];
        for (;;) {
            skipWs();
            if (p >= s.size()) throw std::runtime_error("XML: unterminated tag <" + e->name);
            if (s[p] == '/') { p = s.find('>', p); if (p == std::string::npos) throw std::runtime_error("XML: bad tag"); p++; return e; }
            if (s[p] == '>') { p++; break; }
            std::string k, v;
            while (p < s.size() && s[p] != '=' && !isspace((unsigned char)s[p])) k += s[p++];
            skipWs();
            if (p >= s.size() || s[p] != '=') throw std::runtime_error("XML: attribute without value in <" + e->name);
            p++; skipWs();
            const char q = s[p];
            if (q != '"' && q != '\'') throw std::runtime_error("XML: unquoted attribute in <" + e->name);
            size_t end = s.find(q, p + 1);
            if (end == std::string::npos) throw std::runtime_error("XML: unterminated attribute");
            v = s.substr(p + 1, end - p - 1); p = end + 1;
            e->attrs.emplace_back(k, unescape(v));
        }
        for (;;) {
            size_t lt = s.find('<', p);
            if (lt == std::string::npos) throw std::runtime_error("XML: missing </" + e->name + ">");
            e->text.append(s, p, lt - p);
            p = lt;
            if (starts("</")) { p = s.find('>', p); if (p == std::string::npos) throw std::runtime_error("XML: bad end tag"); p++; break; }
            if (starts("<!--")) { p = s.find("-->", p); if (p == std::string::npos) throw std::runtime_error("unterminated comment"); p += 3; continue; }
            if (starts("<![CDATA[")) { size_t end = s.find("]]>", p); if (end == std::string::npos) throw std::runtime_error("unterminated CDATA"); e->text.append(s, p + 9, end - p - 9); p = end + 3; continue; }
            e->kids.push_back(parseElement());
        }
        e->text = unescape(e->text);
        return e;
    }
};

// ---- small math ---------------------------------------------------------------------------------------------------------
struct V3 { float x = 0, y = 0, z = 0; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 normalizeSafe(V3 a) { float l = sqrtf(dot(a, a)); return l > 0.f ? a * (1.0f / l) : a; }

struct M4 {
    float m[16];   // row-major
    static M4 identity() { M4 r; for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.f : 0.f; return r; }
    M4 operator*(const M4& b) const
    {
        M4 r;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0; for (int k = 0; k < 4; k++) s += m[4 * i + k] * b.m[4 * k + j]; r.m[4 * i + j] = s; }
        return r;
    }
    V3 point(V3 p) const { return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]}; }
    V3 dir(V3 p) const { return {m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z, m[8] * p.x + m[9] * p.y + m[10] * p.z}; }
    // inverse transpose of the upper 3x3, for normals (assimp PretransformVertices: mWorldIT = m.Inverse().Transpose())
    M4 normalMatrix() const
    {
        const float a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
        const float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
        M4 r = identity();
        if (det == 0.f) return r;
        const float id = 1.0f / det;
        // inverse (row-major) then transposed == cofactor matrix / det
        r.m[0] = (e * i - f * h) * id; r.m[1] = (f * g - d * i) * id; r.m[2] = (d * h - e * g) * id;
        r.m[4] = (c * h - b * i) * id; r.m[5] = (a * i - c * g) * id; r.m[6] = (b * g - a * h) * id;
        r.m[8] = (b * f - c * e) * id; r.m[9] = (c * d - a * f) * id; r.m[10] = (a * e - b * d) * id;
        return r;
    }
};

std::vector<float> floats(const std::string& t)
{
    std::vector<float> v; const char* p = t.c_str(); char* e;
    for (;;) { float f = strtof(p, &e); if (e == p) break; v.push_back(f); p = e; }
    return v;
}
std::vector<long> ints(const std::string& t)
{
    std::vector<long> v; const char* p = t.c_str(); char* e;
    for (;;) { long f = strtol(p, &e, 10); if (e == p) break; v.push_back(f); p = e; }
    return v;
}
std::string frag(const std::string& url) { return !url.empty() && url[0] == '#' ? url.substr(1) : url; }

// ---- what assimp would hand to Scene.cpp ---------------------------------------------------------------------------------
struct AiMaterial {
    std::string name;
    V3 emissive, diffuse{0.6f, 0.6f, 0.6f}, reflective;     // ColladaParser defaults: diffuse 0.6, emission / reflective 0
    float refracti = 1.0f;
    std::string diffuseTex, normalsTex;
};
struct AiMesh {
    std::vector<V3> vertices, normals, tangents, bitangents;
    std::vector<float> uvs;          // 2 per vertex, empty when absent
    std::vector<int> faces;          // 3 per face
    int materialIndex = 0;
};
struct AiLight { int type; V3 color, position, direction; float innerCone; };   // 1 = point, 2 = spot
struct AiCamera { V3 position, lookAt{0, 0, -1}, up{0, 1, 0}; float hfovRad = 0.f; };

struct Source { std::vector<float> data; int stride = 1; };

struct Importer {
    std::string dir;
    const Xml* root = nullptr;
    std::map<std::string, const Xml*> byId;
    std::vector<AiMaterial> materials; std::map<std::string, int> materialIndex;
    std::vector<AiMesh> meshes; std::vector<AiLight> lights; std::vector<AiCamera> cameras;

    void index(const Xml* e) { if (const char* id = e->attr("id")) byId[id] = e; for (auto& c : e->kids) index(c.get()); }

    static V3 color(const Xml* e, V3 dflt)
    {
        if (!e) return dflt;
        const Xml* c = e->child("color");
        if (!c) return dflt;
        auto f = floats(c->text);
        return f.size() >= 3 ? V3{f[0], f[1], f[2]} : dflt;
    }
    std::string texturePath(const Xml* effectProfile, const Xml* slot) const
    {
        if (!slot) return "";
        const Xml* t = slot->child("texture");
        if (!t) return "";
        std::string ref = t->attrs_or("texture");
        // sampler sid -> surface sid -> image id (COLLADA 1.4 newparam chain); tolerate a direct image id
        for (int hop = 0; hop < 2; hop++)
            for (const Xml* np : effectProfile->children("newparam"))
                if (np->attrs_or("sid") == ref) {
                    if (const Xml* s2 = np->child("sampler2D")) { if (const Xml* src = s2->child("source")) { ref = src->text; } }
                    else if (const Xml* sf = np->child("surface")) { if (const Xml* in = sf->child("init_from")) { ref = in->text; } }
                }
        ref.erase(0, ref.find_first_not_of(" \t\r\n")); ref.erase(ref.find_last_not_of(" \t\r\n") + 1);
        auto it = byId.find(ref);
        if (it == byId.end()) return "";
        const Xml* init = it->second->find("init_from");
        if (!init) return "";
        std::string f = init->text;
        f.erase(0, f.find_first_not_of(" \t\r\n")); f.erase(f.find_last_not_of(" \t\r\n") + 1);
        if (f.compare(0, 7, "file://") == 0) f = f.substr(7);
        return f;
    }
    int materialFor(const std::string& materialId)
    {
        auto it = materialIndex.find(materialId);
        if (it != materialIndex.end()) return it->second;
        AiMaterial m; m.name = materialId;
        auto mi = byId.find(materialId);
        if (mi != byId.end()) {
            m.name = mi->second->attrs_or("name", materialId.c_str());
            if (const Xml* ie = mi->second->child("instance_effect")) {
                auto ei = byId.find(frag(ie->attrs_or("url")));
                if (ei != byId.end()) {
                    const Xml* prof = ei->second->child("profile_COMMON");
                    const Xml* tech = prof ? prof->child("technique") : nullptr;
                    const Xml* shader = nullptr;
                    if (tech) for (const char* n : {"phong", "lambert", "blinn", "constant"}) if ((shader = tech->child(n))) break;
                    if (shader) {
                        m.emissive = color(shader->child("emission"), m.emissive);
                        m.diffuse = color(shader->child("diffuse"), m.diffuse);
                        m.reflective = color(shader->child("reflective"), m.reflective);
                        if (const Xml* ior = shader->child("index_of_refraction")) if (const Xml* f = ior->child("float")) { auto v = floats(f->text); if (!v.empty()) m.refracti = v[0]; }
                        m.diffuseTex = texturePath(prof, shader->child("diffuse"));
                        const Xml* bump = ei->second->find("bump");      // <extra> FCOLLADA / OpenCOLLADA3dsMax bump slot -> aiTextureType_NORMALS
                        if (bump) m.normalsTex = texturePath(prof, bump);
                    }
                }
            }
        }
        materials.push_back(m);
        return materialIndex[materialId] = (int)materials.size() - 1;
    }

    void addGeometry(const Xml* geom, const M4& world, const std::map<std::string, std::string>& bind)
    {
        const Xml* mesh = geom->child("mesh");
        if (!mesh) return;
        std::map<std::string, Source> sources;
        for (const Xml* s : mesh->children("source")) {
            Source src;
            if (const Xml* fa = s->child("float_array")) src.data = floats(fa->text);
            if (const Xml* acc = s->find("accessor")) src.stride = std::max(1, atoi(acc->attrs_or("stride", "1").c_str()));
            sources[s->attrs_or("id")] = std::move(src);
        }
        struct Input { std::string semantic, source; int offset = 0; };
        std::map<std::string, std::vector<Input>> vertexInputs;
        for (const Xml* v : mesh->children("vertices")) {
            std::vector<Input> ins;
            for (const Xml* in : v->children("input")) ins.push_back({in->attrs_or("semantic"), frag(in->attrs_or("source")), 0});
            vertexInputs[v->attrs_or("id")] = ins;
        }
        const M4 nm = world.normalMatrix();
        for (auto& prim : mesh->kids) {
            const bool tris = prim->name == "triangles", plist = prim->name == "polylist", polys = prim->name == "polygons";
            if (!tris && !plist && !polys) continue;
            std::vector<Input> inputs; int stride = 0;
            for (const Xml* in : prim->children("input")) {
                Input i{in->attrs_or("semantic"), frag(in->attrs_or("source")), atoi(in->attrs_or("offset", "0").c_str())};
                if (i.offset < 0 || i.offset > 255) continue;        // malformed file: an index tuple has no such slot
                stride = std::max(stride, i.offset + 1);             // the slot exists in the index stream even if the input is not used
                const std::string set = in->attrs_or("set", "0");
                if (i.semantic == "TEXCOORD" && set != "0" && set != "1") continue;
                if (i.semantic == "VERTEX") { for (Input vi : vertexInputs[i.source]) { vi.offset = i.offset; inputs.push_back(vi); } }
                else inputs.push_back(i);
            }
            const Source *pos = nullptr, *nor = nullptr, *uv = nullptr; int posOff = 0, norOff = 0, uvOff = 0;
            for (auto& i : inputs) {
                auto it = sources.find(i.source);
                if (it == sources.end()) continue;
                if (i.semantic == "POSITION" && !pos) { pos = &it->second; posOff = i.offset; }
                else if (i.semantic == "NORMAL" && !nor) { nor = &it->second; norOff = i.offset; }
                else if (i.semantic == "TEXCOORD" && !uv) { uv = &it->second; uvOff = i.offset; }
            }
            if (!pos || stride == 0) continue;
            // index stream -> polygons
            std::vector<std::vector<long>> polygons;
            if (polys) { for (const Xml* p : prim->children("p")) polygons.push_back(ints(p->text)); }
            else {
                const Xml* p = prim->child("p");
                if (!p) continue;
                std::vector<long> idx = ints(p->text);
                std::vector<long> vc;
                if (plist) { if (const Xml* v = prim->child("vcount")) vc = ints(v->text); }
                size_t cur = 0;
                const size_t npoly = plist ? vc.size() : idx.size() / (3 * (size_t)stride);
                for (size_t k = 0; k < npoly; k++) {
                    if (plist && vc[k] < 0) break;                    // malformed vcount
                    const size_t nv = plist ? (size_t)vc[k] : 3;
                    if (nv > (idx.size() - cur) / (size_t)stride) break;   // (no overflow: the remaining indices cannot hold this polygon)
                    polygons.emplace_back(idx.begin() + cur, idx.begin() + cur + nv * stride);
                    cur += nv * stride;
                }
            }
            AiMesh out;
            std::string sym = prim->attrs_or("material");
            auto b = bind.find(sym);
            out.materialIndex = materialFor(b != bind.end() ? b->second : sym);
            std::map<std::vector<long>, int> unique;     // JoinIdenticalVertices by index tuple
            auto vertexOf = [&](const long* t) {
                std::vector<long> key = {t[posOff], nor ? t[norOff] : -1, uv ? t[uvOff] : -1};
                auto it = unique.find(key);
                if (it != unique.end()) return it->second;
                const size_t pi = (size_t)key[0] * pos->stride;
                V3 P{0, 0, 0};
                if (pi + 2 < pos->data.size()) P = {pos->data[pi], pos->data[pi + 1], pos->data[pi + 2]};
                out.vertices.push_back(world.point(P));
                if (nor) {
                    const size_t ni = (size_t)key[1] * nor->stride;
                    V3 N{0, 0, 0};
                    if (ni + 2 < nor->data.size()) N = {nor->data[ni], nor->data[ni + 1], nor->data[ni + 2]};
                    out.normals.push_back(normalizeSafe(nm.dir(N)));
                }
                if (uv) {
                    const size_t ui = (size_t)key[2] * uv->stride;
                    float u = 0, v = 0;
                    if (ui + 1 < uv->data.size()) { u = uv->data[ui]; v = uv->data[ui + 1]; }
                    out.uvs.push_back(u); out.uvs.push_back(v);
                }
                return unique[key] = (int)out.vertices.size() - 1;
            };
            for (auto& poly : polygons) {
                const size_t nv = poly.size() / stride;
                if (nv < 3) continue;                                 // points and lines are removed (Scene.cpp:91-92)
                std::vector<int> vi(nv);
                for (size_t k = 0; k < nv; k++) vi[k] = vertexOf(poly.data() + k * stride);
                for (size_t k = 1; k + 1 < nv; k++) { out.faces.push_back(vi[0]); out.faces.push_back(vi[k]); out.faces.push_back(vi[k + 1]); }   // Triangulate
            }
            if (out.faces.empty()) continue;
            if (out.normals.empty()) genSmoothNormals(out);
            if (!out.uvs.empty()) calcTangentSpace(out);
            meshes.push_back(std::move(out));
        }
    }

    // aiProcess_GenSmoothNormals: face normals (cross product, area weighted) summed over all vertices at the same position
    static void genSmoothNormals(AiMesh& m)
    {
        std::vector<V3> acc(m.vertices.size());
        for (size_t f = 0; f + 2 < m.faces.size(); f += 3) {
            const V3 n = cross(m.vertices[m.faces[f + 1]] - m.vertices[m.faces[f]], m.vertices[m.faces[f + 2]] - m.vertices[m.faces[f]]);
            for (int k = 0; k < 3; k++) acc[m.faces[f + k]] = acc[m.faces[f + k]] + n;
        }
        std::map<std::vector<float>, V3> byPos;
        for (size_t i = 0; i < m.vertices.size(); i++) { V3& s = byPos[{m.vertices[i].x, m.vertices[i].y, m.vertices[i].z}]; s = s + acc[i]; }
        m.normals.resize(m.vertices.size());
        for (size_t i = 0; i < m.vertices.size(); i++) m.normals[i] = normalizeSafe(byPos[{m.vertices[i].x, m.vertices[i].y, m.vertices[i].z}]);
    }

    // aiProcess_CalcTangentSpace: tangent / bitangent from the UV gradients of each face, made orthogonal to the vertex normal
    static void calcTangentSpace(AiMesh& m)
    {
        std::vector<V3> T(m.vertices.size()), B(m.vertices.size());
        for (size_t f = 0; f + 2 < m.faces.size(); f += 3) {
            const int i0 = m.faces[f], i1 = m.faces[f + 1], i2 = m.faces[f + 2];
            const V3 v = m.vertices[i1] - m.vertices[i0], w = m.vertices[i2] - m.vertices[i0];
            float sx = m.uvs[2 * i1] - m.uvs[2 * i0], sy = m.uvs[2 * i1 + 1] - m.uvs[2 * i0 + 1];
            float tx = m.uvs[2 * i2] - m.uvs[2 * i0], ty = m.uvs[2 * i2 + 1] - m.uvs[2 * i0 + 1];
            const float dirCorrection = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
            if (sx * ty == sy * tx) { sx = 0.0f; sy = 1.0f; tx = 1.0f; ty = 0.0f; }       // degenerate UVs: assimp's default basis
            const V3 tangent = (w * sy - v * ty) * dirCorrection, bitangent = (w * sx - v * tx) * dirCorrection;
            for (int k : {i0, i1, i2}) {
                const V3 n = m.normals[k];
                V3 lt = normalizeSafe(tangent - n * dot(tangent, n)), lb = normalizeSafe(bitangent - n * dot(bitangent, n));
                T[k] = T[k] + lt; B[k] = B[k] + lb;
            }
        }
        m.tangents.resize(m.vertices.size()); m.bitangents.resize(m.vertices.size());
        for (size_t i = 0; i < m.vertices.size(); i++) { m.tangents[i] = normalizeSafe(T[i]); m.bitangents[i] = normalizeSafe(B[i]); }
    }

    static M4 nodeTransform(const Xml* node)
    {
        M4 t = M4::identity();
        for (auto& c : node->kids) {
            auto f = floats(c->text);
            M4 s = M4::identity();
            if (c->name == "matrix" && f.size() >= 16) { for (int i = 0; i < 16; i++) s.m[i] = f[i]; }
            else if (c->name == "translate" && f.size() >= 3) { s.m[3] = f[0]; s.m[7] = f[1]; s.m[11] = f[2]; }
            else if (c->name == "scale" && f.size() >= 3) { s.m[0] = f[0]; s.m[5] = f[1]; s.m[10] = f[2]; }
            else if (c->name == "rotate" && f.size() >= 4) {
                V3 a = normalizeSafe({f[0], f[1], f[2]});
                const float ang = f[3] * (float)M_PI / 180.f, cs = cosf(ang), sn = sinf(ang), t1 = 1 - cs;
                s.m[0] = cs + a.x * a.x * t1; s.m[1] = a.x * a.y * t1 - a.z * sn; s.m[2] = a.x * a.z * t1 + a.y * sn;
                s.m[4] = a.y * a.x * t1 + a.z * sn; s.m[5] = cs + a.y * a.y * t1; s.m[6] = a.y * a.z * t1 - a.x * sn;
                s.m[8] = a.z * a.x * t1 - a.y * sn; s.m[9] = a.z * a.y * t1 + a.x * sn; s.m[10] = cs + a.z * a.z * t1;
            } else if (c->name == "lookat" && f.size() >= 9) {
                V3 eye{f[0], f[1], f[2]}, at{f[3], f[4], f[5]}, up = normalizeSafe({f[6], f[7], f[8]});
                V3 dir = normalizeSafe(at - eye), right = normalizeSafe(cross(dir, up));
                V3 u2 = cross(right, dir);
                s.m[0] = right.x; s.m[1] = u2.x; s.m[2] = -dir.x; s.m[3] = eye.x;
                s.m[4] = right.y; s.m[5] = u2.y; s.m[6] = -dir.y; s.m[7] = eye.y;
                s.m[8] = right.z; s.m[9] = u2.z; s.m[10] = -dir.z; s.m[11] = eye.z;
            } else continue;
            t = t * s;
        }
        return t;
    }

    void walk(const Xml* node, const M4& parent)
    {
        const M4 world = parent * nodeTransform(node);
        for (const Xml* ig : node->children("instance_geometry")) {
            auto gi = byId.find(frag(ig->attrs_or("url")));
            if (gi == byId.end()) continue;
            std::map<std::string, std::string> bind;
            if (const Xml* bm = ig->child("bind_material")) if (const Xml* tc = bm->child("technique_common"))
                for (const Xml* im : tc->children("instance_material")) bind[im->attrs_or("symbol")] = frag(im->attrs_or("target"));
            addGeometry(gi->second, world, bind);
        }
        for (const Xml* il : node->children("instance_light")) {
            auto li = byId.find(frag(il->attrs_or("url")));
            if (li == byId.end()) continue;
            const Xml* tc = li->second->child("technique_common");
            if (!tc) continue;
            AiLight l{}; l.position = world.point({0, 0, 0}); l.direction = normalizeSafe(world.dir({0, 0, -1}));
            if (const Xml* p = tc->child("point")) { l.type = 1; l.color = color(p, {1, 1, 1}); lights.push_back(l); }
            else if (const Xml* s = tc->child("spot")) {
                l.type = 2; l.color = color(s, {1, 1, 1});
                float fall = 180.f; if (const Xml* fa = s->child("falloff_angle")) { auto v = floats(fa->text); if (!v.empty()) fall = v[0]; }
                l.innerCone = fall * (float)M_PI / 180.f;
                lights.push_back(l);
            }
        }
        for (const Xml* ic : node->children("instance_camera")) {
            auto ci = byId.find(frag(ic->attrs_or("url")));
            if (ci == byId.end()) continue;
            AiCamera c; c.position = world.point({0, 0, 0}); c.lookAt = world.dir({0, 0, -1}); c.up = world.dir({0, 1, 0});
            if (const Xml* persp = ci->second->find("perspective")) {
                float xfov = 0, yfov = 0, aspect = 0;
                if (const Xml* e = persp->child("xfov")) { auto v = floats(e->text); if (!v.empty()) xfov = v[0]; }
                if (const Xml* e = persp->child("yfov")) { auto v = floats(e->text); if (!v.empty()) yfov = v[0]; }
                if (const Xml* e = persp->child("aspect_ratio")) { auto v = floats(e->text); if (!v.empty()) aspect = v[0]; }
                if (xfov == 0 && yfov != 0 && aspect != 0) xfov = 2.0f * atanf(aspect * tanf(yfov * (float)M_PI / 360.f)) * 180.f / (float)M_PI;
                c.hfovRad = xfov * (float)M_PI / 180.f;       // ColladaLoader: mHorizontalFOV = AI_DEG_TO_RAD(xfov)
            }
            cameras.push_back(c);
        }
        for (const Xml* child : node->children("node")) walk(child, world);
    }

    void load(const std::string& path)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("The file that was supplied (" + path + ") does not exist.");     // Scene.cpp:75-79
        std::string src; char buf[1 << 16]; size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) src.append(buf, n);
        fclose(f);
        std::unique_ptr<Xml> doc;
        try { XmlParser p(src); doc = p.parseDocument(); }
        catch (const std::exception& e) { throw std::runtime_error(std::string("An error occurred in Assimp during reading of this file: ") + e.what()); }
        if (doc->name != "COLLADA") throw std::runtime_error("An error occurred in Assimp during reading of this file: not a COLLADA document");
        root = doc.get();
        index(root);
        // every material of <library_materials> gets an index, in document order (assimp's material array)
        if (const Xml* lm = root->child("library_materials")) for (const Xml* m : lm->children("material")) materialFor(m->attrs_or("id"));
        M4 up = M4::identity();
        if (const Xml* asset = root->child("asset")) if (const Xml* ua = asset->child("up_axis")) {
            std::string a = ua->text;
            if (a.find("Z_UP") != std::string::npos) { up = M4::identity(); up.m[5] = 0; up.m[6] = 1; up.m[9] = -1; up.m[10] = 0; }      // (x, y, z) -> (x, z, -y)
            else if (a.find("X_UP") != std::string::npos) { up = M4::identity(); up.m[0] = 0; up.m[1] = -1; up.m[4] = 1; up.m[5] = 0; }   // (x, y, z) -> (-y, x, z)
        }
        const Xml* vs = nullptr;
        if (const Xml* sc = root->child("scene")) if (const Xml* iv = sc->child("instance_visual_scene")) { auto it = byId.find(frag(iv->attrs_or("url"))); if (it != byId.end()) vs = it->second; }
        if (!vs) if (const Xml* lvs = root->child("library_visual_scenes")) vs = lvs->child("visual_scene");
        if (!vs) throw std::runtime_error("An error occurred in Assimp during reading of this file: no visual scene");
        for (const Xml* node : vs->children("node")) walk(node, up);
        if (materials.empty()) materialFor("");
        root = nullptr;
    }
};

// ---- Scene.cpp on top of the imported data ---------------------------------------------------------------------------------
struct FileScene {
    std::string name, dir;
    Importer imp;
    std::vector<opr_material> materials; std::vector<opr_mesh> meshes; std::vector<opr_light> lights; std::vector<opr_texture> textures;
    std::vector<std::vector<uint8_t>> texels; std::vector<std::vector<float>> fbuf; std::vector<std::vector<int32_t>> ibuf;
    opr_scene_desc desc{}; opr_camera camera{}; bool hasCamera = false;
    unsigned numTriangles = 0;
    std::string warnings;

    static bool any(V3 c) { return c.x > 0 || c.y > 0 || c.z > 0; }     // colorHasAnyComponent, Scene.cpp:44-47

    int loadTexture(const std::string& rel)
    {
        const std::string path = dir + "/" + rel;                          // Scene.cpp:209
        int32_t w = 0, h = 0;
        if (opr_image_info(path.c_str(), &w, &h)) throw std::runtime_error(opr_io_last_error());
        texels.emplace_back((size_t)w * h * 4);
        if (opr_image_read(path.c_str(), texels.back().data(), texels.back().size())) throw std::runtime_error(opr_io_last_error());
        opr_texture t{}; t.width = w; t.height = h; t.rgba = texels.back().data();
        textures.push_back(t);
        return (int)textures.size() - 1;
    }

    void build(const char* path, int flags)
    {
        name = path;
        char* abs = realpath(path, nullptr);
        if (abs) { name = abs; free(abs); }
        dir = name.find('/') == std::string::npos ? "." : name.substr(0, name.rfind('/'));
        imp.load(path);
        // ---- loadSceneMaterials, Scene.cpp:178-265: emissive -> textured -> glass -> mirror -> diffuse (-> red fallback)
        for (const AiMaterial& am : imp.materials) {
            opr_material m{}; m.diffuseTexture = m.normalTexture = -1;
            auto set3 = [](float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
            if (any(am.emissive)) { m.kind = OPR_MAT_DIFFUSE_EMITTER; set3(m.power, am.emissive); set3(m.Kd, am.diffuse); }        // DiffuseEmitter(power, Kd)
            else if (!am.diffuseTex.empty()) {
                m.kind = OPR_MAT_TEXTURE; m.diffuseTexture = loadTexture(am.diffuseTex);
                if (!am.normalsTex.empty()) m.normalTexture = loadTexture(am.normalsTex);
            }
            else if (am.refracti > 1.0f) { m.kind = OPR_MAT_GLASS; m.indexOfRefraction = am.refracti; set3(m.Kr, {1, 1, 1}); set3(m.Kt, {1, 1, 1}); }
            else if (any(am.reflective)) { m.kind = OPR_MAT_MIRROR; set3(m.Kr, am.reflective); }
            else { m.kind = OPR_MAT_DIFFUSE; set3(m.Kd, am.diffuse); }
            materials.push_back(m);
        }
        // ---- loadLightSources, Scene.cpp:267-285
        for (const AiLight& l : imp.lights) {
            opr_light o{};
            const float p[3] = {l.color.x, l.color.y, l.color.z}, pos[3] = {l.position.x, l.position.y, l.position.z};
            if (l.type == 1) opr_make_point_light(&o, p, pos);
            else {   // Light::Light(power, position, direction, angle), renderer/Light.cpp:42-49
                opr_make_point_light(&o, p, pos);
                o.lightType = OPR_LIGHT_SPOT; o.normal[0] = l.direction.x; o.normal[1] = l.direction.y; o.normal[2] = l.direction.z; o.area_or_angle = l.innerCone;
            }
            lights.push_back(o);
        }
        // ---- meshes: emitters become quad lights (loadMeshLightSource, Scene.cpp:287-311); AABB quirk (:147-160)
        V3 mn{1e33f, 1e33f, 1e33f}, mx{-1e33f, -1e33f, -1e33f};
        std::vector<opr_light> emitterLights;
        for (AiMesh& am : imp.meshes) {
            opr_material& mat = materials[am.materialIndex];
            if (mat.kind == OPR_MAT_DIFFUSE_EMITTER) {
                const unsigned nf = (unsigned)am.faces.size() / 3;
                if (nf < 1 || nf > 2) warnings += "Material " + imp.materials[am.materialIndex].name + ": Does only support quadrangle light source NumFaces: " + std::to_string(nf) + ".\n";
                const V3 anchor = am.vertices[am.faces[0]], p1 = am.vertices[am.faces[1]], p2 = am.vertices[am.faces[2]];
                const V3 v1 = p1 - anchor, v2 = p2 - anchor;
                opr_light li{};
                const float a[3] = {anchor.x, anchor.y, anchor.z}, e1[3] = {v1.x, v1.y, v1.z}, e2[3] = {v2.x, v2.y, v2.z};
                opr_make_area_light(&li, mat.power, a, e1, e2);
                emitterLights.push_back(li);
                mat.inverseArea = li.inverseArea;                          // DiffuseEmitter::setInverseArea
            }
            numTriangles += (unsigned)am.faces.size() / 3;
            for (size_t f = 0; f + 2 < am.faces.size(); f += 3) {
                const V3 p = am.vertices[am.faces[f]];                     // face.mIndices[0] three times
                mn = {std::min(mn.x, p.x), std::min(mn.y, p.y), std::min(mn.z, p.z)};
                mx = {std::max(mx.x, p.x), std::max(mx.y, p.y), std::max(mx.z, p.z)};
            }
            opr_mesh om{};
            auto pack3 = [&](const std::vector<V3>& v) -> const float* {
                if (v.empty()) return nullptr;
                fbuf.emplace_back(); fbuf.back().reserve(v.size() * 3);
                for (const V3& p : v) { fbuf.back().push_back(p.x); fbuf.back().push_back(p.y); fbuf.back().push_back(p.z); }
                return fbuf.back().data();
            };
            om.numVertices = (int)am.vertices.size(); om.numTriangles = (int)am.faces.size() / 3;
            om.vertices = pack3(am.vertices); om.normals = pack3(am.normals); om.tangents = pack3(am.tangents); om.bitangents = pack3(am.bitangents);
            if (!am.uvs.empty()) { fbuf.push_back(am.uvs); om.uvs = fbuf.back().data(); }
            ibuf.emplace_back(am.faces.begin(), am.faces.end()); om.indices = ibuf.back().data();
            om.material = am.materialIndex;
            meshes.push_back(om);
        }
        // the reference appends every emitter light in createFromFile (:141) and again while it builds the scene graph (:508);
        // the renderer reads the list after both (OptixRenderer.cpp:463-473)
        const int copies = (flags & 1) ? 1 : 2;
        for (int c = 0; c < copies; c++) lights.insert(lights.end(), emitterLights.begin(), emitterLights.end());
        if (!imp.cameras.empty()) {   // loadDefaultSceneCamera, Scene.cpp:469-485
            const AiCamera& c = imp.cameras[0];
            const V3 eye = c.position, lookAt = eye + c.lookAt, up = normalizeSafe(c.up);
            memset(&camera, 0, sizeof camera);
            camera.eye[0] = eye.x; camera.eye[1] = eye.y; camera.eye[2] = eye.z;
            camera.lookat[0] = lookAt.x; camera.lookat[1] = lookAt.y; camera.lookat[2] = lookAt.z;
            camera.up[0] = up.x; camera.up[1] = up.y; camera.up[2] = up.z;
            camera.hfov = camera.vfov = c.hfovRad * 365.0f / (2.0f * (float)M_PI);
            camera.aperture = 0.f; camera.aspectRatioMode = OPR_KEEP_HORIZONTAL;
            opr_camera_setup(&camera);
            hasCamera = true;
        }
        desc.name = name.c_str();
        desc.numMeshes = (int)meshes.size(); desc.meshes = meshes.data();
        desc.numMaterials = (int)materials.size(); desc.materials = materials.data();
        desc.numTextures = (int)textures.size(); desc.textures = textures.data();
        desc.numLights = (int)lights.size(); desc.lights = lights.data();
        desc.aabbMin[0] = mn.x; desc.aabbMin[1] = mn.y; desc.aabbMin[2] = mn.z;
        desc.aabbMax[0] = mx.x; desc.aabbMax[1] = mx.y; desc.aabbMax[2] = mx.z;
    }
};

}  // namespace

extern "C" {

const char* opr_scene_file_last_error(void) { return g_importError.c_str(); }

// Scene::createFromFile (scene/Scene.cpp:73-176).  flags bit 0: append emitter lights once instead of the reference's twice.
void* opr_scene_file_load(const char* path, int32_t flags)
{
    try {
        std::unique_ptr<FileScene> s(new FileScene);
        s->build(path, flags);
        return s.release();
    } catch (const std::exception& e) {
        g_importError = e.what();
        return nullptr;
    }
}
void opr_scene_file_destroy(void* scene) { delete static_cast<FileScene*>(scene); }
const opr_scene_desc* opr_scene_file_desc(void* scene) { return &static_cast<FileScene*>(scene)->desc; }
int opr_scene_file_camera(void* scene, opr_camera* out)
{
    FileScene* s = static_cast<FileScene*>(scene);
    if (!s->hasCamera) return 1;
    *out = s->camera;
    return 0;
}
uint32_t opr_scene_file_num_triangles(void* scene) { return static_cast<FileScene*>(scene)->numTriangles; }
uint32_t opr_scene_file_num_meshes(void* scene) { return (uint32_t)static_cast<FileScene*>(scene)->meshes.size(); }
const char* opr_scene_file_warnings(void* scene) { return static_cast<FileScene*>(scene)->warnings.c_str(); }

}  // extern "C"
