// vxl_assets.cu -- the reference's on-disk formats.  Host code only: no kernel lives here.
//
//   .v   VoxAsset::Serialize        Sources/Asset/VoxAsset.h:42-50      int32 SizeX, SizeY, SizeZ + SizeX*SizeY*SizeZ bytes (x fastest)
//   .p   PalleteAsset::Serialize    Sources/Asset/PalleteAsset.h:63-69  256 x VoxMaterial {r, g, b, a, roughness, metallic, emit} (7 bytes)
//        PalleteCache::UploadPallete Sources/Vox/PalleteCache.cpp:5-25  -> colour texel (r, g, b, 255), material texel (roughness, metallic, emit, 0)
//   GUID Assets::Hash               Sources/Asset/Assets.h:207-210      std::hash<std::string> of the path relative to Mods/ -- FNV-1a 64 under
//                                                                        MSVC (the reference's only platform); the GUIDs stored in the shipped
//                                                                        prefabs confirm it ("default/ModernHouse/0.v" -> 44B7A418296B6797)
//   .pf  PrefabAsset::Spawn         Sources/Asset/PrefabAsset.cpp:30-141 JSON array of entities: Id, Parent, Name, Transform, VoxRenderer, Light
//        TransformSystem::RealculateMatrix  Sources/World/Systems/TransformSystem.cpp:124-135: Matrix = T * Rz * Ry * Rx * S, World = Parent * Matrix,
//        with the arithmetic of glm 0.9.9.9's translate / rotate / scale (ext/matrix_transform.inl), entities in file order (parents first).
// Components the path does not consume (IKChain, Script, Character) are skipped.  Nested prefab instances ("Instance": GUID) are
// expanded by vxl_scene_load, which resolves GUIDs like ModLoader does (hash of every file path under the Mods directory).
#include "vxl_internal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace vxl;

namespace {

// ---- a small JSON reader (the reference uses json11; only reading is needed) -----------------------------------------
struct JValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const char* key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};
struct JParser {
    const char* p; const char* end; std::string err;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    bool parse_string(std::string& out) {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': {
                        if (end - p < 5) return fail("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += *p;
                }
                ++p;
            } else out += *p++;
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool parse(JValue& v, int depth = 0) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.kind = JValue::Object; ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string key;
                if (!parse_string(key)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                JValue child;
                if (!parse(child, depth + 1)) return false;
                v.obj.emplace_back(std::move(key), std::move(child));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JValue::Array; ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                JValue child;
                if (!parse(child, depth + 1)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.kind = JValue::String; return parse_string(v.str); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; return true; }
        char* e = nullptr;
        std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 64));
        v.num = strtod(tmp.c_str(), &e);
        if (e == tmp.c_str()) return fail("unexpected character");
        v.kind = JValue::Number; p += e - tmp.c_str();
        return true;
    }
};

bool read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); err = "ftell failed"; return false; }
    out.resize((size_t)n);
    const size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) { err = std::string("short read on ") + path; return false; }
    return true;
}

// readVec3 (PrefabAsset.cpp:8-18): three decimal floats separated by single spaces; from_chars == correctly rounded == strtof
void read_vec3(const std::string& s, float* v) {
    const size_t s1 = s.find(' '), s2 = s1 == std::string::npos ? std::string::npos : s.find(' ', s1 + 1);
    v[0] = strtof(s.c_str(), nullptr);
    v[1] = s1 == std::string::npos ? 0.0f : strtof(s.c_str() + s1 + 1, nullptr);
    v[2] = s2 == std::string::npos ? 0.0f : strtof(s.c_str() + s2 + 1, nullptr);
}

// ---- glm 0.9.9.9 ext/matrix_transform.inl, restated on float[16] column-major ------------------------------------------
typedef float M4[16];
void m4_identity(float* m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }
// translate :6-11: Result[3] = m[0]*v[0] + m[1]*v[1] + m[2]*v[2] + m[3]
void m4_translate(float* m, const float* v) {
    for (int r = 0; r < 4; ++r) m[12 + r] = ((m[0 + r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r];
}
// rotate :14-43 (axis is one of the unit axes; normalize() leaves it unchanged)
void m4_rotate(float* m, float angle, int axis_index) {
    const float c = std::cos(angle), s = std::sin(angle);
    float axis[3] = {0.0f, 0.0f, 0.0f};
    axis[axis_index] = 1.0f;
    const float inv = 1.0f / std::sqrt((axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2]);   // glm::normalize
    for (int i = 0; i < 3; ++i) axis[i] = axis[i] * inv;
    float temp[3];
    for (int i = 0; i < 3; ++i) temp[i] = (1.0f - c) * axis[i];
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0]; R[0][1] = temp[0] * axis[1] + s * axis[2]; R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2]; R[1][1] = c + temp[1] * axis[1]; R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1]; R[2][1] = temp[2] * axis[1] - s * axis[0]; R[2][2] = c + temp[2] * axis[2];
    float out[12];
    for (int j = 0; j < 3; ++j)
        for (int r = 0; r < 4; ++r) out[j * 4 + r] = (m[0 + r] * R[j][0] + m[4 + r] * R[j][1]) + m[8 + r] * R[j][2];
    memcpy(m, out, sizeof out);
}
// scale :46-54
void m4_scale(float* m, const float* v) {
    for (int j = 0; j < 3; ++j)
        for (int r = 0; r < 4; ++r) m[j * 4 + r] = m[j * 4 + r] * v[j];
}
// mat4 * mat4 (type_mat4x4.inl:630-646)
void m4_mul(const float* A, const float* B, float* R) {
    float out[16];
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            out[j * 4 + r] = ((A[0 + r] * B[j * 4 + 0] + A[4 + r] * B[j * 4 + 1]) + A[8 + r] * B[j * 4 + 2]) + A[12 + r] * B[j * 4 + 3];
    memcpy(R, out, sizeof out);
}

// One entity's own fields from its JSON object (PrefabAsset.cpp:87-139); `o` keeps what it already holds for absent components.
void apply_components(const JValue& e, vxl_prefab_entity& o) {
    if (const JValue* nm = e.get("Name")) if (nm->kind == JValue::String) { memset(o.name, 0, sizeof o.name); strncpy(o.name, nm->str.c_str(), sizeof o.name - 1); }
    if (const JValue* t = e.get("Transform")) {
        o.has |= VXL_PF_TRANSFORM;
        for (int i = 0; i < 3; ++i) { o.position[i] = 0.0f; o.rotation[i] = 0.0f; o.scale[i] = 1.0f; }          // a fresh Transform replaces the old one
        if (const JValue* v = t->get("Position")) read_vec3(v->str, o.position);
        if (const JValue* v = t->get("Rotation")) read_vec3(v->str, o.rotation);
        if (const JValue* v = t->get("Scale")) read_vec3(v->str, o.scale);
    }
    if (const JValue* r = e.get("VoxRenderer")) {
        o.has |= VXL_PF_VOX;
        o.vox_guid = o.pallete_guid = 0; o.pivot[0] = o.pivot[1] = o.pivot[2] = 0.0f;
        if (const JValue* v = r->get("Pallete")) o.pallete_guid = strtoull(v->str.c_str(), nullptr, 16);
        if (const JValue* v = r->get("Vox")) o.vox_guid = strtoull(v->str.c_str(), nullptr, 16);
        if (const JValue* v = r->get("Pivot")) if (v->kind == JValue::String) read_vec3(v->str, o.pivot);
    }
    if (const JValue* l = e.get("Light")) {
        o.has |= VXL_PF_LIGHT;
        auto num = [&](const char* k) { const JValue* v = l->get(k); return v && v->kind == JValue::Number ? (float)v->num : 0.0f; };
        // PrefabAsset.cpp:111-119 assigns every field from the JSON (a missing number reads 0)
        o.light_type = (int)num("LightType"); o.intensity = num("Intensity");
        if (const JValue* v = l->get("Color")) read_vec3(v->str, o.color);
        o.attenuation = num("Attenuation"); o.range = num("Range"); o.angle = num("Angle"); o.angle_attenuation = num("AngleAttenuation");
    }
}

// PrefabAsset::Spawn (PrefabAsset.cpp:30-141): appends the prefab's entities to `ents` in file order under `parent_index`.
// paths: GUID -> path relative to the Mods directory (ModLoader's table) for nested instances; nullptr = instances are an error.
int spawn_prefab(const std::string& file, int parent_index, const std::map<uint64_t, std::string>* paths, const std::string& mods_dir,
                 std::vector<vxl_prefab_entity>& ents, int* root_out, int depth) {
    if (depth > 16) { set_error("prefab instances nest deeper than 16 levels (cycle?)"); return VXL_ERR_LIMIT; }
    std::vector<uint8_t> buf;
    std::string err;
    if (!read_file(file.c_str(), buf, err)) { set_error("prefab: " + err); return VXL_ERR_INVALID; }
    JParser P{(const char*)buf.data(), (const char*)buf.data() + buf.size(), ""};
    JValue root;
    if (!P.parse(root) || root.kind != JValue::Array) { set_error("prefab: " + file + " is not a JSON array of entities (" + P.err + ")"); return VXL_ERR_INVALID; }
    std::map<int, int> index_of;                                 // "Id" -> index in `ents` (newEntityMap, PrefabAsset.cpp:34)
    *root_out = -1;
    for (const JValue& e : root.arr) {
        if (e.kind != JValue::Object) { set_error("prefab: entity is not an object"); return VXL_ERR_INVALID; }
        int parent = -1;
        if (const JValue* par = e.get("Parent")) {
            if (par->kind == JValue::Number) {
                auto it = index_of.find((int)par->num);
                if (it == index_of.end()) { set_error("prefab: Parent refers to an entity that does not precede it"); return VXL_ERR_INVALID; }
                parent = it->second;
            }
        }
        const bool is_root = parent < 0;
        if (is_root) parent = parent_index;
        const JValue* id = e.get("Id");
        const int eid = id && id->kind == JValue::Number ? (int)id->num : 0;
        int idx;
        const JValue* inst = e.get("Instance");
        if (inst && inst->kind != JValue::Null) {                // :47-56: the nested prefab's root stands for this entity
            if (!paths) { set_error("vxl_prefab_file_read: nested prefab instances need vxl_scene_load (a Mods directory to resolve GUIDs)"); return VXL_ERR_INVALID; }
            const uint64_t g = strtoull(inst->str.c_str(), nullptr, 16);
            auto it = paths->find(g);
            if (it == paths->end()) { set_error("prefab: instance GUID " + inst->str + " is not a file under the Mods directory"); return VXL_ERR_INVALID; }
            int nested_root = -1;
            if (int rc = spawn_prefab((std::filesystem::path(mods_dir) / it->second).string(), parent, paths, mods_dir, ents, &nested_root, depth + 1)) return rc;
            if (nested_root < 0) { set_error("prefab: instance " + it->second + " has no root entity"); return VXL_ERR_INVALID; }
            idx = nested_root;
            ents[(size_t)idx].has |= VXL_PF_INSTANCE;
            ents[(size_t)idx].instance_guid = g;
            ents[(size_t)idx].parent = parent;
        } else {
            vxl_prefab_entity o;
            memset(&o, 0, sizeof o);
            o.parent = parent;
            o.scale[0] = o.scale[1] = o.scale[2] = 1.0f;                                        // Components.h:60-67 defaults
            o.light_type = 0; o.intensity = 2.0f; o.color[0] = o.color[1] = o.color[2] = 1.0f;   // Components.h:37-57 defaults
            o.attenuation = 2.0f; o.range = 10.0f; o.angle = 0.3f; o.angle_attenuation = 1.0f;
            ents.push_back(o);
            idx = (int)ents.size() - 1;
        }
        ents[(size_t)idx].id = eid;
        index_of[eid] = idx;
        if (is_root) *root_out = idx;
        apply_components(e, ents[(size_t)idx]);
    }
    return VXL_OK;
}

// TransformSystem::RealculateMatrix (TransformSystem.cpp:124-135) over the finished list (parents precede their children)
int finish_prefab(std::vector<vxl_prefab_entity>& ents, vxl_prefab_entity* out, int cap, int* n_out) {
    *n_out = (int)ents.size();
    if (cap == 0) return VXL_OK;                                 // count query
    if (cap < *n_out) { set_error("prefab: output array too small"); return VXL_ERR_LIMIT; }
    for (size_t i = 0; i < ents.size(); ++i) {
        vxl_prefab_entity& o = ents[i];
        m4_identity(o.matrix);
        m4_translate(o.matrix, o.position);
        m4_rotate(o.matrix, o.rotation[2], 2);
        m4_rotate(o.matrix, o.rotation[1], 1);
        m4_rotate(o.matrix, o.rotation[0], 0);
        m4_scale(o.matrix, o.scale);
        if (o.parent >= 0) m4_mul(ents[(size_t)o.parent].world, o.matrix, o.world);
        else { float I[16]; m4_identity(I); m4_mul(I, o.matrix, o.world); }
        out[i] = o;
    }
    return VXL_OK;
}

}  // namespace

extern "C" {

int vxl_asset_guid(const char* path, uint64_t* out) {
    if (!path || !out) { set_error("vxl_asset_guid: bad argument"); return VXL_ERR_INVALID; }
    uint64_t h = 14695981039346656037ull;                        // FNV-1a 64: offset basis, prime 1099511628211
    for (const unsigned char* p = (const unsigned char*)path; *p; ++p) { h ^= *p; h *= 1099511628211ull; }
    *out = h;
    return VXL_OK;
}

int vxl_vox_file_read(const char* path, int32_t dims[3], uint8_t* out, uint64_t cap) {
    if (!path || !dims) { set_error("vxl_vox_file_read: bad argument"); return VXL_ERR_INVALID; }
    std::vector<uint8_t> buf;
    std::string err;
    if (!read_file(path, buf, err)) { set_error("vxl_vox_file_read: " + err); return VXL_ERR_INVALID; }
    if (buf.size() < 12) { set_error("vxl_vox_file_read: file shorter than its header"); return VXL_ERR_INVALID; }
    memcpy(dims, buf.data(), 12);
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || dims[0] > 4096 || dims[1] > 4096 || dims[2] > 4096) { set_error("vxl_vox_file_read: bad dimensions"); return VXL_ERR_INVALID; }
    const uint64_t n = (uint64_t)dims[0] * (uint64_t)dims[1] * (uint64_t)dims[2];
    if (buf.size() - 12 < n) { set_error("vxl_vox_file_read: file shorter than SizeX*SizeY*SizeZ"); return VXL_ERR_INVALID; }
    if (!out) return VXL_OK;                                     // size query
    if (cap < n) { set_error("vxl_vox_file_read: output buffer too small"); return VXL_ERR_LIMIT; }
    memcpy(out, buf.data() + 12, (size_t)n);
    return VXL_OK;
}

int vxl_model_load_v(vxl_ctx* ctx, const char* path, int* out_model_id) {
    if (!ctx || !out_model_id) { set_error("vxl_model_load_v: bad argument"); return VXL_ERR_INVALID; }
    int32_t dims[3];
    if (int e = vxl_vox_file_read(path, dims, nullptr, 0)) return e;
    std::vector<uint8_t> vox((size_t)dims[0] * dims[1] * dims[2]);
    if (int e = vxl_vox_file_read(path, dims, vox.data(), vox.size())) return e;
    return vxl_model_create(ctx, vox.data(), dims[0], dims[1], dims[2], out_model_id);
}

int vxl_pallete_file_read(const char* path, uint32_t* color256, uint32_t* material256) {
    if (!path || !color256 || !material256) { set_error("vxl_pallete_file_read: bad argument"); return VXL_ERR_INVALID; }
    std::vector<uint8_t> buf;
    std::string err;
    if (!read_file(path, buf, err)) { set_error("vxl_pallete_file_read: " + err); return VXL_ERR_INVALID; }
    if (buf.size() < 256 * 7) { set_error("vxl_pallete_file_read: file shorter than 256 VoxMaterial records"); return VXL_ERR_INVALID; }
    for (int i = 0; i < 256; ++i) {
        const uint8_t* m = buf.data() + (size_t)i * 7;           // r g b a roughness metallic emit
        color256[i] = (uint32_t)m[0] | ((uint32_t)m[1] << 8) | ((uint32_t)m[2] << 16) | (255u << 24);
        material256[i] = (uint32_t)m[4] | ((uint32_t)m[5] << 8) | ((uint32_t)m[6] << 16);
    }
    return VXL_OK;
}

int vxl_prefab_file_read(const char* path, vxl_prefab_entity* out, int cap, int* n_out) {
    if (!path || !n_out || cap < 0 || (cap > 0 && !out)) { set_error("vxl_prefab_file_read: bad argument"); return VXL_ERR_INVALID; }
    std::vector<vxl_prefab_entity> ents;
    int root = -1;
    if (int e = spawn_prefab(path, -1, nullptr, "", ents, &root, 0)) return e;
    return finish_prefab(ents, out, cap, n_out);
}

int vxl_scene_load(const char* mods_dir, const char* prefab_path, vxl_prefab_entity* out, int cap, int* n_out) {
    if (!mods_dir || !prefab_path || !n_out || cap < 0 || (cap > 0 && !out)) { set_error("vxl_scene_load: bad argument"); return VXL_ERR_INVALID; }
    std::map<uint64_t, std::string> paths;
    std::error_code ec;
    const std::filesystem::path base(mods_dir);
    for (std::filesystem::recursive_directory_iterator it(base, ec), end; !ec && it != end; it.increment(ec)) {
        if (!it->is_regular_file(ec)) continue;
        const std::string rel = std::filesystem::relative(it->path(), base, ec).generic_string();
        uint64_t g;
        vxl_asset_guid(rel.c_str(), &g);
        paths[g] = rel;
    }
    if (ec) { set_error(std::string("vxl_scene_load: cannot walk ") + mods_dir); return VXL_ERR_INVALID; }
    std::vector<vxl_prefab_entity> ents;
    int root = -1;
    if (int e = spawn_prefab((base / prefab_path).string().c_str(), -1, &paths, base.string(), ents, &root, 0)) return e;
    return finish_prefab(ents, out, cap, n_out);
}

}  // extern "C"
