// vxl_voximport.cu -- the reference's MagicaVoxel .vox importer (last item).  Host code only: no kernel lives here.
//
//   VoxImportContext::Import   Sources/Editor/Importer/VoxImporter.cpp:284-394   chunk loop (MAIN PACK SIZE XYZI RGBA MATL LAYR IMAP rOBJ nTRN nGRP nSHP)
//   VoxTransformMatrix         :37-84    the `_r` byte: axis map + sign bits of a signed permutation
//   readMATL                   :159-203  _type -> (emit, roughness, metallic) * 255 truncated to uint8
//   readnTRN                   :209-255  `_t` "x y z" and `_r` by std::from_chars
//   CreateEntity               :397-476  node tree -> entities; shape voxels re-oriented (z up -> y up) into a VoxAsset whose sizes are
//                                        rounded up to a multiple of 4 (Sources/Asset/VoxAsset.h:26-30)
//   VoxImporter::Import        :478-520  palette records, asset paths <path>/<file>/<shape>.v, <path>/<file>/<file>.p, <path>/<file>.pf
//   PrefabAsset::FromWorld     Sources/Asset/PrefabAsset.cpp:142-241 + json11's dump + fmt's "{}" of a float: the .pf text
//
// The reference ships three .vox files next to the .v / .p / .pf files its importer wrote from them (Assets/Mods/default); the tests
// compare this code with those files byte for byte.  Chunks are read field by field like the reference does (chunk sizes are not used
// to skip), so a file the reference would misread is misread the same way -- but every read is bounded and malformed counts are errors.
#include "vxl_internal.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

using namespace vxl;

struct vxl_vox_scene {
    struct Model { std::string name; int32_t dims[3]; std::vector<uint8_t> data; };
    std::vector<Model> models;
    std::vector<vxl_vox_import_entity> entities;
    uint8_t records[256 * 7];                                    // .p file contents: r g b a roughness metallic emit
};

namespace {

struct Reader {                                                  // FileReader: a read past the end leaves the destination as it was
    const uint8_t* p; size_t n, pos = 0; bool bad = false;
    void take(void* dst, size_t k) {
        const size_t got = pos < n ? std::min(k, n - pos) : 0;
        if (got) memcpy(dst, p + pos, got);
        pos += k;
    }
    int32_t i32() { int32_t v = 0; take(&v, 4); return v; }
    std::string str() {
        const int32_t k = i32();
        if (k < 0 || (size_t)k > n) { bad = true; return std::string(); }
        std::string s((size_t)k, '\0');
        take(s.data(), (size_t)k);
        return s;
    }
    std::map<std::string, std::string> dict() {
        std::map<std::string, std::string> m;
        const int32_t k = i32();
        if (k < 0 || (size_t)k > n) { bad = true; return m; }
        for (int32_t i = 0; i < k && !bad; ++i) { std::string key = str(), val = str(); m[key] = val; }
        return m;
    }
};

struct Matrix { int rx, ry, rz, sx, sy, sz; };
Matrix matrix_of(uint8_t r) {                                    // VoxTransformMatrix(uint8) :45-58
    Matrix m;
    m.rx = r & 3; m.ry = (r >> 2) & 3; m.rz = 3 - (m.rx | m.ry);
    m.sx = (r >> 4) & 1; m.sy = (r >> 5) & 1; m.sz = (r >> 6) & 1;
    return m;
}

struct Node { char kind = 0; std::string name; int child = 0, t[3] = {0, 0, 0}; Matrix m{}; std::vector<int> children; int shape = 0; };
struct Shape { int size[3]; std::vector<uint32_t> voxels; };

void from_chars_range(const std::string& s, long b, long e, int& out) {      // std::from_chars(str.data() + b, str.data() + e, out)
    if (b < 0 || e > (long)s.size() || b >= e) return;
    std::from_chars(s.data() + b, s.data() + e, out);
}
uint8_t u8_of(float f) {                                         // (uint8)(float): the low byte of the truncated integer (cvttss2si)
    if (!(f > -2147483648.0f && f < 2147483648.0f)) return 0;
    return (uint8_t)((int32_t)f & 0xFF);
}

struct Importer {
    std::vector<Node> nodes;
    std::vector<Shape> shapes;
    uint8_t pallete[257][4];
    uint8_t surfaces[257][3];                                    // e, r, m
    int size[3] = {0, 0, 0};
    std::string err;

    bool parse(const uint8_t* data, size_t n) {
        memset(pallete, 0, sizeof pallete);                      // uninitialised in the reference; zero here (and in the tests' checker)
        memset(surfaces, 0, sizeof surfaces);
        Reader s{data, n};
        char magic[4] = {0, 0, 0, 0};
        s.take(magic, 4);
        if (memcmp(magic, "VOX ", 4)) { err = "not a .vox file (missing 'VOX ' magic)"; return false; }
        s.i32();                                                 // version
        std::string node_name;
        for (;;) {
            char h[4] = {' ', ' ', ' ', ' '};
            s.take(h, 4);
            s.i32(); s.i32();                                    // chunk content size, children size
            bool running = true;
            switch (h[0]) {
            case 'M':
                if (h[2] == 'T') {                               // MATL :159-203
                    const int32_t id = s.i32();
                    auto p = s.dict();
                    float rough = 0.0f, emit = 0.0f, metal = 0.0f;
                    const std::string type = p["_type"];
                    auto num = [&](const char* k, float& v) { auto it = p.find(k); if (it != p.end()) std::from_chars(it->second.data(), it->second.data() + it->second.size(), v); };
                    if (type == "_metal" || type == "_blend") { num("_rough", rough); num("_metal", metal); }
                    else if (type == "_emit") num("_emit", emit);
                    else if (type == "_diffuse") rough = 0.9f;
                    if (id < 0 || id > 256) { err = "MATL id out of range"; return false; }
                    surfaces[id][0] = u8_of(emit * 255.0f); surfaces[id][1] = u8_of(rough * 255.0f); surfaces[id][2] = u8_of(metal * 255.0f);
                }
                break;                                           // MAIN: nothing
            case 'P': s.i32(); break;
            case 'S': size[0] = s.i32(); size[1] = s.i32(); size[2] = s.i32(); break;
            case 'X': {                                          // XYZI :135-157
                const int32_t count = s.i32();
                if (count < 0 || (size_t)count > n / 4) { err = "XYZI voxel count exceeds the file"; return false; }
                Shape sh;
                sh.size[0] = size[0]; sh.size[1] = size[1]; sh.size[2] = size[2];
                sh.voxels.assign((size_t)count, 0u);
                s.take(sh.voxels.data(), (size_t)count * 4);
                shapes.push_back(std::move(sh));
                break;
            }
            case 'R': s.take(&pallete[1][0], 256 * 4); break;    // RGBA :130-133: file colour i lands on palette index i + 1
            case 'L': s.i32(); s.dict(); s.i32(); break;
            case 'I': { uint8_t remap[256]; s.take(remap, 256); break; }
            case 'r': s.dict(); break;
            case 'n': {
                s.i32();                                         // node id
                auto nd = s.dict();
                node_name = nd.count("_name") ? nd["_name"] : "";
                Node node;
                if (h[1] == 'T') {                               // nTRN :209-255
                    node.kind = 'T'; node.name = node_name;
                    node.child = s.i32(); s.i32(); s.i32(); s.i32();
                    auto m = s.dict();
                    const std::string t = m["_t"];
                    if (!t.empty()) {
                        long b = 0, e = (long)(int)(t.find(' ', (size_t)b + 1) + 1);
                        from_chars_range(t, b, e, node.t[0]);
                        b = e; e = (long)(int)(t.find(' ', (size_t)b + 1) + 1);
                        from_chars_range(t, b, e, node.t[1]);
                        b = e; e = (long)t.size();
                        from_chars_range(t, b, e, node.t[2]);
                    }
                    int r = 4;
                    const std::string rs = m["_r"];
                    if (!rs.empty()) std::from_chars(rs.data(), rs.data() + rs.size(), r);
                    node.m = matrix_of((uint8_t)r);
                    nodes.push_back(std::move(node));
                } else if (h[1] == 'G') {                        // nGRP :257-270
                    node.kind = 'G';
                    const int32_t k = s.i32();
                    if (k < 0 || (size_t)k > n / 4) { err = "nGRP child count exceeds the file"; return false; }
                    for (int32_t i = 0; i < k; ++i) node.children.push_back(s.i32());
                    nodes.push_back(std::move(node));
                } else if (h[1] == 'S') {                        // nSHP :272-282
                    node.kind = 'S';
                    s.i32(); node.shape = s.i32();
                    nodes.push_back(std::move(node));
                    s.dict();
                }
                break;
            }
            default: running = false;
            }
            if (s.bad) { err = "malformed string / dictionary length"; return false; }
            if (!running) break;
        }
        return true;
    }

    // CreateEntity :397-476
    // A valid scene graph is a tree: a node that is referenced twice (a DAG re-instantiates every shared sub-tree, so a
    // 2 KB file of 24 doubled levels would ask for 2^25 entities) or that closes a cycle is rejected.
    std::vector<char> visited;
    size_t model_bytes = 0;                                      // all models of this import together
    static constexpr size_t MAX_MODEL_BYTES = (size_t)1 << 30;
    bool visit(int id) {
        if (visited.size() != nodes.size()) visited.assign(nodes.size(), 0);
        if (visited[(size_t)id]) { err = "scene node referenced twice (the node graph must be a tree)"; return false; }
        visited[(size_t)id] = 1;
        return true;
    }
    int create(vxl_vox_scene& out, const Node& root, int parent, int& counter, int depth) {
        if (depth > 64) { err = "node tree deeper than 64 levels (cycle?)"; return -2; }
        if (root.child < 0 || root.child >= (int)nodes.size()) { err = "nTRN child id does not name a node"; return -2; }
        if (!visit(root.child)) return -2;
        const Node& node = nodes[(size_t)root.child];
        vxl_vox_import_entity e;
        memset(&e, 0, sizeof e);
        e.parent = parent; e.model = -1;
        float p[3] = {(float)root.t[0] * 0.1f, (float)root.t[2] * 0.1f, (float)(-root.t[1]) * 0.1f};          // Flip-Z-Axis
        const int idx = (int)out.entities.size();
        if (node.kind == 'G') {
            memcpy(e.position, p, sizeof p);
            out.entities.push_back(e);
            for (int c : node.children) {
                if (c < 0 || c >= (int)nodes.size() || nodes[(size_t)c].kind != 'T') { err = "nGRP child is not a transform node"; return -2; }
                if (!visit(c)) return -2;
                if (create(out, nodes[(size_t)c], idx, counter, depth + 1) == -2) return -2;
            }
            return idx;
        }
        if (node.kind == 'S') {
            if (node.shape < 0 || node.shape >= (int)shapes.size()) { err = "nSHP names a model the file does not hold"; return -2; }
            const Shape& sh = shapes[(size_t)node.shape];
            const Matrix& M = root.m;
            if (M.rx > 2 || M.ry > 2 || M.rz < 0 || M.rz > 2) { err = "nTRN _r is not a permutation"; return -2; }   // the reference CHECK(0)s
            const int ts[3] = {sh.size[M.rx], sh.size[M.ry], sh.size[M.rz]};
            if (ts[0] <= 0 || ts[1] <= 0 || ts[2] <= 0 || ts[0] > 256 || ts[1] > 256 || ts[2] > 256) {
                err = "bad SIZE (XYZI coordinates are bytes: a model is at most 256^3)"; return -2; }
            const int center[3] = {M.sx ? ts[0] - ts[0] / 2 : ts[0] / 2, M.sz ? ts[2] - ts[2] / 2 : ts[2] / 2, !M.sy ? ts[1] - ts[1] / 2 : ts[1] / 2};
            for (int i = 0; i < 3; ++i) e.position[i] = p[i] - (float)center[i] * 0.1f;
            vxl_vox_scene::Model mdl;
            auto pad = [](int v) { return ((v - 1) & ~3) + 4; };                       // VoxAsset(int32, int32, int32)
            mdl.dims[0] = pad(ts[0]); mdl.dims[1] = pad(ts[2]); mdl.dims[2] = pad(ts[1]);
            model_bytes += (size_t)mdl.dims[0] * mdl.dims[1] * mdl.dims[2];
            if (model_bytes > MAX_MODEL_BYTES) { err = "the models of this file exceed 1 GiB together"; return -2; }
            mdl.data.assign((size_t)mdl.dims[0] * mdl.dims[1] * mdl.dims[2], 0);
            for (uint32_t d : sh.voxels) {
                const int v[3] = {(int)(d & 0xFF), (int)((d >> 8) & 0xFF), (int)((d >> 16) & 0xFF)};
                const int x = v[M.rx], y = v[M.ry], z = v[M.rz];
                const int cx = M.sx ? ts[0] - x - 1 : x, cy = M.sy ? ts[1] - y - 1 : y, cz = M.sz ? ts[2] - z - 1 : z;
                const int X = cx, Y = cz, Z = ts[1] - 1 - cy;                            // Flip-Z-Axis
                if (X < 0 || Y < 0 || Z < 0 || X >= mdl.dims[0] || Y >= mdl.dims[1] || Z >= mdl.dims[2]) { err = "XYZI voxel outside SIZE"; return -2; }
                mdl.data[(size_t)X + (size_t)Y * mdl.dims[0] + (size_t)Z * mdl.dims[0] * mdl.dims[1]] = (uint8_t)(d >> 24);
            }
            mdl.name = root.name.empty() ? std::to_string(counter++) : root.name.substr(0, sizeof e.name - 1);   // one name for the file, the entity and the .pf
            strncpy(e.name, mdl.name.c_str(), sizeof e.name - 1);
            e.model = (int)out.models.size();
            out.models.push_back(std::move(mdl));
            out.entities.push_back(e);
            return idx;
        }
        return -1;                                               // entt::null: the entity exists in the world but is never saved
    }
};

// fmt's "{}" of a float: the shortest digits that round-trip, fixed notation with ".0" for integers, exponent form outside [1e-4, 1e16)
std::string fmt_float(float v) {
    if (std::isnan(v)) return "nan";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[64];
    const float a = std::fabs(v);
    if (a != 0.0f && (a < 1e-4f || a >= 1e16f)) {
        auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
        return std::string(buf, r.ptr);                          // d.ddde-XX, two exponent digits at least: the same as fmt
    }
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    std::string s(buf, r.ptr);
    if (s.find('.') == std::string::npos) s += ".0";
    return s;
}
std::string write_vec3(const float* v) { return fmt_float(v[0]) + " " + fmt_float(v[1]) + " " + fmt_float(v[2]); }
std::string json_string(const std::string& s) {                  // json11 dump(const std::string&)
    std::string o = "\"";
    for (unsigned char ch : s) {
        switch (ch) {
            case '\\': o += "\\\\"; break; case '"': o += "\\\""; break; case '\b': o += "\\b"; break; case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break; case '\r': o += "\\r"; break; case '\t': o += "\\t"; break;
            default:
                if (ch <= 0x1f) { char b[8]; snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
                else o += (char)ch;
        }
    }
    return o + "\"";
}
std::string hex_upper(uint64_t g) { char b[32]; snprintf(b, sizeof b, "%llX", (unsigned long long)g); return b; }

bool write_file(const std::filesystem::path& p, const void* a, size_t na, const void* b, size_t nb, std::string& err) {
    std::error_code ec;
    std::filesystem::create_directories(p.parent_path(), ec);
    FILE* f = fopen(p.string().c_str(), "wb");
    if (!f) { err = "cannot create " + p.string(); return false; }
    const bool ok = (na == 0 || fwrite(a, 1, na, f) == na) && (nb == 0 || fwrite(b, 1, nb, f) == nb);
    fclose(f);
    if (!ok) err = "short write on " + p.string();
    return ok;
}

}  // namespace

extern "C" {

int vxl_vox_import_memory(const void* data, uint64_t size, vxl_vox_scene** out) {
    if (!data || !out) { set_error("vxl_vox_import_memory: bad argument"); return VXL_ERR_INVALID; }
    try {
    Importer im;
    if (!im.parse((const uint8_t*)data, (size_t)size)) { set_error("vxl_vox_import: " + im.err); return VXL_ERR_INVALID; }
    if (im.nodes.empty() || im.nodes[0].kind != 'T') { set_error("vxl_vox_import: the file has no root transform node"); return VXL_ERR_INVALID; }
    auto sc = std::make_unique<vxl_vox_scene>();
    for (int i = 0; i < 256; ++i) {                              // VoxImporter::Import :489-497; `a` is uninitialised in the reference, 0 here
        uint8_t* m = sc->records + i * 7;
        m[0] = im.pallete[i][0]; m[1] = im.pallete[i][1]; m[2] = im.pallete[i][2]; m[3] = 0;
        m[4] = im.surfaces[i][1]; m[5] = im.surfaces[i][2]; m[6] = im.surfaces[i][0];
    }
    int counter = 0;
    if (im.create(*sc, im.nodes[0], -1, counter, 0) == -2) { set_error("vxl_vox_import: " + im.err); return VXL_ERR_INVALID; }
    *out = sc.release();
    return VXL_OK;
    } catch (const std::bad_alloc&) {                             // nothing crosses the C boundary
        set_error("vxl_vox_import: out of memory");
        return VXL_ERR_OOM;
    } catch (const std::exception& ex) {
        set_error(std::string("vxl_vox_import: ") + ex.what());
        return VXL_ERR_INVALID;
    }
}

int vxl_vox_import(const char* vox_path, vxl_vox_scene** out) {
    if (!vox_path || !out) { set_error("vxl_vox_import: bad argument"); return VXL_ERR_INVALID; }
    FILE* f = fopen(vox_path, "rb");
    if (!f) { set_error(std::string("vxl_vox_import: cannot open ") + vox_path); return VXL_ERR_INVALID; }
    std::vector<uint8_t> buf;
    try {
        uint8_t chunk[65536];
        size_t k;
        while ((k = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + k);
    } catch (const std::exception&) {
        fclose(f);
        set_error("vxl_vox_import: out of memory reading the file");
        return VXL_ERR_OOM;
    }
    fclose(f);
    return vxl_vox_import_memory(buf.data(), buf.size(), out);
}

int vxl_vox_scene_counts(const vxl_vox_scene* sc, int* n_entities, int* n_models) {
    if (!sc) { set_error("vxl_vox_scene_counts: bad argument"); return VXL_ERR_INVALID; }
    if (n_entities) *n_entities = (int)sc->entities.size();
    if (n_models) *n_models = (int)sc->models.size();
    return VXL_OK;
}

int vxl_vox_scene_entities(const vxl_vox_scene* sc, vxl_vox_import_entity* out, int cap) {
    if (!sc || !out || cap < (int)sc->entities.size()) { set_error("vxl_vox_scene_entities: bad argument / array too small"); return VXL_ERR_INVALID; }
    for (size_t i = 0; i < sc->entities.size(); ++i) out[i] = sc->entities[i];
    return VXL_OK;
}

int vxl_vox_scene_model(const vxl_vox_scene* sc, int model, int32_t dims[3], char name[64], uint8_t* out, uint64_t cap) {
    if (!sc || !dims || model < 0 || model >= (int)sc->models.size()) { set_error("vxl_vox_scene_model: bad argument"); return VXL_ERR_INVALID; }
    const auto& m = sc->models[(size_t)model];
    memcpy(dims, m.dims, sizeof m.dims);
    if (name) { memset(name, 0, 64); strncpy(name, m.name.c_str(), 63); }
    if (!out) return VXL_OK;                                     // size query
    if (cap < m.data.size()) { set_error("vxl_vox_scene_model: output buffer too small"); return VXL_ERR_LIMIT; }
    memcpy(out, m.data.data(), m.data.size());
    return VXL_OK;
}

int vxl_vox_scene_pallete(const vxl_vox_scene* sc, uint8_t records[1792]) {
    if (!sc || !records) { set_error("vxl_vox_scene_pallete: bad argument"); return VXL_ERR_INVALID; }
    memcpy(records, sc->records, sizeof sc->records);
    return VXL_OK;
}

int vxl_vox_scene_write(const vxl_vox_scene* sc, const char* mods_dir, const char* path, const char* file_name) {
    if (!sc || !mods_dir || !path || !file_name || !*file_name) { set_error("vxl_vox_scene_write: bad argument"); return VXL_ERR_INVALID; }
    namespace fs = std::filesystem;
    try {
    const fs::path mods(mods_dir), P(path), F(file_name);
    std::string err;
    // shape names come verbatim from the (untrusted) .vox file and become file names under Mods/<path>/<file_name>/: a name that is
    // absolute, climbs with "..", or carries a separator / NUL would write outside that directory
    auto unsafe = [](const std::string& n) {
        return n.empty() || n == "." || n == ".." || n.find('/') != std::string::npos || n.find('\\') != std::string::npos ||
               n.find('\0') != std::string::npos || n.find(':') != std::string::npos;
    };
    for (const auto& m : sc->models)
        if (unsafe(m.name)) { set_error("vxl_vox_scene_write: shape name '" + m.name + "' is not a plain file name"); return VXL_ERR_INVALID; }
    if (unsafe(F.generic_string()) || P.is_absolute()) { set_error("vxl_vox_scene_write: file_name must be a plain name and path relative"); return VXL_ERR_INVALID; }
    for (const auto& part : P) if (part == "..") { set_error("vxl_vox_scene_write: path must not contain '..'"); return VXL_ERR_INVALID; }
    // Assets::CreateAsset (Assets.cpp:24-42): GUID = Hash(path relative to Mods/), file = Mods/<path>
    const std::string p_rel = (P / F / F).replace_extension("p").generic_string();
    uint64_t p_guid; vxl_asset_guid(p_rel.c_str(), &p_guid);
    if (!write_file(mods / p_rel, sc->records, sizeof sc->records, nullptr, 0, err)) { set_error("vxl_vox_scene_write: " + err); return VXL_ERR_INVALID; }
    std::vector<uint64_t> v_guid(sc->models.size());
    for (size_t i = 0; i < sc->models.size(); ++i) {
        const auto& m = sc->models[i];
        const std::string rel = (P / F / m.name).replace_extension("v").generic_string();
        vxl_asset_guid(rel.c_str(), &v_guid[i]);
        if (!write_file(mods / rel, m.dims, sizeof m.dims, m.data.data(), m.data.size(), err)) { set_error("vxl_vox_scene_write: " + err); return VXL_ERR_INVALID; }
    }
    // PrefabAsset::FromWorld: depth first from the root, which is how `entities` is ordered; ids = creation order in a fresh registry
    std::string js = "[";
    for (size_t i = 0; i < sc->entities.size(); ++i) {
        const auto& e = sc->entities[i];
        const float zero[3] = {0, 0, 0}, one[3] = {1, 1, 1};
        if (i) js += ", ";
        js += "{\"Id\": " + std::to_string(i) + ", \"Name\": " + json_string(i == 0 ? F.generic_string() : std::string(e.name));   // W->SetName(root, _FileName)
        if (i) js += ", \"Parent\": " + std::to_string(e.parent);
        js += ", \"Transform\": {\"Position\": " + json_string(write_vec3(e.position)) + ", \"Rotation\": " + json_string(write_vec3(zero)) +
              ", \"Scale\": " + json_string(write_vec3(one)) + "}";
        if (e.model >= 0)
            js += ", \"VoxRenderer\": {\"Pallete\": " + json_string(hex_upper(p_guid)) + ", \"Pivot\": " + json_string(write_vec3(zero)) +
                  ", \"Vox\": " + json_string(hex_upper(v_guid[(size_t)e.model])) + "}";
        js += "}";
    }
    js += "]";
    fs::path pf = F;
    const std::string pf_rel = (P / pf.replace_extension("pf")).generic_string();
    if (!write_file(mods / pf_rel, js.data(), js.size(), nullptr, 0, err)) { set_error("vxl_vox_scene_write: " + err); return VXL_ERR_INVALID; }
    return VXL_OK;
    } catch (const std::bad_alloc&) {
        set_error("vxl_vox_scene_write: out of memory");
        return VXL_ERR_OOM;
    } catch (const std::exception& ex) {                         // std::filesystem errors
        set_error(std::string("vxl_vox_scene_write: ") + ex.what());
        return VXL_ERR_INVALID;
    }
}

int vxl_vox_scene_free(vxl_vox_scene* sc) {
    delete sc;
    return VXL_OK;
}

}  // extern "C"
