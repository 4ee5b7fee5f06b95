// ORACLE build (test infrastructure): C entry points over the reference's OWN scene-file layer, compiled where it lies
// (/root/reference/src/factory/btrc/factory/node/{parser,node}.cpp) into oracle/_ref/libbtrcref.so.
// Used by tests/test_reference_host_pins_cpu.py to pin the product's and the oracle's JSON front ends: comments, $include,
// $reference resolution, the std::to_string(float) round trip, duplicate keys, typed values.  Nothing here is shipped.
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <string>

#include <btrc/core/spectrum.h>
#include <btrc/factory/node/parser.h>
#include <btrc/utils/file.h>

using namespace btrc;
using namespace btrc::factory;

static thread_local std::string g_text, g_err;

// utils/file.cpp does not compile with gcc (a bare `#elif` in its platform section, file.cpp:6); the one function the parser
// needs from it is restated here: read_txt_file (file.cpp:14-22) = the whole file through a text-mode ifstream.
BTRC_BEGIN
std::string read_txt_file(const std::string &filename) {
    std::ifstream fin(filename, std::ios_base::in);
    if (!fin) throw BtrcException("failed to open text file: " + filename);
    std::stringstream sst;
    sst << fin.rdbuf();
    return sst.str();
}
BTRC_END

// Canonical dump of a resolved tree: groups in first-insertion key order {"k":v,...}, arrays [v,...], values "text".
static void dump(const RC<const Node> &n, std::string &out, int depth) {
    if (depth > 200) { out += "<deep>"; return; }
    switch (n->get_type()) {
    case Node::Type::Group: {
        auto g = n->as_group();
        out += "{";
        bool first = true;
        for (auto &key : g->get_ordered_keys()) {
            if (!first) out += ",";
            first = false;
            out += "\"" + std::string(key) + "\":";
            dump(g->find_child_node(key), out, depth + 1);
        }
        out += "}";
        break;
    }
    case Node::Type::Array: {
        auto a = n->as_array();
        out += "[";
        for (size_t i = 0; i < a->get_size(); ++i) {
            if (i) out += ",";
            dump(a->get_element(i), out, depth + 1);
        }
        out += "]";
        break;
    }
    case Node::Type::Value:
        out += "\"" + n->as_value()->get_string() + "\"";
        break;
    }
}

static RC<Group> parse_file(const char *path) {  // cli/main.cpp:35-40
    JSONParser parser;
    parser.set_source(read_txt_file(path));
    parser.add_include_directory(std::filesystem::path(path).parent_path());
    parser.parse();
    return parser.get_result();
}

static RC<const Node> walk(RC<const Node> n, const char *node_path) {  // "a/b/0/c" from the root
    std::string cur;
    auto step = [&](const std::string &s) {
        if (s.empty()) return;
        if (n->get_type() == Node::Type::Array) n = n->as_array()->get_element(std::stoul(s));
        else n = n->as_group()->child_node(s);
    };
    for (const char *p = node_path; *p; ++p) {
        if (*p == '/') { step(cur); cur.clear(); }
        else cur += *p;
    }
    step(cur);
    return n;
}

extern "C" {

const char *ref_json_last_error() { return g_err.c_str(); }

// Dump of the parsed + resolved scene file, or NULL (ref_json_last_error()).
const char *ref_json_dump(const char *path) {
    try {
        g_text.clear();
        dump(parse_file(path), g_text, 0);
        return g_text.c_str();
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

// Node::parse<T> (node.cpp:276-325) of the node at node_path.  kind: 0 int, 1 float, 2 bool, 3 Vec3f, 4 Spectrum, 5 Degree, 6 Radian,
// 7 Transform3D (16 floats, Mat4::at(r, c) row-major), 8 Transform2D (9 floats, Mat3::at(r, c) row-major).  Returns the float count, -1 on error.
int ref_json_parse(const char *path, const char *node_path, int kind, float *out) {
    try {
        auto n = walk(parse_file(path), node_path);
        switch (kind) {
        case 0: out[0] = (float)n->parse<int>(); return 1;
        case 1: out[0] = n->parse<float>(); return 1;
        case 2: out[0] = n->parse<bool>() ? 1.0f : 0.0f; return 1;
        case 3: { Vec3f v = n->parse<Vec3f>(); out[0] = v.x; out[1] = v.y; out[2] = v.z; return 3; }
        case 4: { Spectrum s = n->parse<Spectrum>(); out[0] = s.r; out[1] = s.g; out[2] = s.b; return 3; }
        case 5: out[0] = n->parse<Degree>().value; return 1;
        case 6: out[0] = n->parse<Radian>().value; return 1;
        case 7: { Transform3D t = n->parse<Transform3D>(); for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = t.mat.at(r, c); return 16; }
        case 8: { Transform2D t = n->parse<Transform2D>(); for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[3 * r + c] = t.mat.at(r, c); return 9; }
        }
        g_err = "unknown kind";
        return -1;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

}  // extern "C"
