// JSON scene parser and typed node access (see node.h for the reference lines each piece mirrors).
#include "node.h"

#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <sstream>

namespace btrc {

// ----------------------------------------------------------------------------------------------
// host_math.h: transform factories (mat4.cpp:108-184, transform.cpp:92-125)
// ----------------------------------------------------------------------------------------------
static Mat4 mat_rotate_axis(Vec3 axis, float rad) {  // mat4.cpp:117-145
    // Drop-in quirk: the reference fills its column-major storage (Mat4::at(r, c) = data[c][r], mat4.cpp:98-106) with the
    // textbook row-major entries, so "rotate" about an axis is the transpose of the usual matrix, i.e. it turns by -rad.
    // Pinned against the reference's own code (tests/test_reference_host_pins_cpu.py); reproduced, not fixed.
    const Vec3 a = normalize(axis);
    const float sinv = std::sin(rad), cosv = std::cos(rad);
    Mat4 r = Mat4::identity();
    r.m[0][0] = a.x * a.x + (1 - a.x * a.x) * cosv;
    r.m[1][0] = a.x * a.y * (1 - cosv) - a.z * sinv;
    r.m[2][0] = a.x * a.z * (1 - cosv) + a.y * sinv;
    r.m[0][1] = a.x * a.y * (1 - cosv) + a.z * sinv;
    r.m[1][1] = a.y * a.y + (1 - a.y * a.y) * cosv;
    r.m[2][1] = a.y * a.z * (1 - cosv) - a.x * sinv;
    r.m[0][2] = a.x * a.z * (1 - cosv) - a.y * sinv;
    r.m[1][2] = a.y * a.z * (1 - cosv) + a.x * sinv;
    r.m[2][2] = a.z * a.z + (1 - a.z * a.z) * cosv;
    return r;
}
static Mat4 mat_rotate_x(float rad) {
    const float S = std::sin(rad), C = std::cos(rad);
    return Mat4::rows(1, 0, 0, 0, 0, C, -S, 0, 0, S, C, 0, 0, 0, 0, 1);
}
static Mat4 mat_rotate_y(float rad) {
    const float S = std::sin(rad), C = std::cos(rad);
    return Mat4::rows(C, 0, S, 0, 0, 1, 0, 0, -S, 0, C, 0, 0, 0, 0, 1);
}
static Mat4 mat_rotate_z(float rad) {
    const float S = std::sin(rad), C = std::cos(rad);
    return Mat4::rows(C, -S, 0, 0, S, C, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1);
}

Transform3D Transform3D::translate(float x, float y, float z) {
    Transform3D t;
    t.mat = Mat4::rows(1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z, 0, 0, 0, 1);
    t.inv = Mat4::rows(1, 0, 0, -x, 0, 1, 0, -y, 0, 0, 1, -z, 0, 0, 0, 1);
    return t;
}
Transform3D Transform3D::rotate(Vec3 axis, float rad) {
    Transform3D t;
    t.mat = mat_rotate_axis(axis, rad);
    t.inv = mat_rotate_axis(axis, -rad);
    return t;
}
Transform3D Transform3D::rotate_x(float rad) { Transform3D t; t.mat = mat_rotate_x(rad); t.inv = mat_rotate_x(-rad); return t; }
Transform3D Transform3D::rotate_y(float rad) { Transform3D t; t.mat = mat_rotate_y(rad); t.inv = mat_rotate_y(-rad); return t; }
Transform3D Transform3D::rotate_z(float rad) {
    // Drop-in quirk: the reference builds BOTH matrices from -rad (transform.cpp:112-115), so
    // "rotate_z" turns the other way and carries a wrong inverse.  Scene files written for Btrc
    // depend on it, so it is reproduced, not fixed.
    Transform3D t; t.mat = mat_rotate_z(-rad); t.inv = mat_rotate_z(-rad); return t;
}
Transform3D Transform3D::scale(float x, float y, float z) {
    Transform3D t;
    t.mat = Mat4::rows(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1);
    t.inv = Mat4::rows(1 / x, 0, 0, 0, 0, 1 / y, 0, 0, 0, 0, 1 / z, 0, 0, 0, 0, 1);
    return t;
}

// ----------------------------------------------------------------------------------------------
// JSON text -> node tree
// ----------------------------------------------------------------------------------------------
namespace {

struct JsonReader {
    const std::string &s;
    size_t i = 0;
    explicit JsonReader(const std::string &src) : s(src) {}

    [[noreturn]] void fail(const std::string &what) const {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); ++k) if (s[k] == '\n') ++line;
        throw Error("json parse error at line " + std::to_string(line) + ": " + what);
    }
    void skip_ws() {
        for (;;) {
            while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
            if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {  // comments are allowed (parser.cpp:63)
                while (i < s.size() && s[i] != '\n') ++i;
            } else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
                size_t e = s.find("*/", i + 2);
                if (e == std::string::npos) fail("unterminated comment");
                i = e + 2;
            } else
                return;
        }
    }
    std::string parse_string() {
        if (s[i] != '"') fail("string expected");
        ++i;
        std::string out;
        while (i < s.size() && s[i] != '"') {
            char c = s[i++];
            if (c == '\\') {
                if (i >= s.size()) fail("bad escape");
                char e = s[i++];
                switch (e) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    if (i + 4 > s.size()) fail("bad \\u escape");
                    unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
                    i += 4;
                    if (cp < 0x80) out += char(cp);
                    else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
                    else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += e;
                }
            } else
                out += c;
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    int depth = 0;
    struct DepthGuard {
        int &d;
        explicit DepthGuard(int &dd) : d(dd) { ++d; }
        ~DepthGuard() { --d; }
    };
    NodeP parse_value() {
        DepthGuard guard(depth);
        if (depth > 512) fail("nesting too deep");  // recursion bound: a damaged file must not overflow the stack
        skip_ws();
        if (i >= s.size()) fail("unexpected end of input");
        auto n = std::make_shared<Node>();
        const char c = s[i];
        if (c == '{') {
            n->type = Node::Type::Group;
            ++i;
            skip_ws();
            if (i < s.size() && s[i] == '}') { ++i; return n; }
            for (;;) {
                skip_ws();
                std::string key = parse_string();
                skip_ws();
                if (i >= s.size() || s[i] != ':') fail("':' expected");
                ++i;
                NodeP v = parse_value();
                bool replaced = false;  // ordered_json keeps the first position, last value
                for (auto &kv : n->children) if (kv.first == key) { kv.second = v; replaced = true; }
                if (!replaced) n->children.emplace_back(std::move(key), std::move(v));
                skip_ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; break; }
                fail("',' or '}' expected");
            }
            return n;
        }
        if (c == '[') {
            n->type = Node::Type::Array;
            ++i;
            skip_ws();
            if (i < s.size() && s[i] == ']') { ++i; return n; }
            for (;;) {
                n->elements.push_back(parse_value());
                skip_ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; break; }
                fail("',' or ']' expected");
            }
            return n;
        }
        n->type = Node::Type::Value;
        if (c == '"') { n->value = parse_string(); return n; }
        if (s.compare(i, 4, "true") == 0) { i += 4; n->value = "true"; return n; }
        if (s.compare(i, 5, "false") == 0) { i += 5; n->value = "false"; return n; }
        if (s.compare(i, 4, "null") == 0) fail("null is not a valid scene value");
        // number
        size_t b = i;
        bool is_float = false;
        if (i < s.size() && (s[i] == '-' || s[i] == '+')) ++i;
        while (i < s.size() && (std::isdigit((unsigned char)s[i]) || s[i] == '.' || s[i] == 'e' || s[i] == 'E' || s[i] == '-' || s[i] == '+')) {
            if (s[i] == '.' || s[i] == 'e' || s[i] == 'E') is_float = true;
            ++i;
        }
        if (b == i) fail("unexpected character");
        const std::string tok = s.substr(b, i - b);
        if (is_float) {
            // parser.cpp:38-39: std::to_string(json.get<float>()) -- "%f" of the value rounded to float
            const float f = static_cast<float>(std::strtod(tok.c_str(), nullptr));
            n->value = std::to_string(f);
        } else {
            n->value = std::to_string(static_cast<int>(std::strtoll(tok.c_str(), nullptr, 10)));
        }
        return n;
    }
};

std::vector<std::string> split_path(const std::string &p) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : p) {
        if (c == '/') { if (!cur.empty()) out.push_back(cur); cur.clear(); }
        else cur += c;
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

bool parse_index(const std::string &s, size_t &out) {
    if (s.empty()) return false;
    for (char c : s) if (!std::isdigit((unsigned char)c)) return false;
    out = std::strtoul(s.c_str(), nullptr, 10);
    return true;
}

NodeP find_node(std::vector<NodeP> &cur_path, const std::string &path, int hops = 0);

void resolve_references(std::vector<NodeP> &cur_path, NodeP &node) {  // parser.cpp:135-165
    if (node->is_group()) {
        cur_path.push_back(node);
        for (auto &kv : node->children) resolve_references(cur_path, kv.second);
        cur_path.pop_back();
    } else if (node->is_array()) {
        cur_path.push_back(node);
        for (auto &e : node->elements) resolve_references(cur_path, e);
        cur_path.pop_back();
    } else {
        const std::string &str = node->value;
        if (str.rfind("$reference{", 0) == 0) {
            if (str.back() != '}') throw Error("'}' for '$reference{' is not found");
            NodeP target = find_node(cur_path, str.substr(11, str.size() - 12));
            // a reference to one of its own ancestors would make the tree cyclic (the reference recurses until it crashes)
            for (const NodeP &a : cur_path) if (a == target) throw Error("cyclic $reference: " + str);
            node = target;
        }
    }
}

NodeP find_node(std::vector<NodeP> &cur_path, const std::string &path, int hops) {  // parser.cpp:167-252
    if (hops > 64) throw Error("cyclic $reference: " + path);
    const auto secs = split_path(path);
    if (secs.empty()) throw Error("invalid node path: " + path);
    const std::string &start = secs.front();
    int si;
    if (start == "$root") si = 0;
    else {
        si = (int)cur_path.size() - 1;
        while (si >= 0) {
            const NodeP &n = cur_path[si];
            if (n->is_group()) { if (n->find(start)) break; }
            else if (n->is_array()) { size_t idx; if (parse_index(start, idx) && idx < n->elements.size()) break; }
            --si;
        }
    }
    if (si < 0) throw Error("invalid node path: " + path);
    NodeP node = cur_path[si];
    for (size_t i = 0; i < secs.size(); ++i) {
        const std::string &sec = secs[i];
        if (i == 0 && sec == "$root") continue;
        if (node->is_group()) {
            node = node->find(sec);
            if (!node) throw Error("invalid node path: " + path);
        } else if (node->is_array()) {
            size_t idx;
            if (!parse_index(sec, idx) || idx >= node->elements.size()) throw Error("invalid node path: " + path);
            node = node->elements[idx];
        } else
            throw Error("invalid node path: " + path);
    }
    if (node->is_value() && node->value.find("$reference{") != std::string::npos) {
        const std::string &str = node->value;
        if (str.back() != '}') throw Error("'}' for '$reference{' is not found");
        node = find_node(cur_path, str.substr(11, str.size() - 12), hops + 1);
    }
    return node;
}

}  // namespace

std::string read_text_file(const std::string &filename) {
    std::ifstream fin(filename, std::ios::in | std::ios::binary);
    if (!fin) throw Error("failed to open text file: " + filename);
    std::stringstream ss;
    ss << fin.rdbuf();
    return ss.str();
}

NodeP parse_scene_text(std::string src, const std::string &include_dir) {
    namespace fs = std::filesystem;
    for (int expansions = 0;; ++expansions) {  // parser.cpp:106-133
        size_t beg = src.find("\"$include{");
        if (beg == std::string::npos) break;
        if (expansions > 4096 || src.size() > (1u << 28)) throw Error("too many '$include's (a file that includes itself?)");
        beg += 10;
        const size_t end = src.find("}\"", beg);
        if (end == std::string::npos) throw Error("failed to find the enclosing '}' of a '$include{'");
        const std::string inc = src.substr(beg, end - beg);
        fs::path p(inc);
        if (p.is_relative()) {
            p = fs::path(include_dir) / p;
            if (!fs::exists(p)) throw Error("included file '" + inc + "' is not found");
        }
        const std::string included = read_text_file(p.lexically_normal().string());
        src = src.substr(0, beg - 10) + included + src.substr(end + 2);
    }
    JsonReader rd(src);
    NodeP root = rd.parse_value();
    if (!root->is_group()) throw Error("root node is not a group");
    std::vector<NodeP> path = {root};
    for (auto &kv : root->children) resolve_references(path, kv.second);
    return root;
}

NodeP parse_scene_file(const std::string &filename) {
    namespace fs = std::filesystem;
    const std::string dir = fs::absolute(fs::path(filename)).parent_path().string();
    return parse_scene_text(read_text_file(filename), dir);
}

// ----------------------------------------------------------------------------------------------
// typed access
// ----------------------------------------------------------------------------------------------
NodeP Node::find(const std::string &name) const {
    if (!is_group()) throw Error("group node expected");
    for (auto &kv : children) if (kv.first == name) return kv.second;
    return nullptr;
}
NodeP Node::child(const std::string &name) const {
    NodeP r = find(name);
    if (!r) throw Error("child " + name + " not found");
    return r;
}
int Node::as_int() const {
    if (!is_value()) throw Error("value node expected");
    try { return std::stoi(value); } catch (...) { throw Error("invalid integer: " + value); }
}
float Node::as_float() const {
    if (!is_value()) throw Error("value node expected");
    try { return std::stof(value); } catch (...) { throw Error("invalid number: " + value); }
}
bool Node::as_bool() const {
    if (!is_value()) throw Error("value node expected");
    return value == "true";
}
const std::string &Node::as_string() const {
    if (!is_value()) throw Error("value node expected");
    return value;
}
Vec3 Node::as_vec3() const {  // node.cpp:50-74
    if (is_array()) {
        if (elements.size() == 1) { const float v = elements[0]->as_float(); return Vec3(v, v, v); }
        if (elements.size() == 3) return Vec3(elements[0]->as_float(), elements[1]->as_float(), elements[2]->as_float());
        throw Error("unexpected array size: " + std::to_string(elements.size()));
    }
    if (is_value()) { const float v = as_float(); return Vec3(v, v, v); }
    throw Error("unexpected group node");
}
float Node::as_degree() const {  // node.cpp:76-86
    if (!is_group()) throw Error("group expected");
    if (auto c = find("deg")) return c->as_float();
    if (auto c = find("rad")) return 180.0f / 3.1415926535f * c->as_float();
    throw Error("invalid angle node");
}
float Node::as_radian() const {  // node.cpp:88-98
    if (!is_group()) throw Error("group expected");
    if (auto c = find("deg")) return 3.1415926535f / 180.0f * c->as_float();
    if (auto c = find("rad")) return c->as_float();
    throw Error("invalid angle node");
}
Transform3D Node::as_transform3d() const {  // node.cpp:166-213
    if (is_array()) {
        Transform3D ret;
        for (auto &e : elements) ret = e->as_transform3d() * ret;
        return ret;
    }
    if (auto t = find("translate")) { const Vec3 v = t->as_vec3(); return Transform3D::translate(v.x, v.y, v.z); }
    if (auto r = find("rotate")) return Transform3D::rotate(r->child("axis")->as_vec3(), r->child("angle")->as_radian());
    if (auto r = find("rotate_x")) return Transform3D::rotate_x(r->as_radian());
    if (auto r = find("rotate_y")) return Transform3D::rotate_y(r->as_radian());
    if (auto r = find("rotate_z")) return Transform3D::rotate_z(r->as_radian());
    if (auto s = find("scale")) { const Vec3 v = s->as_vec3(); return Transform3D::scale(v.x, v.y, v.z); }
    throw Error("unknown transform3d type");
}
// node.cpp:126-164.  translate / scale go through parse<Vec2f>, which the reference never instantiates
// (node.cpp:321-324 throws "unimplemented"); kept.  Products are true 3x3 products (the reference's Mat3 operator*
// reads a fourth column that does not exist: utils/math/mat3.cpp:108-123).
Mat3 Node::as_transform2d() const {
    if (is_array()) {
        Mat3 ret = Mat3::identity();
        for (auto &e : elements) ret = e->as_transform2d() * ret;
        return ret;
    }
    if (find("translate") || find("scale")) throw Error("unimplemented");
    auto rows = [](float a, float b, float c, float d, float e, float f, float g, float h, float i) {
        Mat3 m = {{{a, b, c}, {d, e, f}, {g, h, i}}};
        return m;
    };
    if (auto r = find("rotate")) {
        const float rad = r->child("angle")->as_radian();
        const float S = std::sin(rad), C = std::cos(rad);
        return rows(C, -S, 0, S, C, 0, 0, 0, 1);
    }
    if (auto n = find("inverse_u"); n && n->as_bool()) return rows(-1, 0, 1, 0, 1, 0, 0, 0, 0);
    if (auto n = find("inverse_v"); n && n->as_bool()) return rows(1, 0, 0, 0, -1, 1, 0, 0, 1);
    if (auto n = find("swap_uv"); n && n->as_bool()) return rows(0, 1, 0, 1, 0, 0, 0, 0, 1);
    throw Error("unknown transform2d type");
}
int Node::int_or(const std::string &name, int d) const { auto c = find(name); return c ? c->as_int() : d; }
float Node::float_or(const std::string &name, float d) const { auto c = find(name); return c ? c->as_float() : d; }
bool Node::bool_or(const std::string &name, bool d) const { auto c = find(name); return c ? c->as_bool() : d; }
std::string Node::string_or(const std::string &name, const std::string &d) const { auto c = find(name); return c ? c->as_string() : d; }

}  // namespace btrc
