// Scene-description node tree and typed parsing: the drop-in boundary for Btrc's JSON scene files.
//
// Mirrors the behaviour (not the code) of the reference's factory layer:
//   factory/btrc/factory/node/parser.cpp:16-45   every JSON leaf is stored as a string; floats go through
//                                               std::to_string(float) ("%f", 6 decimals) and back
//   parser.cpp:106-133                           "$include{file}" textual splice
//   parser.cpp:135-252                           "$reference{a/b/0}" sharing of nodes
//   node/node.cpp:50-214                         Vec3f / Spectrum / angle / Transform3D / Transform2D parsing
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "host_math.h"

namespace btrc {

// reference: BtrcException (common.h:31-36) with BTRC_HI_TRY/WRAP context chaining
class Error : public std::runtime_error {
public:
    explicit Error(const std::string &msg) : std::runtime_error(msg) {}
};

class Node;
using NodeP = std::shared_ptr<Node>;

class Node {
public:
    enum class Type { Group, Array, Value };
    Type type = Type::Value;
    std::string value;                                  // Value
    std::vector<NodeP> elements;                        // Array
    std::vector<std::pair<std::string, NodeP>> children;  // Group (insertion ordered)

    bool is_group() const { return type == Type::Group; }
    bool is_array() const { return type == Type::Array; }
    bool is_value() const { return type == Type::Value; }

    NodeP find(const std::string &name) const;   // group expected; nullptr when absent
    NodeP child(const std::string &name) const;  // throws "child <name> not found"

    // typed values (node.cpp:14-48)
    int as_int() const;
    float as_float() const;
    bool as_bool() const;
    const std::string &as_string() const;
    Vec3 as_vec3() const;       // also Spectrum
    float as_degree() const;
    float as_radian() const;
    Transform3D as_transform3d() const;
    Mat3 as_transform2d() const;

    int int_or(const std::string &name, int d) const;
    float float_or(const std::string &name, float d) const;
    bool bool_or(const std::string &name, bool d) const;
    std::string string_or(const std::string &name, const std::string &d) const;
};

// Parses a scene file: includes resolved against `include_dir`, comments allowed, references resolved.
NodeP parse_scene_text(std::string src, const std::string &include_dir);
NodeP parse_scene_file(const std::string &filename);
std::string read_text_file(const std::string &filename);

}  // namespace btrc
