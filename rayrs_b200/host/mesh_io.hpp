// Mesh file I/O of the host mirror: PLY (the `ply` crate) and Wavefront OBJ (wavefront_obj.rs).
//
// The reference's `ply` crate is a design sketch — every line of ply/src/lib.rs is commented out
//  — so there is no behaviour to match bit for bit; this module implements that sketch
// with its names, its header grammar and its error messages:
//   PlyKeyword::from_line        ply/src/lib.rs:31-136   one header line -> one token
//   PlyFormat / PlyPropertyType  ply/src/lib.rs:137-187
//   PlyHeader, PlyHeaderParser   ply/src/lib.rs:189-317  the Start -> Format -> StartElement ->
//                                                        NewElement -> InElement -> End state machine
//   Ply::load                    ply/src/lib.rs:319-350
// and completes what the sketch leaves open (`add_element` / `add_property` are empty bodies and the
// element data is never read): ascii, binary_little_endian and binary_big_endian bodies, scalar and
// list properties of all eight types.  `load_ply_file` is the PLY twin of
// `wavefront_obj::load_obj_file` (wavefront_obj.rs:15-44): file -> Vec<Triangle>.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "rayrs_host.hpp"

namespace rayrs {

// io::Error of the reference's signatures: kind + message
struct IoError : std::runtime_error {
    enum Kind { InvalidInput, InvalidData, Other, NotFound, UnexpectedEof } kind;
    IoError(Kind k, const std::string& msg) : std::runtime_error(msg), kind(k) {}
};

// geometry.rs:312-355 / 78-101 — what the loaders return; Object::from_triangles / from_spheres wrap them
struct Triangle {
    Vec3 p1, p2, p3;
};
struct Sphere {
    double radius;
    Vec3 origin;
};

namespace ply {

enum class PlyFormat { Ascii, BinaryBigEndian, BinaryLittleEndian };
PlyFormat format_from_string(const std::string& s);  // PlyFormat::from_string, ply/src/lib.rs:144-156

enum class PlyPropertyType { Char, Uchar, Short, Ushort, Int, Uint, Float, Double };
PlyPropertyType property_type_from_string(const std::string& s);  // PlyPropertyType::from_string, :170-186
size_t property_type_size(PlyPropertyType t);

// enum PlyKeyword, ply/src/lib.rs:5-29
struct PlyKeyword {
    enum Tag { Ply, Format, Comment, Element, Property, ListProperty, EndHeader } tag = Ply;
    PlyFormat format = PlyFormat::Ascii;  // Format
    std::string version;                  // Format
    std::string comment;                  // Comment (the whole line, as the sketch stores it)
    std::string name;                     // Element / Property / ListProperty
    size_t length = 0;                    // Element
    PlyPropertyType typ = PlyPropertyType::Float;      // Property
    PlyPropertyType lentype = PlyPropertyType::Uchar;  // ListProperty
    PlyPropertyType elemtype = PlyPropertyType::Int;   // ListProperty
    static PlyKeyword from_line(const std::string& line);  // ply/src/lib.rs:49-73
    bool operator==(const PlyKeyword& o) const;
};

struct PlyProperty {
    std::string name;
    bool is_list = false;
    PlyPropertyType typ = PlyPropertyType::Float;      // scalar type, or the element type of a list
    PlyPropertyType lentype = PlyPropertyType::Uchar;  // list only
};

struct PlyElement {
    std::string name;
    size_t length = 0;
    std::vector<PlyProperty> properties;
    // body, filled by Ply::load: scalars[p][i] for scalar property p; lists[p][i] for list property p
    std::vector<std::vector<double>> scalars;
    std::vector<std::vector<std::vector<double>>> lists;
    int property_index(const std::string& name) const;
};

// struct PlyHeader, ply/src/lib.rs:189-217
struct PlyHeader {
    std::string version;
    PlyFormat format = PlyFormat::Ascii;
    std::vector<PlyElement> elements;
    std::vector<std::string> comments;
    void add_comment(const std::string& c) { comments.push_back(c); }
    void add_element(const std::string& name, size_t length);
    void add_property(const std::string& name, PlyPropertyType typ);
    void add_list_property(const std::string& name, PlyPropertyType lentype, PlyPropertyType elemtype);
};

// enum PlyHeaderParser, ply/src/lib.rs:219-317
struct PlyHeaderParser {
    enum State { Start, Format, StartElement, NewElement, InElement, End } state = Start;
    PlyHeader header;
    void handle_input(const PlyKeyword& inp);  // throws IoError(Other, "expected ...") like the sketch
};

struct Ply {
    PlyHeader header;
    static Ply load(const std::string& path);                    // ply/src/lib.rs:322-349 + the body
    static Ply parse(const std::string& bytes);                  // same, from memory
    const PlyElement* element(const std::string& name) const;
    // vertex x/y/z + face vertex_indices|vertex_index; polygons are fan-triangulated
    std::vector<Triangle> triangles() const;
};

// writer for the synthetic meshes (float32 xyz, uchar-count int32 faces)
void write_ply(const std::string& path, const std::vector<float>& xyz, const std::vector<int32_t>& faces3, PlyFormat format);

}  // namespace ply

std::vector<Triangle> load_ply_file(const std::string& filename);  // PLY twin of load_obj_file
std::vector<Triangle> load_obj_file(const std::string& filename);  // wavefront_obj.rs:15-44
std::vector<Sphere> load_obj_file_spheres(const std::string& filename, double radius);  // wavefront_obj.rs:46-66: a sphere per vertex

}  // namespace rayrs
