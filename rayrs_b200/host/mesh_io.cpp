// PLY / OBJ loaders of the host mirror — see mesh_io.hpp for what follows the reference sketch
// (ply/src/lib.rs, commented out in the reference) and what completes it.
#include "mesh_io.hpp"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rayrs {
namespace ply {

namespace {

// Rust's str::split(' '): consecutive separators yield empty pieces
std::vector<std::string> split_space(const std::string& s) {
    std::vector<std::string> out;
    size_t start = 0;
    for (;;) {
        size_t p = s.find(' ', start);
        if (p == std::string::npos) {
            out.push_back(s.substr(start));
            return out;
        }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
}

bool host_is_little_endian() {
    const uint16_t v = 1;
    uint8_t b;
    std::memcpy(&b, &v, 1);
    return b == 1;
}

struct Cursor {
    const std::string& s;
    size_t pos;
    void need(size_t n) const {
        if (pos + n > s.size()) throw IoError(IoError::UnexpectedEof, "unexpected EOF");
    }
};

double read_binary(Cursor& c, PlyPropertyType t, bool swap) {
    const size_t n = property_type_size(t);
    c.need(n);
    unsigned char b[8];
    std::memcpy(b, c.s.data() + c.pos, n);
    c.pos += n;
    if (swap)
        for (size_t i = 0; i < n / 2; ++i) std::swap(b[i], b[n - 1 - i]);
    switch (t) {
        case PlyPropertyType::Char: { int8_t v; std::memcpy(&v, b, 1); return v; }
        case PlyPropertyType::Uchar: { uint8_t v; std::memcpy(&v, b, 1); return v; }
        case PlyPropertyType::Short: { int16_t v; std::memcpy(&v, b, 2); return v; }
        case PlyPropertyType::Ushort: { uint16_t v; std::memcpy(&v, b, 2); return v; }
        case PlyPropertyType::Int: { int32_t v; std::memcpy(&v, b, 4); return v; }
        case PlyPropertyType::Uint: { uint32_t v; std::memcpy(&v, b, 4); return v; }
        case PlyPropertyType::Float: { float v; std::memcpy(&v, b, 4); return v; }
        default: { double v; std::memcpy(&v, b, 8); return v; }
    }
}

double read_ascii(Cursor& c, PlyPropertyType t) {
    const std::string& s = c.s;
    while (c.pos < s.size() && (s[c.pos] == ' ' || s[c.pos] == '\n' || s[c.pos] == '\r' || s[c.pos] == '\t')) ++c.pos;
    if (c.pos >= s.size()) throw IoError(IoError::UnexpectedEof, "unexpected EOF");
    const char* begin = s.c_str() + c.pos;
    char* end = nullptr;
    errno = 0;
    double v;
    if (t == PlyPropertyType::Float || t == PlyPropertyType::Double) {
        v = std::strtod(begin, &end);
        if (t == PlyPropertyType::Float) v = (double)(float)v;  // a `float` property holds 32 bits in every format
    } else {
        v = (double)std::strtoll(begin, &end, 10);
    }
    if (end == begin) throw IoError(IoError::InvalidData, "invalid value in element data");
    c.pos += (size_t)(end - begin);
    return v;
}

void write_scalar(std::string& out, const void* p, size_t n, bool swap) {
    const char* b = static_cast<const char*>(p);
    if (!swap) {
        out.append(b, n);
    } else {
        for (size_t i = 0; i < n; ++i) out.push_back(b[n - 1 - i]);
    }
}

}  // namespace

PlyFormat format_from_string(const std::string& s) {
    if (s == "ascii") return PlyFormat::Ascii;
    if (s == "binary_big_endian") return PlyFormat::BinaryBigEndian;
    if (s == "binary_little_endian") return PlyFormat::BinaryLittleEndian;
    throw IoError(IoError::InvalidData, "invalid format: " + s);
}

PlyPropertyType property_type_from_string(const std::string& s) {
    // the eight names of the sketch, plus the sized aliases that files in the wild use
    if (s == "char" || s == "int8") return PlyPropertyType::Char;
    if (s == "uchar" || s == "uint8") return PlyPropertyType::Uchar;
    if (s == "short" || s == "int16") return PlyPropertyType::Short;
    if (s == "ushort" || s == "uint16") return PlyPropertyType::Ushort;
    if (s == "int" || s == "int32") return PlyPropertyType::Int;
    if (s == "uint" || s == "uint32") return PlyPropertyType::Uint;
    if (s == "float" || s == "float32") return PlyPropertyType::Float;
    if (s == "double" || s == "float64") return PlyPropertyType::Double;
    throw IoError(IoError::InvalidData, "invalid property type: " + s);
}

size_t property_type_size(PlyPropertyType t) {
    switch (t) {
        case PlyPropertyType::Char:
        case PlyPropertyType::Uchar: return 1;
        case PlyPropertyType::Short:
        case PlyPropertyType::Ushort: return 2;
        case PlyPropertyType::Int:
        case PlyPropertyType::Uint:
        case PlyPropertyType::Float: return 4;
        default: return 8;
    }
}

bool PlyKeyword::operator==(const PlyKeyword& o) const {
    return tag == o.tag && format == o.format && version == o.version && comment == o.comment && name == o.name &&
           length == o.length && typ == o.typ && lentype == o.lentype && elemtype == o.elemtype;
}

PlyKeyword PlyKeyword::from_line(const std::string& line_in) {
    std::string line = line_in;
    if (!line.empty() && line.back() == '\r') line.pop_back();  // BufRead::lines strips "\r\n" too
    std::vector<std::string> parts = split_space(line);
    const std::string head = parts[0];
    std::vector<std::string> rest(parts.begin() + 1, parts.end());
    PlyKeyword k;
    if (head == "ply") {
        k.tag = Ply;
    } else if (head == "format") {  // parse_format
        if (rest.size() != 2) throw IoError(IoError::InvalidData, "invalid format specifier");
        if (rest[1] != "1.0") throw IoError(IoError::InvalidData, "invalid version: " + rest[1] + ", valid versions: 1.0");
        k.tag = Format;
        k.format = format_from_string(rest[0]);
        k.version = rest[1];
    } else if (head == "comment" || head == "obj_info") {  // obj_info: not in the sketch, common in real files
        k.tag = Comment;
        k.comment = line;
    } else if (head == "element") {  // parse_element
        if (rest.size() != 2) throw IoError(IoError::InvalidData, "invalid element");
        const std::string& num = rest[1];
        if (num.empty() || num.find_first_not_of("0123456789") != std::string::npos)
            throw IoError(IoError::InvalidData, "invalid digit found in string");
        k.tag = Element;
        k.name = rest[0];
        k.length = (size_t)std::strtoull(num.c_str(), nullptr, 10);
    } else if (head == "property") {  // parse_property
        if (rest.size() == 2) {
            k.tag = Property;
            k.typ = property_type_from_string(rest[0]);
            k.name = rest[1];
        } else if (rest.size() == 4 && rest[0] == "list") {
            k.tag = ListProperty;
            k.lentype = property_type_from_string(rest[1]);
            k.elemtype = property_type_from_string(rest[2]);
            k.name = rest[3];
        } else {
            throw IoError(IoError::InvalidData, "invalid property");
        }
    } else if (head == "end_header") {
        k.tag = EndHeader;
    } else {
        throw IoError(IoError::InvalidInput, "unknown ply keyword: " + head);
    }
    return k;
}

int PlyElement::property_index(const std::string& n) const {
    for (size_t i = 0; i < properties.size(); ++i)
        if (properties[i].name == n) return (int)i;
    return -1;
}

void PlyHeader::add_element(const std::string& name, size_t length) {
    PlyElement e;
    e.name = name;
    e.length = length;
    elements.push_back(std::move(e));
}
void PlyHeader::add_property(const std::string& name, PlyPropertyType typ) {
    PlyProperty p;
    p.name = name;
    p.typ = typ;
    elements.back().properties.push_back(p);
}
void PlyHeader::add_list_property(const std::string& name, PlyPropertyType lentype, PlyPropertyType elemtype) {
    PlyProperty p;
    p.name = name;
    p.is_list = true;
    p.lentype = lentype;
    p.typ = elemtype;
    elements.back().properties.push_back(p);
}

void PlyHeaderParser::handle_input(const PlyKeyword& inp) {
    switch (state) {
        case Start:
            if (inp.tag != PlyKeyword::Ply) throw IoError(IoError::Other, "expected 'ply' identifier");
            state = Format;
            return;
        case Format:
            if (inp.tag != PlyKeyword::Format) throw IoError(IoError::Other, "expected format specification");
            header.version = inp.version;
            header.format = inp.format;
            state = StartElement;
            return;
        case StartElement:
            if (inp.tag == PlyKeyword::Comment) {
                header.add_comment(inp.comment);
            } else if (inp.tag == PlyKeyword::Element) {
                header.add_element(inp.name, inp.length);
                state = NewElement;
            } else {
                throw IoError(IoError::Other, "expected 'element' keyword");
            }
            return;
        case NewElement:
            if (inp.tag == PlyKeyword::Comment) {
                header.add_comment(inp.comment);
            } else if (inp.tag == PlyKeyword::Property) {
                header.add_property(inp.name, inp.typ);
                state = InElement;
            } else if (inp.tag == PlyKeyword::ListProperty) {
                header.add_list_property(inp.name, inp.lentype, inp.elemtype);
                state = InElement;
            } else {
                throw IoError(IoError::Other, "expected 'property' keyword");
            }
            return;
        case InElement:
            if (inp.tag == PlyKeyword::Comment) {
                header.add_comment(inp.comment);
            } else if (inp.tag == PlyKeyword::Element) {
                header.add_element(inp.name, inp.length);
                state = NewElement;
            } else if (inp.tag == PlyKeyword::Property) {
                header.add_property(inp.name, inp.typ);
            } else if (inp.tag == PlyKeyword::ListProperty) {
                header.add_list_property(inp.name, inp.lentype, inp.elemtype);
            } else if (inp.tag == PlyKeyword::EndHeader) {
                state = End;
            } else {
                throw IoError(IoError::Other, "expected properties or new element");
            }
            return;
        case End:
            throw Panic("Parser in end state cannot accept more input.");
    }
}

Ply Ply::parse(const std::string& bytes) {
    PlyHeaderParser machine;
    size_t pos = 0;
    while (machine.state != PlyHeaderParser::End) {
        if (pos >= bytes.size()) throw IoError(IoError::InvalidInput, "unexpected EOF");
        size_t nl = bytes.find('\n', pos);
        std::string line = bytes.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = nl == std::string::npos ? bytes.size() : nl + 1;
        machine.handle_input(PlyKeyword::from_line(line));
    }
    Ply out;
    out.header = std::move(machine.header);
    // ---- element data (not in the sketch) ----
    Cursor c{bytes, pos};
    const bool ascii = out.header.format == PlyFormat::Ascii;
    const bool swap = !ascii && ((out.header.format == PlyFormat::BinaryLittleEndian) != host_is_little_endian());
    for (PlyElement& e : out.header.elements) {
        const size_t np = e.properties.size();
        e.scalars.assign(np, {});
        e.lists.assign(np, {});
        for (size_t p = 0; p < np; ++p) {
            if (e.properties[p].is_list) e.lists[p].resize(e.length);
            else e.scalars[p].resize(e.length);
        }
        for (size_t i = 0; i < e.length; ++i) {
            for (size_t p = 0; p < np; ++p) {
                const PlyProperty& pr = e.properties[p];
                if (!pr.is_list) {
                    e.scalars[p][i] = ascii ? read_ascii(c, pr.typ) : read_binary(c, pr.typ, swap);
                } else {
                    const double len = ascii ? read_ascii(c, pr.lentype) : read_binary(c, pr.lentype, swap);
                    if (len < 0 || len > 1e6) throw IoError(IoError::InvalidData, "invalid list length");
                    std::vector<double>& l = e.lists[p][i];
                    l.resize((size_t)len);
                    for (double& v : l) v = ascii ? read_ascii(c, pr.typ) : read_binary(c, pr.typ, swap);
                }
            }
        }
    }
    return out;
}

Ply Ply::load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw IoError(IoError::NotFound, "No such file or directory: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
}

const PlyElement* Ply::element(const std::string& name) const {
    for (const PlyElement& e : header.elements)
        if (e.name == name) return &e;
    return nullptr;
}

std::vector<Triangle> Ply::triangles() const {
    const PlyElement* v = element("vertex");
    const PlyElement* f = element("face");
    if (!v || !f) throw IoError(IoError::InvalidData, "ply file has no vertex/face elements");
    const int ix = v->property_index("x"), iy = v->property_index("y"), iz = v->property_index("z");
    if (ix < 0 || iy < 0 || iz < 0 || v->properties[ix].is_list || v->properties[iy].is_list || v->properties[iz].is_list)
        throw IoError(IoError::InvalidData, "vertex element has no scalar x/y/z properties");
    int il = f->property_index("vertex_indices");
    if (il < 0) il = f->property_index("vertex_index");
    if (il < 0 || !f->properties[il].is_list) throw IoError(IoError::InvalidData, "face element has no vertex_indices list");
    std::vector<Triangle> tris;
    tris.reserve(f->length);
    auto vertex = [&](double idx) {
        if (idx < 0 || idx >= (double)v->length) throw IoError(IoError::InvalidData, "face index out of range");
        const size_t k = (size_t)idx;
        return Vec3(v->scalars[ix][k], v->scalars[iy][k], v->scalars[iz][k]);
    };
    for (size_t i = 0; i < f->length; ++i) {
        const std::vector<double>& l = f->lists[il][i];
        for (size_t k = 1; k + 1 < l.size(); ++k) tris.push_back(Triangle{vertex(l[0]), vertex(l[k]), vertex(l[k + 1])});
    }
    return tris;
}

void write_ply(const std::string& path, const std::vector<float>& xyz, const std::vector<int32_t>& faces3, PlyFormat format) {
    if (xyz.size() % 3 || faces3.size() % 3) throw Panic("write_ply: xyz and faces must hold triples");
    const size_t nv = xyz.size() / 3, nf = faces3.size() / 3;
    std::string out = "ply\nformat ";
    out += format == PlyFormat::Ascii ? "ascii" : (format == PlyFormat::BinaryBigEndian ? "binary_big_endian" : "binary_little_endian");
    out += " 1.0\ncomment rayrs_b200 synthetic mesh\nelement vertex " + std::to_string(nv) +
           "\nproperty float x\nproperty float y\nproperty float z\nelement face " + std::to_string(nf) +
           "\nproperty list uchar int vertex_indices\nend_header\n";
    if (format == PlyFormat::Ascii) {
        char buf[128];
        for (size_t i = 0; i < nv; ++i) {
            std::snprintf(buf, sizeof buf, "%.9g %.9g %.9g\n", xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
            out += buf;
        }
        for (size_t i = 0; i < nf; ++i) {
            std::snprintf(buf, sizeof buf, "3 %d %d %d\n", faces3[3 * i], faces3[3 * i + 1], faces3[3 * i + 2]);
            out += buf;
        }
    } else {
        const bool swap = (format == PlyFormat::BinaryLittleEndian) != host_is_little_endian();
        out.reserve(out.size() + nv * 12 + nf * 13);
        for (float v : xyz) write_scalar(out, &v, 4, swap);
        for (size_t i = 0; i < nf; ++i) {
            out.push_back((char)3);
            for (int k = 0; k < 3; ++k) write_scalar(out, &faces3[3 * i + k], 4, swap);
        }
    }
    std::ofstream f(path, std::ios::binary);
    if (!f) throw IoError(IoError::Other, "cannot open for writing: " + path);
    f.write(out.data(), (std::streamsize)out.size());
}

}  // namespace ply

std::vector<Object> Object::from_triangles(const std::vector<Triangle>& tris, Material mat, Emission emission) {
    std::vector<Object> out;
    out.reserve(tris.size());
    for (const Triangle& t : tris) out.push_back(Object::triangle(t.p1, t.p2, t.p3, mat, emission));
    return out;
}

std::vector<Object> Object::from_spheres(const std::vector<Sphere>& spheres, Material mat, Emission emission) {
    std::vector<Object> out;
    out.reserve(spheres.size());
    for (const Sphere& sp : spheres) out.push_back(Object::sphere(sp.radius, sp.origin, mat, emission));
    return out;
}

std::vector<Triangle> load_ply_file(const std::string& filename) { return ply::Ply::load(filename).triangles(); }

// wavefront_obj::load_obj_file (wavefront_obj.rs:15-44): "v x y z" and "f i j k" lines split on single
// spaces, 1-based indices, anything else ignored; malformed numbers / missing fields panic (unwrap / index).
std::vector<Triangle> load_obj_file(const std::string& filename) {
    std::ifstream f(filename);
    if (!f) throw IoError(IoError::NotFound, "No such file or directory: " + filename);
    std::vector<Vec3> vertices;
    std::vector<Triangle> triangles;
    std::string text;
    auto parse_f64 = [](const std::string& s) {
        char* end = nullptr;
        double v = std::strtod(s.c_str(), &end);
        if (s.empty() || end != s.c_str() + s.size()) throw Panic("called `Result::unwrap()` on an `Err` value: ParseFloatError");
        return v;
    };
    auto parse_usize = [](const std::string& s) {
        if (s.empty() || s.find_first_not_of("0123456789") != std::string::npos)
            throw Panic("called `Result::unwrap()` on an `Err` value: ParseIntError");
        return (size_t)std::strtoull(s.c_str(), nullptr, 10);
    };
    while (std::getline(f, text)) {
        if (!text.empty() && text.back() == '\r') text.pop_back();
        std::vector<std::string> v = ply::split_space(text);
        if (v[0] == "v") {
            if (v.size() < 4) throw Panic("index out of bounds");
            vertices.push_back(Vec3(parse_f64(v[1]), parse_f64(v[2]), parse_f64(v[3])));
        } else if (v[0] == "f") {
            if (v.size() < 4) throw Panic("index out of bounds");
            size_t i = parse_usize(v[1]), j = parse_usize(v[2]), k = parse_usize(v[3]);
            if (i == 0 || j == 0 || k == 0 || i > vertices.size() || j > vertices.size() || k > vertices.size())
                throw Panic("index out of bounds");
            triangles.push_back(Triangle{vertices[i - 1], vertices[j - 1], vertices[k - 1]});
        }
    }
    return triangles;
}

// wavefront_obj::load_obj_file_spheres (wavefront_obj.rs:46-66): one sphere of `radius` per "v x y z" line
std::vector<Sphere> load_obj_file_spheres(const std::string& filename, double radius) {
    std::ifstream f(filename);
    if (!f) throw IoError(IoError::NotFound, "No such file or directory: " + filename);
    std::vector<Sphere> spheres;
    std::string text;
    auto parse_f64 = [](const std::string& s) {
        char* end = nullptr;
        double v = std::strtod(s.c_str(), &end);
        if (s.empty() || end != s.c_str() + s.size()) throw Panic("called `Result::unwrap()` on an `Err` value: ParseFloatError");
        return v;
    };
    while (std::getline(f, text)) {
        if (!text.empty() && text.back() == '\r') text.pop_back();
        std::vector<std::string> v = ply::split_space(text);
        if (v[0] == "v") {
            if (v.size() < 4) throw Panic("index out of bounds");
            spheres.push_back(Sphere{radius, Vec3(parse_f64(v[1]), parse_f64(v[2]), parse_f64(v[3]))});
        }
    }
    return spheres;
}

}  // namespace rayrs
