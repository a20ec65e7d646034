// formats.hpp -- the input formats either side of the hot path, host side, C++17 + zlib:
//   readPngRgb / filterObstacles   the PNG occupancy scenario's reader and colour filter
//                                   (demo/png_2d_scenario.hpp:50-69 FilterColor, :192-265 readAndFilterPng; libpng there)
//   ScenarioConfig                  OMPL-style .cfg files: [section] / key = value (demo/scenario_config.hpp:47-160)
//   readObjTriangles                triangle soups for the rigid-body scenario (the reference loads meshes through
//                                   assimp and fan-triangulates faces, demo/se3_rigid_body_scenario.hpp:164-204)
//   readColladaTriangles            the same from COLLADA (.dae) -- the format of the reference's SE(3) inputs
//                                   (OMPL's resources, named by the .cfg files): geometry library + scene graph with
//                                   node transforms, visited as demo/se3_rigid_body_scenario.hpp:73-133 does
//   readMeshTriangles               .obj / .dae by extension; recentre on the vertex mean as :181-193 does for the robot
// libpng and assimp are not on this machine; the decoders below are written from the format specifications
// (PNG: RFC 2083 -- chunk layout, zlib stream over IDAT, the five scan-line filters incl. Paeth; Wavefront OBJ: v / f
// records).  The transformations libpng applies in the reference are reproduced: palette -> RGB, 16 -> 8 bits (high
// byte), < 8 bits unpacked, alpha stripped; grey images are expanded to RGB (the reference reads three bytes per pixel
// and would mis-read them).  Interlaced PNGs are rejected.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <array>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <system_error>
#include <vector>

namespace mptg::formats {

struct Image {
    int width = 0, height = 0;
    std::vector<std::uint8_t> rgb;  // row-major, 3 bytes per pixel
};

namespace detail {
inline std::uint32_t be32(const unsigned char* p) { return (std::uint32_t)p[0] << 24 | (std::uint32_t)p[1] << 16 | (std::uint32_t)p[2] << 8 | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace detail

inline Image readPngRgb(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::system_error(errno, std::system_category(), "failed to open '" + path + "'");
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    static const unsigned char magic[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (file.size() < 8 || !std::equal(magic, magic + 8, file.begin())) throw std::invalid_argument(path + ": not a PNG file");
    std::uint32_t width = 0, height = 0;
    int depth = 0, colour = -1, interlace = 0;
    std::vector<unsigned char> idat;
    std::vector<std::array<std::uint8_t, 3>> palette;
    for (std::size_t pos = 8; pos + 12 <= file.size();) {
        const std::uint32_t len = detail::be32(&file[pos]);
        const std::string type(file.begin() + pos + 4, file.begin() + pos + 8);
        if (pos + 12 + (std::size_t)len > file.size()) throw std::invalid_argument(path + ": truncated chunk");
        const unsigned char* data = &file[pos + 8];
        if (type == "IHDR") {
            if (len != 13) throw std::invalid_argument(path + ": bad IHDR");
            width = detail::be32(data), height = detail::be32(data + 4);
            depth = data[8], colour = data[9], interlace = data[12];
        } else if (type == "PLTE") {
            for (std::uint32_t i = 0; i + 2 < len; i += 3) palette.push_back({data[i], data[i + 1], data[i + 2]});
        } else if (type == "IDAT") {
            idat.insert(idat.end(), data, data + len);
        } else if (type == "IEND") {
            break;
        }
        pos += 12 + (std::size_t)len;
    }
    if (colour < 0 || width == 0 || height == 0) throw std::invalid_argument(path + ": no IHDR");
    if (interlace != 0) throw std::invalid_argument(path + ": interlaced PNGs are not supported");
    int channels;
    switch (colour) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: throw std::invalid_argument(path + ": bad colour type");
    }
    if (!(depth == 8 || depth == 16 || ((colour == 0 || colour == 3) && (depth == 1 || depth == 2 || depth == 4))))
        throw std::invalid_argument(path + ": unsupported bit depth");
    const std::size_t bitsPerPixel = (std::size_t)channels * depth;
    const std::size_t rowBytes = ((std::size_t)width * bitsPerPixel + 7) / 8;
    const std::size_t bpp = std::max<std::size_t>(1, bitsPerPixel / 8);  // filter unit
    std::vector<unsigned char> raw((rowBytes + 1) * (std::size_t)height);
    {
        uLongf outLen = (uLongf)raw.size();
        const int rc = uncompress(raw.data(), &outLen, idat.data(), (uLong)idat.size());
        if (rc != Z_OK || outLen != raw.size()) throw std::invalid_argument(path + ": bad image data (zlib " + std::to_string(rc) + ")");
    }
    // undo the scan-line filters in place
    std::vector<unsigned char> prev(rowBytes, 0);
    Image img;
    img.width = (int)width, img.height = (int)height;
    img.rgb.resize((std::size_t)width * height * 3);
    for (std::uint32_t y = 0; y < height; ++y) {
        unsigned char* row = &raw[(rowBytes + 1) * (std::size_t)y];
        const int filter = row[0];
        unsigned char* cur = row + 1;
        for (std::size_t i = 0; i < rowBytes; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = cur[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += detail::paeth(a, b, c); break;
                default: throw std::invalid_argument(path + ": bad scan-line filter");
            }
            cur[i] = (unsigned char)v;
        }
        std::copy(cur, cur + rowBytes, prev.begin());
        // samples -> 8-bit RGB
        auto sample = [&](std::size_t px, int ch) -> int {  // value scaled like libpng's transformations
            if (depth == 8) return cur[px * channels + ch];
            if (depth == 16) return cur[(px * channels + ch) * 2];  // png_set_strip_16: the high byte
            const std::size_t bit = px * depth;                      // packed grey / palette index
            return (cur[bit / 8] >> (8 - depth - (int)(bit % 8))) & ((1 << depth) - 1);
        };
        for (std::uint32_t x = 0; x < width; ++x) {
            std::uint8_t* out = &img.rgb[((std::size_t)y * width + x) * 3];
            if (colour == 3) {
                const int idx = sample(x, 0);
                if ((std::size_t)idx >= palette.size()) throw std::invalid_argument(path + ": palette index out of range");
                out[0] = palette[idx][0], out[1] = palette[idx][1], out[2] = palette[idx][2];
            } else if (colour == 0 || colour == 4) {
                int g = sample(x, 0);
                if (depth < 8) g = g * 255 / ((1 << depth) - 1);
                out[0] = out[1] = out[2] = (std::uint8_t)g;
            } else {
                out[0] = (std::uint8_t)sample(x, 0), out[1] = (std::uint8_t)sample(x, 1), out[2] = (std::uint8_t)sample(x, 2);
            }
        }
    }
    return img;
}

// demo/png_2d_scenario.hpp:50-69
struct FilterColor {
    int r, g, b, tol;
    FilterColor(int r_, int g_, int b_, int tol_) : r(r_), g(g_), b(b_), tol(tol_) {}
    bool isObstacle(int pr, int pg, int pb) const {
        return !((pr < r - tol || pr > r + tol) || (pg < g - tol || pg > g + tol) || (pb < b - tol || pb > b + tol));
    }
};

// demo/png_2d_scenario.hpp:246-265: pixel (x, y) is an obstacle iff any filter matches; index y * width + x
inline std::vector<std::uint8_t> filterObstacles(const Image& img, const std::vector<FilterColor>& filters) {
    std::vector<std::uint8_t> obstacles((std::size_t)img.width * img.height);
    for (std::size_t i = 0; i < obstacles.size(); ++i) {
        const std::uint8_t* px = &img.rgb[i * 3];
        bool hit = false;
        for (const FilterColor& c : filters)
            if (c.isObstacle(px[0], px[1], px[2])) {
                hit = true;
                break;
            }
        obstacles[i] = hit ? 1 : 0;
    }
    return obstacles;
}

// demo/scenario_config.hpp:47-160 ([section], key = value, numeric / vector / angle-axis getters)
class ScenarioConfig {
    std::map<std::string, std::map<std::string, std::string>> properties_;

    static std::string trim(const std::string& s) {
        std::size_t b = 0, e = s.size();
        while (b < e && std::isspace((unsigned char)s[b])) ++b;
        while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
        return s.substr(b, e - b);
    }

public:
    explicit ScenarioConfig(const std::string& fileName) {
        std::ifstream in(fileName);
        if (!in) throw std::system_error(errno, std::system_category(), "failed to open '" + fileName + "'");
        std::string line, section;
        while (std::getline(in, line)) {
            const std::string t = trim(line);
            if (t.empty()) continue;
            if (t.front() == '[' && t.back() == ']') {
                section = trim(t.substr(1, t.size() - 2));
            } else if (const std::size_t eq = t.find('='); eq != std::string::npos && !section.empty()) {
                properties_[section][trim(t.substr(0, eq))] = trim(t.substr(eq + 1));
            }  // anything else: the reference logs a warning and carries on
        }
    }
    bool hasProp(const std::string& section, const std::string& name) const {
        const auto it = properties_.find(section);
        return it != properties_.end() && it->second.count(name) != 0;
    }
    void load(std::string& prop, const std::string& section, const std::string& name) const {
        const auto it = properties_.find(section);
        if (it == properties_.end()) throw std::invalid_argument("missing section [" + section + "]");
        const auto kv = it->second.find(name);
        if (kv == it->second.end()) throw std::invalid_argument("missing property [" + section + "] " + name);
        prop = kv->second;
    }
    void load(double& prop, const std::string& section, const std::string& name) const {
        std::string str;
        load(str, section, name);
        std::size_t pos = 0;
        prop = std::stod(str, &pos);
        if (pos != str.length()) throw std::invalid_argument("invalid numeric value: " + str);
    }
    void load(float& prop, const std::string& section, const std::string& name) const {
        std::string str;
        load(str, section, name);
        std::size_t pos = 0;
        prop = std::stof(str, &pos);
        if (pos != str.length()) throw std::invalid_argument("invalid numeric value: " + str);
    }
    // name.x / name.y / name.z
    template <typename Scalar>
    void loadVector3(Scalar* v, const std::string& section, const std::string& name) const {
        load(v[0], section, name + ".x"), load(v[1], section, name + ".y"), load(v[2], section, name + ".z");
    }
    // name.theta + name.axis.{x,y,z} -> quaternion (x, y, z, w), as Eigen::Quaternion(AngleAxis) does
    template <typename Scalar>
    void loadRotation(Scalar* xyzw, const std::string& section, const std::string& name) const {
        Scalar theta, axis[3];
        load(theta, section, name + ".theta");
        loadVector3(axis, section, name + ".axis");
        const Scalar s = std::sin(theta / 2), c = std::cos(theta / 2);
        xyzw[0] = axis[0] * s, xyzw[1] = axis[1] * s, xyzw[2] = axis[2] * s, xyzw[3] = c;
    }
    // an SE(3) state "name": rotation then translation (std::tuple<Quaternion, Vector3>, scenario_config.hpp:156-160),
    // in C-ABI order qx qy qz qw tx ty tz
    template <typename Scalar>
    void loadSE3(Scalar* state7, const std::string& section, const std::string& name) const {
        loadRotation(state7, section, name);
        loadVector3(state7 + 4, section, name);
    }
};

// Wavefront OBJ subset: "v x y z" and "f i j k ..." (i, i/t, i/t/n, i//n; negative = relative); faces with more than
// three corners are fan-triangulated (corner 0, i, i+1), as the reference does with assimp's faces
// (demo/se3_rigid_body_scenario.hpp:196-203).  -> nine floats per triangle
inline std::vector<float> readObjTriangles(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::system_error(errno, std::system_category(), "failed to open '" + path + "'");
    std::vector<std::array<float, 3>> verts;
    std::vector<float> tris;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tag;
        if (!(ls >> tag)) continue;
        if (tag == "v") {
            std::array<float, 3> v{};
            if (!(ls >> v[0] >> v[1] >> v[2])) throw std::invalid_argument(path + ": bad vertex record");
            verts.push_back(v);
        } else if (tag == "f") {
            std::vector<std::size_t> corner;
            std::string tok;
            while (ls >> tok) {
                const long idx = std::strtol(tok.c_str(), nullptr, 10);  // stops at '/'
                const long res = idx > 0 ? idx - 1 : (long)verts.size() + idx;
                if (idx == 0 || res < 0 || (std::size_t)res >= verts.size()) throw std::invalid_argument(path + ": face index out of range");
                corner.push_back((std::size_t)res);
            }
            for (std::size_t i = 1; i + 1 < corner.size(); ++i)
                for (std::size_t c : {corner[0], corner[i], corner[i + 1]}) tris.insert(tris.end(), verts[c].begin(), verts[c].end());
        }
    }
    return tris;
}

// ---------------------------------------------------------------------------------- COLLADA subset
namespace xml {
struct Node {
    std::string tag, text;
    std::map<std::string, std::string> attr;
    std::vector<Node> children;
    const Node* child(const std::string& t) const {
        for (const Node& c : children)
            if (c.tag == t) return &c;
        return nullptr;
    }
    std::string get(const std::string& k) const {
        auto it = attr.find(k);
        return it == attr.end() ? std::string() : it->second;
    }
};
// elements, attributes and character data; comments, processing instructions, DOCTYPE and CDATA markers are skipped,
// the five predefined entities are not expanded (none occurs in the numeric content read here)
inline Node parse(const std::string& src, const std::string& what) {
    Node root;
    std::vector<Node*> stack{&root};
    std::size_t i = 0;
    const std::size_t n = src.size();
    auto fail = [&](const char* msg) { return std::invalid_argument(what + ": malformed XML (" + msg + ")"); };
    while (i < n) {
        if (src[i] != '<') {
            const std::size_t j = src.find('<', i);
            stack.back()->text.append(src, i, (j == std::string::npos ? n : j) - i);
            i = j == std::string::npos ? n : j;
            continue;
        }
        if (src.compare(i, 4, "<!--") == 0) {
            const std::size_t j = src.find("-->", i + 4);
            if (j == std::string::npos) throw fail("unterminated comment");
            i = j + 3;
        } else if (src.compare(i, 2, "<?") == 0 || src.compare(i, 2, "<!") == 0) {
            const std::size_t j = src.find('>', i);
            if (j == std::string::npos) throw fail("unterminated declaration");
            i = j + 1;
        } else if (src.compare(i, 2, "</") == 0) {
            const std::size_t j = src.find('>', i);
            if (j == std::string::npos || stack.size() < 2) throw fail("unbalanced end tag");
            stack.pop_back();
            i = j + 1;
        } else {
            std::size_t j = i + 1;
            while (j < n && !std::isspace((unsigned char)src[j]) && src[j] != '>' && src[j] != '/') ++j;
            Node el;
            el.tag = src.substr(i + 1, j - i - 1);
            bool selfClosed = false;
            while (j < n && src[j] != '>') {
                if (src[j] == '/') {
                    selfClosed = true;
                    ++j;
                    continue;
                }
                if (std::isspace((unsigned char)src[j])) {
                    ++j;
                    continue;
                }
                const std::size_t eq = src.find('=', j);
                if (eq == std::string::npos) throw fail("attribute without value");
                std::string key = src.substr(j, eq - j);
                while (!key.empty() && std::isspace((unsigned char)key.back())) key.pop_back();
                std::size_t q = eq + 1;
                while (q < n && std::isspace((unsigned char)src[q])) ++q;
                if (q >= n || (src[q] != '"' && src[q] != '\'')) throw fail("unquoted attribute");
                const std::size_t end = src.find(src[q], q + 1);
                if (end == std::string::npos) throw fail("unterminated attribute");
                el.attr[key] = src.substr(q + 1, end - q - 1);
                j = end + 1;
            }
            if (j >= n) throw fail("unterminated tag");
            i = j + 1;
            stack.back()->children.push_back(std::move(el));
            if (!selfClosed) stack.push_back(&stack.back()->children.back());
        }
    }
    if (stack.size() != 1) throw fail("unclosed element");
    return root;
}
template <typename T>
std::vector<T> numbers(const std::string& text) {
    std::vector<T> out;
    const char* p = text.c_str();
    char* end;
    for (;;) {
        const double v = std::strtod(p, &end);
        if (end == p) break;
        out.push_back((T)v);
        p = end;
    }
    return out;
}
}  // namespace xml

namespace detail {
struct Mat4 {  // row-major affine transform
    double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    Mat4 operator*(const Mat4& o) const {
        Mat4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double a = 0;
                for (int k = 0; k < 4; ++k) a += m[i * 4 + k] * o.m[k * 4 + j];
                r.m[i * 4 + j] = a;
            }
        return r;
    }
    std::array<double, 3> apply(const double* v) const {
        return {m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3], m[4] * v[0] + m[5] * v[1] + m[6] * v[2] + m[7],
                m[8] * v[0] + m[9] * v[1] + m[10] * v[2] + m[11]};
    }
};
struct DaeGeometry {
    std::vector<double> positions;           // xyz
    std::vector<std::vector<std::size_t>> faces;  // position indices per polygon
};
}  // namespace detail

// COLLADA 1.4 / 1.5 subset, enough for rigid meshes: <library_geometries> (<mesh> with <source>/<float_array>,
// <vertices> POSITION input, <triangles> / <polylist> / <polygons> / <trifans> / <tristrips> primitives with interleaved <p>
// indices), <library_visual_scenes> (<node> trees with <matrix>, <translate>, <rotate>, <scale> in document order,
// <instance_geometry>, <instance_node> into <library_nodes>) and <asset><up_axis> (X_UP / Z_UP are turned to Y up at the
// root as assimp's importer does; <unit> is not applied, as in assimp).  Every instanced geometry is visited with the
// product of the transforms from the root down (demo/se3_rigid_body_scenario.hpp:96-133) and its polygons are fanned
// out round corner 0 (:118-125).  Files without a scene instance every geometry once, untransformed.
// shiftToCentre: subtract the mean of the visited vertices (:181-193; the reference's mean runs over assimp's joined
// vertex list, here over the positions each instanced geometry references -- the same set of points unless a position
// is shared by corners that differ in normal or texture coordinate, which assimp then counts more than once).
inline std::vector<float> readColladaTriangles(const std::string& path, bool shiftToCentre = false) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::system_error(errno, std::system_category(), "failed to open '" + path + "'");
    std::stringstream buf;
    buf << in.rdbuf();
    const xml::Node doc = xml::parse(buf.str(), path);
    const xml::Node* collada = doc.child("COLLADA");
    if (!collada) throw std::invalid_argument(path + ": not a COLLADA document");
    auto strip = [](std::string ref) { return !ref.empty() && ref[0] == '#' ? ref.substr(1) : ref; };

    std::map<std::string, detail::DaeGeometry> geometries;
    std::vector<std::string> geometryOrder;
    if (const xml::Node* lib = collada->child("library_geometries"))
        for (const xml::Node& g : lib->children) {
            if (g.tag != "geometry") continue;
            const xml::Node* mesh = g.child("mesh");
            if (!mesh) continue;  // splines, convex meshes: not triangle data
            std::map<std::string, const xml::Node*> sources;
            for (const xml::Node& c : mesh->children)
                if (c.tag == "source") sources[c.get("id")] = &c;
            std::string positionSource, verticesId;
            if (const xml::Node* v = mesh->child("vertices")) {
                verticesId = v->get("id");
                for (const xml::Node& inp : v->children)
                    if (inp.tag == "input" && inp.get("semantic") == "POSITION") positionSource = strip(inp.get("source"));
            }
            auto src = sources.find(positionSource);
            if (src == sources.end()) throw std::invalid_argument(path + ": geometry '" + g.get("id") + "' has no POSITION source");
            const xml::Node* fa = src->second->child("float_array");
            if (!fa) throw std::invalid_argument(path + ": POSITION source without <float_array>");
            detail::DaeGeometry geo;
            std::size_t stride = 3;
            if (const xml::Node* tc = src->second->child("technique_common"))
                if (const xml::Node* acc = tc->child("accessor"))
                    if (!acc->get("stride").empty()) stride = (std::size_t)std::strtoul(acc->get("stride").c_str(), nullptr, 10);
            const std::vector<double> raw = xml::numbers<double>(fa->text);
            if (stride < 3) throw std::invalid_argument(path + ": POSITION stride below 3");
            for (std::size_t i = 0; i + 3 <= raw.size(); i += stride) geo.positions.insert(geo.positions.end(), raw.begin() + i, raw.begin() + i + 3);
            for (const xml::Node& prim : mesh->children) {
                const bool tri = prim.tag == "triangles", plist = prim.tag == "polylist", polys = prim.tag == "polygons";
                const bool fans = prim.tag == "trifans", strips = prim.tag == "tristrips";
                if (!(tri || plist || polys || fans || strips)) continue;
                std::size_t inputs = 0, vOffset = 0;
                bool haveVertex = false;
                for (const xml::Node& inp : prim.children)
                    if (inp.tag == "input") {
                        const std::size_t off = (std::size_t)std::strtoul(inp.get("offset").c_str(), nullptr, 10);
                        inputs = std::max(inputs, off + 1);
                        if (inp.get("semantic") == "VERTEX" && strip(inp.get("source")) == verticesId) vOffset = off, haveVertex = true;
                    }
                if (!haveVertex) throw std::invalid_argument(path + ": primitive without a VERTEX input");
                auto corners = [&](const std::string& text) {
                    const std::vector<long> all = xml::numbers<long>(text);
                    std::vector<std::size_t> out;
                    for (std::size_t i = vOffset; i < all.size(); i += inputs) {
                        if (all[i] < 0 || (std::size_t)all[i] * 3 + 2 >= geo.positions.size()) throw std::invalid_argument(path + ": vertex index out of range");
                        out.push_back((std::size_t)all[i]);
                    }
                    return out;
                };
                if (tri || plist) {
                    const xml::Node* p = prim.child("p");
                    if (!p) continue;
                    const std::vector<std::size_t> idx = corners(p->text);
                    std::vector<long> vcount;
                    if (plist) {
                        const xml::Node* vc = prim.child("vcount");
                        if (!vc) throw std::invalid_argument(path + ": <polylist> without <vcount>");
                        vcount = xml::numbers<long>(vc->text);
                    } else {
                        vcount.assign(idx.size() / 3, 3);
                    }
                    std::size_t at = 0;
                    for (long c : vcount) {
                        if (c < 0 || at + (std::size_t)c > idx.size()) throw std::invalid_argument(path + ": <vcount> exceeds <p>");
                        geo.faces.emplace_back(idx.begin() + at, idx.begin() + at + c);
                        at += (std::size_t)c;
                    }
                } else {
                    for (const xml::Node& p : prim.children) {
                        if (p.tag != "p") continue;
                        const std::vector<std::size_t> idx = corners(p.text);
                        if (polys || fans) {
                            geo.faces.push_back(idx);  // a fan IS the triangulation used below
                        } else {
                            for (std::size_t i = 0; i + 2 < idx.size(); ++i)
                                geo.faces.push_back(i % 2 == 0 ? std::vector<std::size_t>{idx[i], idx[i + 1], idx[i + 2]}
                                                               : std::vector<std::size_t>{idx[i + 1], idx[i], idx[i + 2]});
                        }
                    }
                }
            }
            geometryOrder.push_back(g.get("id"));
            geometries[g.get("id")] = std::move(geo);
        }
    if (geometries.empty()) throw std::invalid_argument("mesh file '" + path + "' does not contain meshes");  // :175-176

    // instances: (geometry, accumulated transform)
    std::vector<std::pair<const detail::DaeGeometry*, detail::Mat4>> instances;
    std::map<std::string, const xml::Node*> libraryNodes;
    std::function<void(const xml::Node&)> indexNodes = [&](const xml::Node& nd) {
        for (const xml::Node& c : nd.children)
            if (c.tag == "node") {
                if (!c.get("id").empty()) libraryNodes[c.get("id")] = &c;
                indexNodes(c);
            }
    };
    if (const xml::Node* lib = collada->child("library_nodes")) indexNodes(*lib);
    const double pi = 3.14159265358979323846;
    std::function<void(const xml::Node&, detail::Mat4, int)> visit = [&](const xml::Node& nd, detail::Mat4 xf, int depth) {
        if (depth > 64) throw std::invalid_argument(path + ": node hierarchy too deep (cycle through <instance_node>?)");
        for (const xml::Node& c : nd.children) {
            if (c.tag == "matrix") {
                const std::vector<double> v = xml::numbers<double>(c.text);
                if (v.size() != 16) throw std::invalid_argument(path + ": <matrix> needs 16 numbers");
                detail::Mat4 m;
                std::copy(v.begin(), v.end(), m.m);
                xf = xf * m;
            } else if (c.tag == "translate") {
                const std::vector<double> v = xml::numbers<double>(c.text);
                if (v.size() != 3) throw std::invalid_argument(path + ": <translate> needs 3 numbers");
                detail::Mat4 m;
                m.m[3] = v[0], m.m[7] = v[1], m.m[11] = v[2];
                xf = xf * m;
            } else if (c.tag == "scale") {
                const std::vector<double> v = xml::numbers<double>(c.text);
                if (v.size() != 3) throw std::invalid_argument(path + ": <scale> needs 3 numbers");
                detail::Mat4 m;
                m.m[0] = v[0], m.m[5] = v[1], m.m[10] = v[2];
                xf = xf * m;
            } else if (c.tag == "rotate") {
                const std::vector<double> v = xml::numbers<double>(c.text);
                if (v.size() != 4) throw std::invalid_argument(path + ": <rotate> needs axis and angle");
                const double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                if (len > 0) {
                    const double x = v[0] / len, y = v[1] / len, z = v[2] / len, a = v[3] * pi / 180.0, cs = std::cos(a), sn = std::sin(a), t = 1 - cs;
                    detail::Mat4 m;
                    m.m[0] = t * x * x + cs, m.m[1] = t * x * y - sn * z, m.m[2] = t * x * z + sn * y;
                    m.m[4] = t * x * y + sn * z, m.m[5] = t * y * y + cs, m.m[6] = t * y * z - sn * x;
                    m.m[8] = t * x * z - sn * y, m.m[9] = t * y * z + sn * x, m.m[10] = t * z * z + cs;
                    xf = xf * m;
                }
            }
        }
        for (const xml::Node& c : nd.children) {
            if (c.tag == "instance_geometry") {
                auto it = geometries.find(strip(c.get("url")));
                if (it == geometries.end()) throw std::invalid_argument(path + ": <instance_geometry> of unknown geometry " + c.get("url"));
                instances.push_back({&it->second, xf});
            } else if (c.tag == "instance_node") {
                auto it = libraryNodes.find(strip(c.get("url")));
                if (it == libraryNodes.end()) throw std::invalid_argument(path + ": <instance_node> of unknown node " + c.get("url"));
                visit(*it->second, xf, depth + 1);
            } else if (c.tag == "node") {
                visit(c, xf, depth + 1);
            }
        }
    };
    detail::Mat4 root;
    if (const xml::Node* asset = collada->child("asset"))
        if (const xml::Node* up = asset->child("up_axis")) {
            std::string axis = up->text;
            axis.erase(std::remove_if(axis.begin(), axis.end(), [](unsigned char ch) { return std::isspace(ch); }), axis.end());
            if (axis == "Z_UP") {  // (x, y, z) -> (x, z, -y)
                const double m[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 1};
                std::copy(m, m + 16, root.m);
            } else if (axis == "X_UP") {  // (x, y, z) -> (-y, x, z)
                const double m[16] = {0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
                std::copy(m, m + 16, root.m);
            }
        }
    const xml::Node* sceneNode = nullptr;
    if (const xml::Node* lib = collada->child("library_visual_scenes")) {
        std::string want;
        if (const xml::Node* sc = collada->child("scene"))
            if (const xml::Node* ivs = sc->child("instance_visual_scene")) want = strip(ivs->get("url"));
        for (const xml::Node& vs : lib->children)
            if (vs.tag == "visual_scene" && (!sceneNode || vs.get("id") == want)) sceneNode = &vs;
    }
    if (sceneNode) visit(*sceneNode, root, 0);
    if (instances.empty())
        for (const std::string& id : geometryOrder) instances.push_back({&geometries[id], root});

    std::array<double, 3> centre{0, 0, 0};
    if (shiftToCentre) {
        std::size_t count = 0;
        for (auto& [geo, xf] : instances) {
            std::vector<std::uint8_t> used(geo->positions.size() / 3, 0);
            for (auto& f : geo->faces)
                for (std::size_t c : f) used[c] = 1;
            for (std::size_t v = 0; v < used.size(); ++v)
                if (used[v]) {
                    const auto p = xf.apply(&geo->positions[3 * v]);
                    for (int k = 0; k < 3; ++k) centre[k] += p[k];
                    ++count;
                }
        }
        if (count)
            for (double& c : centre) c /= (double)count;
    }
    std::vector<float> tris;
    for (auto& [geo, xf] : instances)
        for (auto& f : geo->faces) {
            if (f.size() < 3) continue;  // :109-110
            for (std::size_t i = 1; i + 1 < f.size(); ++i)
                for (std::size_t c : {f[0], f[i], f[i + 1]}) {
                    const auto p = xf.apply(&geo->positions[3 * c]);
                    for (int k = 0; k < 3; ++k) tris.push_back((float)(p[k] - centre[k]));
                }
        }
    return tris;
}

// recentre a triangle soup on the mean of its corners' distinct positions (for OBJ input; the COLLADA reader does it itself)
inline void recentreTriangles(std::vector<float>& tris) {
    std::vector<std::array<float, 3>> pts;
    for (std::size_t i = 0; i + 3 <= tris.size(); i += 3) pts.push_back({tris[i], tris[i + 1], tris[i + 2]});
    std::sort(pts.begin(), pts.end());
    pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
    double c[3] = {0, 0, 0};
    for (auto& p : pts)
        for (int k = 0; k < 3; ++k) c[k] += p[k];
    if (!pts.empty())
        for (double& v : c) v /= (double)pts.size();
    for (std::size_t i = 0; i < tris.size(); ++i) tris[i] = (float)(tris[i] - c[i % 3]);
}

// the mesh file a .cfg names (robot / world): COLLADA or OBJ by extension
inline std::vector<float> readMeshTriangles(const std::string& path, bool shiftToCentre = false) {
    std::string ext = path.size() >= 4 ? path.substr(path.size() - 4) : std::string();
    std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
    if (ext == ".dae") return readColladaTriangles(path, shiftToCentre);
    if (ext == ".obj") {
        std::vector<float> tris = readObjTriangles(path);
        if (shiftToCentre) recentreTriangles(tris);
        return tris;
    }
    throw std::invalid_argument("mesh file '" + path + "': only .dae and .obj are read here (the reference goes through assimp)");
}

}  // namespace mptg::formats
