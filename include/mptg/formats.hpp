// formats.hpp -- the input formats either side of the hot path, host side, C++17 + zlib:
//   readPngRgb / filterObstacles   the PNG occupancy scenario's reader and colour filter
//                                   (demo/png_2d_scenario.hpp:50-69 FilterColor, :192-265 readAndFilterPng; libpng there)
//   ScenarioConfig                  OMPL-style .cfg files: [section] / key = value (demo/scenario_config.hpp:47-160)
//   readObjTriangles                triangle soups for the rigid-body scenario (the reference loads meshes through
//                                   assimp and fan-triangulates faces, demo/se3_rigid_body_scenario.hpp:164-204)
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

}  // namespace mptg::formats
