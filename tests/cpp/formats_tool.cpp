// formats_tool.cpp -- command-line front end of include/mptg/formats.hpp for tests/test_formats.py
//   formats_tool png FILE OUT      OUT: int32 width, int32 height, then width*height*3 bytes RGB
//   formats_tool occ FILE OUT      OUT: int32 width, int32 height, then width*height bytes (1 = obstacle) with the colour
//                                  filters of demo/png_2d_planning.cpp:69-72
//   formats_tool obj FILE OUT      OUT: float32 triangles, nine per triangle
//   formats_tool dae FILE OUT [centre]   the same from a COLLADA file (optionally recentred on the vertex mean)
//   formats_tool cfg FILE          prints the SE(3) start / goal states, the volume and the mesh names of an OMPL .cfg
#include <cstdio>
#include <cstring>
#include <fstream>

#include "mptg/formats.hpp"

using namespace mptg::formats;

int main(int argc, char** argv) {
    try {
        if (argc < 3) throw std::invalid_argument("usage: formats_tool png|occ|obj|cfg FILE [OUT]");
        const std::string cmd = argv[1], file = argv[2];
        if (cmd == "png" || cmd == "occ") {
            const Image img = readPngRgb(file);
            std::ofstream out(argv[3], std::ios::binary);
            const std::int32_t hdr[2] = {img.width, img.height};
            out.write((const char*)hdr, sizeof hdr);
            if (cmd == "png") {
                out.write((const char*)img.rgb.data(), (std::streamsize)img.rgb.size());
            } else {
                const std::vector<FilterColor> filters{{126, 106, 61, 15}, {61, 53, 6, 15}, {255, 255, 255, 5}};
                const auto occ = filterObstacles(img, filters);
                out.write((const char*)occ.data(), (std::streamsize)occ.size());
            }
        } else if (cmd == "obj") {
            const auto tris = readObjTriangles(file);
            std::ofstream out(argv[3], std::ios::binary);
            out.write((const char*)tris.data(), (std::streamsize)(tris.size() * sizeof(float)));
            std::printf("%zu triangles\n", tris.size() / 9);
        } else if (cmd == "dae" || cmd == "mesh") {
            const bool centre = argc > 4 && std::string(argv[4]) == "centre";
            const auto tris = cmd == "dae" ? readColladaTriangles(file, centre) : readMeshTriangles(file, centre);
            std::ofstream out(argv[3], std::ios::binary);
            out.write((const char*)tris.data(), (std::streamsize)(tris.size() * sizeof(float)));
            std::printf("%zu triangles\n", tris.size() / 9);
        } else if (cmd == "cfg") {
            ScenarioConfig cfg(file);
            double start[7], goal[7], vmin[3], vmax[3];
            std::string world, robot;
            cfg.loadSE3(start, "problem", "start");
            cfg.loadSE3(goal, "problem", "goal");
            cfg.loadVector3(vmin, "problem", "volume.min");
            cfg.loadVector3(vmax, "problem", "volume.max");
            cfg.load(world, "problem", "world");
            cfg.load(robot, "problem", "robot");
            std::printf("world=%s robot=%s\n", world.c_str(), robot.c_str());
            std::printf("start");
            for (double v : start) std::printf(" %.17g", v);
            std::printf("\ngoal");
            for (double v : goal) std::printf(" %.17g", v);
            std::printf("\nvolume %.17g %.17g %.17g %.17g %.17g %.17g\n", vmin[0], vmin[1], vmin[2], vmax[0], vmax[1], vmax[2]);
            if (cfg.hasProp("planner", "rrt.range")) {
                double range;
                cfg.load(range, "planner", "rrt.range");
                std::printf("range %.17g\n", range);
            }
            bool threw = false;
            try {
                double d;
                cfg.load(d, "problem", "no.such.key");
            } catch (const std::invalid_argument&) {
                threw = true;
            }
            std::printf("missing-key-throws %d\n", (int)threw);
        } else {
            throw std::invalid_argument("unknown command " + cmd);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "formats_tool: %s\n", e.what());
        return 1;
    }
    return 0;
}
