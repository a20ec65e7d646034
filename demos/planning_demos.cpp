// planning_demos.cpp -- the four planning configurations of BASELINE.json (configs[0..3]) on the
// wave planners, mirroring the reference's demo mains:
//   holonomic_2d_point  demo/holonomic_2d_point_planning.cpp:57-102   (PRRT, circles + rectangles)
//   png_2d              demo/png_2d_planning.cpp:60-105               (PRRT*, occupancy grid)
//   se3_rigid_body      demo/se3_rigid_body_planning.cpp:155-233      (PRRT*, mesh vs mesh, DiscreteMotionValidator)
//   link_manipulator    demo/link_manipulator_planning.cpp:59-90      (PPRM, N-link planar arm)
// and, on request (--demo nao_cup), the reference's fifth demo:
//   nao_cup             demo/nao_cup_planning.cpp:155-215             (PRRT*, 10 joints, spheres and capsules)
// Usage: planning_demos [--all | --demo NAME] [--time-ms T] [--check] [--map file.png|file.pgm] [--seed S]
// Prints, per demo, time to first solution, node count and path cost ("solve time" of BASELINE.json).
// The reference's PNG and OMPL meshes are not redistributable inputs of this repository: the map is
// a synthetic one of the same size unless --map gives a binary PGM (tools/png_to_pgm.py converts the
// reference image with the reference's colour filters), the meshes are procedural bent tubes.
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <string>

#include "mptg/formats.hpp"
#include "mptg/planner.hpp"
#include "mptg/scenarios.hpp"

using namespace mptg;
using Clock = std::chrono::steady_clock;

struct Options {
    std::size_t nodes = 0;  // > 0: run until the graph has this many nodes (the reference's -n flag, se3_rigid_body_planning.cpp:80-153)
    double timeMs = 5000;
    bool check = false;
    std::string map;
    std::string cfg;  // se3_rigid_body: an OMPL-style problem file (demo/se3_rigid_body_planning.cpp:240-262) naming .dae / .obj meshes
    std::uint64_t seed = 2026;
    bool devicePrrt = false;  // also run Planner<Scenario, PRRT<device_resident>>: tree and sampling on the GPU
};

static int failures = 0;

template <typename Planner, typename Scenario>
void report(const char* name, const char* algo, Planner& planner, const Scenario& scenario, double firstSolutionS, double totalS,
            const Options& opt) {
    using State = typename Scenario::State;
    std::vector<State> path = planner.solution();
    double cost = 0;
    for (std::size_t i = 1; i < path.size(); ++i) cost += scenario.space().distance(path[i - 1], path[i]);
    planner.printStats();
    bool ok = planner.solved() && path.size() >= 2;
    if (opt.check && ok) {
        Context ctx;
        Geometry g = scenario.makeGeometry(ctx);
        std::vector<State> from(path.begin(), path.end() - 1), to(path.begin() + 1, path.end());
        std::vector<std::uint8_t> v(from.size());
        auto desc = scenario.space().desc();
        double step = 0;
        if constexpr (impl::has_link_step<Scenario>::value) step = scenario.linkStep();
        g.link(&desc, from.data(), to.data(), (std::uint32_t)from.size(), step, v.data());
        for (auto x : v) ok = ok && x == 1;
    }
    std::printf("%s %s [%s]: solved=%d first solution after %.4f s (ran %.3f s, %.0f nodes/s), %zu nodes, %zu waypoints, path cost %.4f\n",
                ok ? "OK" : "FAILED", name, algo, (int)planner.solved(), firstSolutionS, totalS, planner.size() / totalS, planner.size(), path.size(), cost);
    if (!ok) ++failures;
}

// scenario.valid(q) / scenario.link(a,b) for single states through the batched back-end
template <typename Scenario>
struct Probe {
    const Scenario& sc;
    Context ctx;
    Geometry g;
    explicit Probe(const Scenario& s) : sc(s), g(s.makeGeometry(ctx)) {}
    bool valid(const typename Scenario::State& q) {
        std::uint8_t ok = 0;
        g.valid(&q, 1, &ok);
        return ok != 0;
    }
    bool link(const typename Scenario::State& a, const typename Scenario::State& b) {
        std::uint8_t ok = 0;
        auto desc = sc.space().desc();
        double step = 0;
        if constexpr (impl::has_link_step<Scenario>::value) step = sc.linkStep();
        g.link(&desc, &a, &b, 1, step, &ok);
        return ok != 0;
    }
};

static std::size_t g_targetNodes = 0;

template <typename Planner>
std::pair<double, double> runUntilSolved(Planner& planner, double timeMs) {
    const auto t0 = Clock::now();
    double first = -1;
    const bool trace = std::getenv("MPTG_DEMO_TRACE") != nullptr;
    planner.solveFor(
        [&] {
            if (trace) std::fprintf(stderr, "trace: %zu nodes, solved=%d\n", planner.size(), (int)planner.solved());
            if (first < 0 && planner.solved()) first = std::chrono::duration<double>(Clock::now() - t0).count();
            return g_targetNodes ? planner.size() >= g_targetNodes : planner.solved();
        },
        std::chrono::duration<double, std::milli>(timeMs));
    if (first < 0 && planner.solved()) first = std::chrono::duration<double>(Clock::now() - t0).count();
    return {first, std::chrono::duration<double>(Clock::now() - t0).count()};
}

// ------------------------------------------------------------------ C1
void holonomic(const Options& opt) {
    using Scalar = double;
    using Scenario = demo::Holonomic2DPointScenario<Scalar>;
    using State = Scenario::State;
    const int width = 1024, height = 512;
    std::vector<shape::Circle<Scalar>> circles{{170.0, 140.0, 80.0}, {800.0, 70.0, 50.0}, {900.0, 380.0, 70.0}};
    std::vector<shape::Rect<Scalar>> rects{{375, 140, 520, 220}, {200, 320, 390, 390}, {600, 200, 680, 450}};
    State start = makeState<Scalar, 2>({30, 30}), goal = makeState<Scalar, 2>({width - 30.0, height - 30.0});
    Scenario scenario(width, height, circles, rects, goal);
    Planner<Scenario, PRRT<report_stats<true>, wave_size<512>>> planner(scenario, opt.seed);
    planner.addStart(start);
    auto [first, total] = runUntilSolved(planner, opt.timeMs);
    report("holonomic_2d_point", "PRRT", planner, scenario, first, total, opt);
}

// ------------------------------------------------------------------ C2
static std::vector<std::uint8_t> syntheticMap(int w, int h, std::uint64_t seed) {
    std::vector<std::uint8_t> occ((std::size_t)w * h, 0);
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u(0, 1);
    for (int b = 0; b < 125; ++b) {
        const double cx = u(rng) * w, cy = u(rng) * h;
        if (u(rng) < 0.5) {
            const double r = 30 + u(rng) * 170;
            for (int y = std::max(0, (int)(cy - r)); y < std::min(h, (int)(cy + r)); ++y)
                for (int x = std::max(0, (int)(cx - r)); x < std::min(w, (int)(cx + r)); ++x)
                    if ((x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r) occ[(std::size_t)y * w + x] = 1;
        } else {
            const double bw = 40 + u(rng) * 400, bh = 20 + u(rng) * 200;
            for (int y = std::max(0, (int)(cy - bh / 2)); y < std::min(h, (int)(cy + bh / 2)); ++y)
                for (int x = std::max(0, (int)(cx - bw / 2)); x < std::min(w, (int)(cx + bw / 2)); ++x) occ[(std::size_t)y * w + x] = 1;
        }
    }
    return occ;
}

static bool readPgm(const std::string& path, int& w, int& h, std::vector<std::uint8_t>& occ) {
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv;
    if (!(f >> magic >> w >> h >> maxv) || magic != "P5") return false;
    f.get();
    occ.resize((std::size_t)w * h);
    f.read((char*)occ.data(), (std::streamsize)occ.size());
    for (auto& c : occ) c = c ? 1 : 0;  // non-zero = obstacle
    return (bool)f;
}

void png2d(const Options& opt) {
    using Scalar = double;
    using Scenario = demo::PNG2dScenario<Scalar>;
    using State = Scenario::State;
    int width = 3976, height = 2603;  // size of demo/png_planning_input.png
    std::vector<std::uint8_t> occ;
    State start = makeState<Scalar, 2>({430, 1300}), goal = makeState<Scalar, 2>({3150, 950});  // png_2d_planning.cpp:84-86
    bool haveMap = false;
    if (opt.map.size() > 4 && opt.map.substr(opt.map.size() - 4) == ".png") {
        // the reference's own input: decode the PNG and apply its obstacle colour filters (png_2d_planning.cpp:69-72,
        // png_2d_scenario.hpp:192-265)
        const formats::Image img = formats::readPngRgb(opt.map);
        occ = formats::filterObstacles(img, {{126, 106, 61, 15}, {61, 53, 6, 15}, {255, 255, 255, 5}});
        width = img.width, height = img.height, haveMap = true;
    } else if (!opt.map.empty()) {
        haveMap = readPgm(opt.map, width, height, occ);
    }
    if (!haveMap) {
        occ = syntheticMap(width, height, 11);
        // keep the shipped start usable on the synthetic map, then move the goal to the reachable free
        // cell (8 px clearance) closest to the shipped goal
        for (int y = (int)start[1] - 40; y <= (int)start[1] + 40; ++y)
            for (int x = (int)start[0] - 40; x <= (int)start[0] + 40; ++x)
                if (x >= 0 && y >= 0 && x < width && y < height) occ[(std::size_t)y * width + x] = 0;
        std::vector<std::uint8_t> seen((std::size_t)width * height, 0);
        std::vector<std::uint32_t> frontier{(std::uint32_t)((int)start[1] * width + (int)start[0])};
        seen[frontier[0]] = 1;
        while (!frontier.empty()) {
            const std::uint32_t c = frontier.back();
            frontier.pop_back();
            const int x = (int)(c % width), y = (int)(c / width);
            const int nx[4] = {x + 1, x - 1, x, x}, ny[4] = {y, y, y + 1, y - 1};
            for (int i = 0; i < 4; ++i)
                if (nx[i] >= 0 && ny[i] >= 0 && nx[i] < width && ny[i] < height) {
                    const std::uint32_t n = (std::uint32_t)(ny[i] * width + nx[i]);
                    if (!seen[n] && !occ[n]) seen[n] = 1, frontier.push_back(n);
                }
        }
        double best = 1e300;
        State moved = goal;
        for (int y = 8; y < height - 8; y += 4)
            for (int x = 8; x < width - 8; x += 4) {
                if (!seen[(std::size_t)y * width + x]) continue;
                const double d = (x - goal[0]) * (x - goal[0]) + (y - goal[1]) * (y - goal[1]);
                if (d >= best) continue;
                bool clearAround = true;
                for (int yy = y - 8; yy <= y + 8 && clearAround; ++yy)
                    for (int xx = x - 8; xx <= x + 8; ++xx)
                        if (occ[(std::size_t)yy * width + xx]) {
                            clearAround = false;
                            break;
                        }
                if (clearAround) best = d, moved = makeState<Scalar, 2>({(Scalar)x, (Scalar)y});
            }
        goal = moved;
    }
    if (const char* dump = std::getenv("MPTG_DEMO_DUMP_MAP")) {
        std::ofstream f(dump, std::ios::binary);
        f << "P5\n" << width << " " << height << "\n255\n";
        for (auto c : occ) f.put(c ? (char)255 : (char)0);
    }
    Scenario scenario(width, height, goal, occ);
    Planner<Scenario, PRRTStar<report_stats<true>, wave_size<2048>>> planner(scenario, opt.seed);
    planner.addStart(start);
    planner.setRange(200);
    auto [first, total] = runUntilSolved(planner, opt.timeMs);
    report("png_2d", "PRRT*", planner, scenario, first, total, opt);
    if (opt.devicePrrt) {
        Planner<Scenario, PRRT<device_resident, report_stats<true>, wave_size<16384>, max_nodes<(1 << 22)>>> dev(scenario, opt.seed);
        dev.addStart(start);
        dev.setRange(200);
        auto [dFirst, dTotal] = runUntilSolved(dev, opt.timeMs);
        report("png_2d", "PRRT, device-resident", dev, scenario, dFirst, dTotal, opt);
        Planner<Scenario, PRRTStar<device_resident, report_stats<true>, wave_size<8192>, max_nodes<(1 << 21)>>> star(scenario, opt.seed);
        star.addStart(start);
        star.setRange(200);
        auto [sFirst, sTotal] = runUntilSolved(star, opt.timeMs);
        report("png_2d", "PRRT*, device-resident", star, scenario, sFirst, sTotal, opt);
    }
}

// ------------------------------------------------------------------ C3
static void tube(std::vector<float>& tris, double scale, double wobble, int nseg, double phase, double radius, int sides) {
    std::vector<std::array<double, 3>> p(nseg), t(nseg);
    for (int i = 0; i < nseg; ++i) {
        const double s = 1.6 * 3.14159265358979 * i / (nseg - 1);
        p[i] = {scale * std::cos(s), scale * std::sin(s) * (1 + 0.2 * std::sin(3 * s + phase)), wobble * scale * std::sin(2 * s + phase)};
    }
    auto sub = [](auto a, auto b) { return std::array<double, 3>{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; };
    auto norm = [](std::array<double, 3> v) {
        const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        return std::array<double, 3>{v[0] / n, v[1] / n, v[2] / n};
    };
    auto cross = [](auto a, auto b) { return std::array<double, 3>{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; };
    for (int i = 0; i < nseg; ++i) t[i] = norm(sub(p[std::min(i + 1, nseg - 1)], p[std::max(i - 1, 0)]));
    std::array<double, 3> nrm = norm(cross(t[0], std::array<double, 3>{0, 0, 1}));
    std::vector<std::vector<std::array<double, 3>>> rings(nseg, std::vector<std::array<double, 3>>(sides));
    for (int i = 0; i < nseg; ++i) {
        const double d = nrm[0] * t[i][0] + nrm[1] * t[i][1] + nrm[2] * t[i][2];
        nrm = norm({nrm[0] - d * t[i][0], nrm[1] - d * t[i][1], nrm[2] - d * t[i][2]});
        const auto b = cross(t[i], nrm);
        for (int j = 0; j < sides; ++j) {
            const double a = 2 * 3.14159265358979 * j / sides;
            for (int c = 0; c < 3; ++c) rings[i][j][c] = p[i][c] + radius * (std::cos(a) * nrm[c] + std::sin(a) * b[c]);
        }
    }
    auto push = [&](const std::array<double, 3>& a, const std::array<double, 3>& b, const std::array<double, 3>& c) {
        for (auto* v : {&a, &b, &c})
            for (int k = 0; k < 3; ++k) tris.push_back((float)(*v)[k]);
    };
    for (int i = 0; i + 1 < nseg; ++i)
        for (int j = 0; j < sides; ++j) {
            push(rings[i][j], rings[i][(j + 1) % sides], rings[i + 1][j]);
            push(rings[i][(j + 1) % sides], rings[i + 1][(j + 1) % sides], rings[i + 1][j]);
        }
    for (int j = 0; j < sides; ++j) {
        push(p[0], rings[0][(j + 1) % sides], rings[0][j]);
        push(p[nseg - 1], rings[nseg - 1][j], rings[nseg - 1][(j + 1) % sides]);
    }
}

void se3RigidBody(const Options& opt) {
    using Scalar = float;
    using Scenario = demo::SE3RigidBodyScenario<Scalar>;
    using State = Scenario::State;
    std::vector<float> env, robot;
    State start = State::identityAt(-52, -50, 0), goal = State::identityAt(52, 50, 5);
    mptg::State<Scalar, 3> vmin = makeState<Scalar, 3>({-60, -60, -40}), vmax = makeState<Scalar, 3>({60, 60, 40});
    Scalar range = 40;
    if (!opt.cfg.empty()) {  // the reference's own inputs: [problem] robot / world (COLLADA through assimp there), start, goal, volume
        const formats::ScenarioConfig cfg(opt.cfg);
        const std::size_t slash = opt.cfg.find_last_of("\\/");
        const std::string dir = slash == std::string::npos ? "" : opt.cfg.substr(0, slash + 1);
        std::string world, robotName;
        cfg.load(world, "problem", "world");
        cfg.load(robotName, "problem", "robot");
        env = formats::readMeshTriangles(dir + world);
        robot = formats::readMeshTriangles(dir + robotName);  // recentred by the scenario below (se3_rigid_body_scenario.hpp:181-193)
        cfg.loadSE3(start.data(), "problem", "start");
        cfg.loadSE3(goal.data(), "problem", "goal");
        cfg.loadVector3(vmin.data(), "problem", "volume.min");
        cfg.loadVector3(vmax.data(), "problem", "volume.max");
        if (cfg.hasProp("planner", "rrt.range")) cfg.load(range, "planner", "rrt.range");
        std::printf("se3_rigid_body: %s -- world %s (%zu triangles), robot %s (%zu triangles)\n", opt.cfg.c_str(), world.c_str(), env.size() / 9,
                    robotName.c_str(), robot.size() / 9);
    } else {
        tube(env, 30.0, 0.35, 125, 0.3, 4.0, 16);   // ~4k triangles
        tube(robot, 14.0, 0.45, 50, 1.7, 2.0, 10);  // ~1k triangles
    }
    Scenario scenario(env, robot, goal, vmin, vmax, 0.01f);
    {
        Probe<Scenario> probe(scenario);
        if (!probe.valid(start) || !probe.valid(goal)) throw std::runtime_error("se3_rigid_body: start or goal in collision");
        std::printf("se3_rigid_body: direct start->goal edge %s\n", probe.link(start, goal) ? "valid" : "blocked");
    }
    Planner<Scenario, PRRTStar<report_stats<true>, wave_size<1024>>> planner(scenario, opt.seed);
    planner.addStart(start);
    planner.setRange(range);
    auto [first, total] = runUntilSolved(planner, opt.timeMs);
    report("se3_rigid_body", "PRRT*", planner, scenario, first, total, opt);
    if (opt.devicePrrt) {
        Planner<Scenario, PRRT<device_resident, report_stats<true>, wave_size<16384>, max_nodes<(1 << 22)>>> dev(scenario, opt.seed);
        dev.addStart(start);
        dev.setRange(range);
        auto [dFirst, dTotal] = runUntilSolved(dev, opt.timeMs);
        report("se3_rigid_body", "PRRT, device-resident", dev, scenario, dFirst, dTotal, opt);
        Planner<Scenario, PRRTStar<device_resident, report_stats<true>, wave_size<8192>, max_nodes<(1 << 21)>>> star(scenario, opt.seed);
        star.addStart(start);
        star.setRange(range);
        auto [sFirst, sTotal] = runUntilSolved(star, opt.timeMs);
        report("se3_rigid_body", "PRRT*, device-resident", star, scenario, sFirst, sTotal, opt);
    }
}

// ------------------------------------------------------------------ C4
template <int N>
void linkManipulator(const Options& opt) {
    using Scalar = double;
    using Scenario = demo::LinkManipulatorScenario<Scalar, N>;
    using State = typename Scenario::State;
    std::vector<Scalar> lengths(N, 4.0);
    const Scalar reach = 4.0 * N;
    std::vector<shape::Circle<Scalar>> circles;
    std::mt19937_64 rng(5);
    std::uniform_real_distribution<double> u(0, 1);
    while (circles.size() < 8) {
        const double r = reach * (0.35 + 0.6 * u(rng)), a = (0.15 + 0.7 * u(rng)) * 2 * 3.14159265358979;
        circles.push_back({r * std::cos(a), r * std::sin(a), 3.0});
    }
    // straight arm along +x (the angular gap above keeps it free) to the first swung-round pose that is
    // free but not directly connectable, so the roadmap has to go around the circles
    State start = State::Zero(), goal = State::Zero();
    {
        Scenario trial(goal, circles, lengths, 0.5);
        Probe<Scenario> probe(trial);
        if (!probe.valid(start)) throw std::runtime_error("link_manipulator: start in collision");
        bool found = false;
        for (double base = 3.0; base > 0.5 && !found; base -= 0.1) {
            State cand = State::Zero();
            cand[0] = base;
            for (int i = 1; i < N; ++i) cand[i] = (i % 2 ? 0.05 : -0.05);
            if (probe.valid(cand) && !probe.link(start, cand)) goal = cand, found = true;
        }
        if (!found) throw std::runtime_error("link_manipulator: no blocked goal pose found");
    }
    Scenario scenario(goal, circles, lengths, 0.5);
    Planner<Scenario, PPRM<report_stats<true>, wave_size<1024>>> planner(scenario, opt.seed);
    planner.addStart(start);
    planner.addGoal(goal);
    auto [first, total] = runUntilSolved(planner, opt.timeMs);
    char name[64];
    std::snprintf(name, sizeof name, "link_manipulator N=%d", N);
    report(name, "PPRM", planner, scenario, first, total, opt);
    if (opt.devicePrrt) {  // the same roadmap planner with the roadmap kept on the GPU (mptg_pprm_*)
        Planner<Scenario, PPRM<device_resident, report_stats<true>, wave_size<4096>, max_nodes<(1 << 19)>>> dev(scenario, opt.seed);
        dev.addStart(start);
        dev.addGoal(goal);
        auto [dFirst, dTotal] = runUntilSolved(dev, opt.timeMs);
        report(name, "PPRM, device-resident", dev, scenario, dFirst, dTotal, opt);
    }
}

// ------------------------------------------------------------------ the fifth demo of the reference
// demo/nao_cup_planning.cpp:155-215: Planner<NaoCupScenario<S>, Algorithm>, start = nao_init_config, goal = the target
// configuration within 1e-5.  Not part of --all: the straight edge is blocked and a twentieth of the joint box is clear,
// so a solution takes the reference seconds to minutes.
template <typename Scalar>
void naoCup(const Options& opt) {
    using Scenario = demo::NaoCupScenario<Scalar>;
    Scenario scenario;
    const char* name = sizeof(Scalar) == 4 ? "nao_cup float" : "nao_cup double";
    {
        Planner<Scenario, PRRTStar<report_stats<true>, wave_size<1024>>> planner(scenario, opt.seed);
        planner.addStart(scenario.start());
        planner.setRange(1.0);
        auto [first, total] = runUntilSolved(planner, opt.timeMs);
        report(name, "PRRT*", planner, scenario, first, total, opt);
    }
    if (opt.devicePrrt) {
        Planner<Scenario, PRRT<device_resident, report_stats<true>, wave_size<8192>, max_nodes<(1 << 21)>>> dev(scenario, opt.seed);
        dev.addStart(scenario.start());
        dev.setRange(1.0);
        auto [first, total] = runUntilSolved(dev, opt.timeMs);
        report(name, "PRRT, device-resident", dev, scenario, first, total, opt);
    }
}

static void onCrash(int sig) {
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "planning_demos: fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof msg - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char** argv) {
    signal(SIGSEGV, onCrash);
    signal(SIGABRT, onCrash);
    std::setvbuf(stdout, nullptr, _IOLBF, 0);
    Options opt;
    std::string which = "all";
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--all") which = "all";
        else if (a == "--demo" && i + 1 < argc) which = argv[++i];
        else if (a == "--time-ms" && i + 1 < argc) opt.timeMs = std::atof(argv[++i]);
        else if (a == "--check") opt.check = true;
        else if (a == "--device-prrt") opt.devicePrrt = true;
        else if (a == "--nodes" && i + 1 < argc) opt.nodes = g_targetNodes = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--map" && i + 1 < argc) opt.map = argv[++i];
        else if (a == "--cfg" && i + 1 < argc) opt.cfg = argv[++i];
        else if (a == "--seed" && i + 1 < argc) opt.seed = std::strtoull(argv[++i], nullptr, 10);
        else {
            std::fprintf(stderr, "usage: %s [--all | --demo holonomic_2d_point|png_2d|se3_rigid_body|link_manipulator|nao_cup|nao_cup_float] [--time-ms T] [--nodes N] [--check] [--device-prrt (also run the device-resident PRRT / PPRM)] [--map file.png|file.pgm] [--cfg problem.cfg (se3_rigid_body: .dae / .obj meshes, start, goal, volume)] [--seed S]\n", argv[0]);
            return 2;
        }
    }
    try {
        if (which == "all" || which == "holonomic_2d_point") holonomic(opt);
        if (which == "all" || which == "png_2d") png2d(opt);
        if (which == "all" || which == "se3_rigid_body") se3RigidBody(opt);
        if (which == "all" || which == "link_manipulator") {
            linkManipulator<8>(opt);
            linkManipulator<16>(opt);
            linkManipulator<32>(opt);
        }
        if (which == "nao_cup") naoCup<double>(opt);
        if (which == "nao_cup_float") naoCup<float>(opt);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return failures ? 1 : 0;
}
