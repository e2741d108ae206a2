// fish_abm_headless: the reference's main() (fish_mod.cpp:84-161) without the viewer.
//
//   initialize -> get_environment -> for z in steps { move; sample }   with the reference's two commented-out loggers
//   (fish_mod.cpp:97-100,123-124,132-141,149-154) switched on by --log-quality / --log-intake:
//     quality log : one row per replicate: sum_array before the run, then after every step   (the shape of data/*_speed/*.csv)
//     intake log  : one row per 10 steps, one column per fish: cumulative food_intake        (the shape of data/*_int_uptake_speed5/*.csv)
// Everything per-step runs on the GPU through libfishabm.so; --replicates R runs R independent schools as one
// ensemble launch.  OpenCV is not needed; a viewer can read the state through fishabm_get_state every k steps.
#include <chrono>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "fishabm_host.hpp"

using namespace fishabm_host;

static void usage() {
    std::puts(
        "usage: fish_abm_headless [--n 60] [--steps 500] [--speed 5] [--seed S] [--w 2000] [--h 2000] [--quality quality.csv]\n"
        "                         [--replicates 1] [--no-rolling] [--device 0] [--log-quality out.csv] [--log-intake out.csv]\n"
        "                         [--intake-every 10] [--unfused] [--quiet] [--slabs K] [--devices a,b,..]\n"
        "--slabs K splits the school into K x-slabs (one context each, on --devices or all on --device), exchanging halos and\n"
        "migrants through peer memory; results are identical to the unsplit run\n"
        "defaults are the reference's: 60 fish, 500 steps, speed 5, 2000x2000 window (fish_mod.cpp:23,116,126,53-59)");
}

int main(int argc, char** argv) {
    int n = 60, steps = 500, w = 2000, h = 2000, replicates = 1, device = 0, intake_every = 10;
    double speed = 5.0;
    uint64_t seed = (uint64_t)time(nullptr);  // srand(time(0)), fish_mod.cpp:102
    bool rolling = true, unfused = false, quiet = false;
    std::string qpath = "quality.csv", log_q, log_i, devlist;
    int slabs = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "--n") n = std::atoi(next());
        else if (a == "--steps") steps = std::atoi(next());
        else if (a == "--speed") speed = std::atof(next());
        else if (a == "--seed") seed = std::strtoull(next(), nullptr, 0);
        else if (a == "--w") w = std::atoi(next());
        else if (a == "--h") h = std::atoi(next());
        else if (a == "--quality") qpath = next();
        else if (a == "--replicates") replicates = std::atoi(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--no-rolling") rolling = false;
        else if (a == "--log-quality") log_q = next();
        else if (a == "--log-intake") log_i = next();
        else if (a == "--intake-every") intake_every = std::atoi(next());
        else if (a == "--unfused") unfused = true;
        else if (a == "--quiet") quiet = true;
        else if (a == "--slabs") slabs = std::atoi(next());
        else if (a == "--devices") devlist = next();
        else if (a == "--help" || a == "-h") { usage(); return 0; }
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); return 2; }
    }
    try {
        Structure fish_struc(w, h);                                    // fish_mod.cpp:104
        std::vector<int> slab_devices;
        if (slabs > 1) {
            std::vector<int> given;
            for (size_t p0 = 0; p0 < devlist.size();) {
                const size_t p1 = devlist.find(',', p0);
                given.push_back(std::atoi(devlist.substr(p0, p1 - p0).c_str()));
                if (p1 == std::string::npos) break;
                p0 = p1 + 1;
            }
            for (int r = 0; r < slabs; ++r) slab_devices.push_back(given.empty() ? device : given[r % given.size()]);
        }
        Backend gpu(n, fish_struc, seed, device, replicates, rolling, /*speed_norm = num*/ (double)n, 0, slab_devices);
        individuals fish;
        initialize(gpu, fish, n, 2, speed, fish_struc);                // :116
        std::vector<int> quality;
        get_environment(gpu, quality, fish_struc, qpath);              // :121
        const long long q0 = sum_array(gpu);
        if (!quiet) std::printf("fish_abm_headless: %d school(s) x %d fish, %d steps, speed %g, sum(quality) %lld\n", replicates, n, steps, speed, q0 / replicates);

        const auto t0 = std::chrono::steady_clock::now();
        std::vector<int64_t> sumq;
        std::vector<int32_t> intake;
        const bool logging = !log_q.empty() || !log_i.empty();
        const int rows = !log_i.empty() && intake_every > 0 ? (steps + intake_every - 1) / intake_every : 0;
        if (unfused && gpu.slabs() > 1) throw Error(FISHABM_E_ARG, "--unfused is for an unsplit school");
        if (unfused) {                                                 // the reference's literal loop, :126-131
            if (logging) sumq.assign((size_t)steps * replicates, 0);
            if (rows) intake.assign((size_t)rows * replicates * n, 0);
            for (int z = 0; z < steps; ++z) {
                if (!quiet && z % 100 == 0) std::printf("step %d\n", z);   // :127-129
                move(gpu);
                sample(gpu);
                if (!log_q.empty()) gpu.ck(fishabm_sum_quality(gpu.ctx(), sumq.data() + (size_t)z * replicates));
                if (rows && z % intake_every == 0)
                    gpu.ck(fishabm_get_state(gpu.ctx(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, intake.data() + (size_t)(z / intake_every) * replicates * n));
            }
        } else if (logging && gpu.slabs() == 1) {
            if (!log_q.empty()) sumq.assign((size_t)steps * replicates, 0);
            if (rows) intake.assign((size_t)rows * replicates * n, 0);
            gpu.ck(fishabm_run_logged(gpu.ctx(), steps, sumq.empty() ? nullptr : sumq.data(), intake.empty() ? nullptr : intake.data(), intake_every));
        } else if (logging) {  // split school: the loggers read the slabs' partial sums / owned fish after every step
            if (!log_q.empty()) sumq.assign((size_t)steps, 0);
            if (rows) intake.assign((size_t)rows * n, 0);
            for (int z = 0; z < steps; ++z) {
                step(gpu, 1);
                if (!log_q.empty()) sumq[z] = sum_array(gpu);
                if (rows && z % intake_every == 0)
                    gpu.each([&](fishabm_ctx* c) { return fishabm_get_state(c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, intake.data() + (size_t)(z / intake_every) * n); });
            }
        } else {
            step(gpu, steps);
            gpu.each([&](fishabm_ctx* c) { return fishabm_flush(c); });
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        if (!log_q.empty()) {  // one row per replicate, steps+1 columns (data/rolling_speed/*.csv)
            std::ofstream f(log_q);
            for (int r = 0; r < replicates; ++r) {
                f << q0 / replicates;
                for (int z = 0; z < steps; ++z) f << ',' << sumq[(size_t)z * replicates + r];
                f << '\n';
            }
        }
        if (!log_i.empty()) {  // one row per logged step, one column per fish (replicates side by side)
            std::ofstream f(log_i);
            for (int k = 0; k < rows; ++k) {
                for (int c = 0; c < replicates * n; ++c) f << (c ? "," : "") << intake[(size_t)k * replicates * n + c];
                f << '\n';
            }
        }
        gpu.pull(fish);
        long long eaten = 0;
        for (int v : fish.food_intake) eaten += v;
        const long long q1 = sum_array(gpu);
        if (!quiet)
            std::printf("done: %.3f s, %.3e agent-steps/s; sum(quality) %lld -> %lld per school (mean), food_intake %lld, conserved: %s\n", secs,
                        (double)n * replicates * steps / secs, q0 / replicates, q1 / replicates, eaten, (q0 - q1 == eaten) ? "yes" : "NO");
        return (q0 - q1 == eaten) ? 0 : 3;
    } catch (const Error& e) {
        std::fprintf(stderr, "fish_abm_headless: error %d: %s\n", e.code, e.what());
        return 1;
    }
}
