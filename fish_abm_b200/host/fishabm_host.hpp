// Host-side model of the reference (fish_mod.cpp:35-82) over the C ABI of libfishabm.so.
//
// `individuals` and `Structure` keep the reference's member names (runtime-sized instead of `#define num`), and the
// free functions keep its prototypes' names and argument meaning (fish_mod.cpp:64-82): initialize, get_environment,
// move, sample, feed, social, in_zone, get_speed, averageturn, correctangle, correct_coords, sum_array.  The fish
// live on the GPU between calls; the host copy is refreshed on demand (pull) and uploaded when the caller edits it
// (push).  There is no CPU implementation behind these functions.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fishabm.h"

namespace fishabm_host {

struct Structure {  // class Structure, fish_mod.cpp:51-62
    int latitude = 100, longitude = 100, resolution = 1, res_factor = 20;
    int rows = 100, cols = 100, h = 2000, w = 2000;
    Structure() = default;
    Structure(int w_, int h_) : latitude(h_ / 20), longitude(w_ / 20), rows(h_ / 20), cols(w_ / 20), h(h_), w(w_) {}
};

struct individuals {  // class individuals, fish_mod.cpp:35-49 (colours / nfollow / sample[] are viewer-only or unused)
    int n = 0, dim = 2;
    std::vector<int> lag, sample_rate, food_intake;
    std::vector<double> speed;   // double speed[num], fish_mod.cpp:43: one movement speed PER FISH (used per fish at :171-172)
    std::vector<double> dir, speed_factor;
    std::vector<double> coords;  // [n][2], x,y interleaved as in the reference
    void resize(int n_) {
        n = n_;
        lag.assign(n, 0); sample_rate.assign(n, 0); food_intake.assign(n, 0);
        speed.assign(n, 0.0); dir.assign(n, 0.0); speed_factor.assign(n, 1.0); coords.assign(2 * (size_t)n, 0.0);
    }
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// One school (or `replicates` independent schools) resident on one GPU -- or one school split into x-slabs, one
// context per slab, on the GPUs listed in `slab_devices` (all in this process: the slabs are attached to each other's
// inboxes directly, the peer-memory transport of include/fishabm.h).
class Backend {
  public:
    Backend(int n, const Structure& s, uint64_t seed, int device = 0, int replicates = 1, bool rolling = true, double speed_norm = 0.0,
            uint32_t replicate0 = 0, const std::vector<int>& slab_devices = {});
    ~Backend();
    Backend(const Backend&) = delete;
    Backend& operator=(const Backend&) = delete;

    int n() const { return n_; }
    int replicates() const { return reps_; }
    int slabs() const { return (int)ctxs_.size(); }
    fishabm_ctx* ctx(int k = 0) { return ctxs_[k]; }
    // apply f to every context; for a split school the per-fish outputs of the slabs fill disjoint entries of the same
    // whole-school arrays, so most calls are simply repeated on every slab
    template <class F> void each(F f) { for (fishabm_ctx* c : ctxs_) ck(f(c), c); }
    void step(int nsteps);                               // all slabs: enqueue, then wait (they wait for each other on the device)
    long long sum_quality();

    void push(const individuals& inds);                  // host -> device (one school)
    void pull(individuals& inds);                        // device -> host
    void set_quality(const std::vector<int>& q, int rows, int cols);
    void get_quality(std::vector<int>& q, int replicate = 0);
    void ck(int rc, fishabm_ctx* c = nullptr) const;

  private:
    std::vector<fishabm_ctx*> ctxs_;
    bool attached_ = false;   // slabs attached to each other's inboxes (one GPU per slab); else the exchange is driven from step()
    int n_, reps_;
    Structure s_;
};

// ---- the reference's operator API (fish_mod.cpp:64-82) ---------------------------------------------------------------
void initialize(Backend& b, individuals& inds, int n, int dim, double speed, const Structure& struc);   // :183-206
void get_environment(Backend& b, std::vector<int>& quality, const Structure& struc, const std::string& csv_path);  // :603-629
void move(Backend& b);                                                                                  // :163-181
void sample(Backend& b);                                                                                // :540-558
void step(Backend& b, int nsteps);                                                                      // :126-131
void feed(Backend& b, int id);                                                                          // :560-586
std::vector<int> in_zone(Backend& b, int fov, double r);                                                // :226-260, every id
std::vector<double> social(Backend& b);                                                                 // :262-377, every id
std::vector<double> get_speed(Backend& b);                                                              // :588-591, every id
long long sum_array(Backend& b);                                                                        // :593-601
double averageturn(const std::vector<double>& v, bool comb_weight);                                     // :379-408
double correctangle(double dir);                                                                        // :410-418
void correct_coords(double* xy, const Structure& struc);                                                // :466-480

}  // namespace fishabm_host
