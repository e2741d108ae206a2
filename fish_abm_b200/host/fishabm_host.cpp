#include "fishabm_host.hpp"

namespace fishabm_host {

Backend::Backend(int n, const Structure& s, uint64_t seed, int device, int replicates, bool rolling, double speed_norm, uint32_t replicate0)
    : n_(n), reps_(replicates), s_(s) {
    fishabm_params p;
    fishabm_params_default(&p);
    p.n = n; p.w = s.w; p.h = s.h; p.seed = seed; p.device = device; p.replicates = replicates; p.rolling = rolling ? 1 : 0;
    p.speed_norm = speed_norm; p.replicate0 = replicate0;
    const int rc = fishabm_create(&p, &ctx_);
    if (rc != FISHABM_OK) throw Error(rc, std::string("fishabm_create: ") + fishabm_last_error(nullptr));
}

Backend::~Backend() { fishabm_destroy(ctx_); }

void Backend::ck(int rc) const {
    if (rc != FISHABM_OK) throw Error(rc, fishabm_last_error(ctx_));
}

void Backend::push(const individuals& inds) {
    if (inds.n != n_ * reps_) throw Error(FISHABM_E_ARG, "individuals.n does not match the backend");
    std::vector<double> speed((size_t)inds.n, inds.speed);  // the reference has one speed per school (fish_mod.cpp:43)
    ck(fishabm_set_state(ctx_, inds.coords.data(), inds.dir.data(), speed.data(), inds.speed_factor.data(), inds.lag.data(),
                         inds.sample_rate.data(), inds.food_intake.data()));
}

void Backend::pull(individuals& inds) {
    if (inds.n != n_ * reps_) inds.resize(n_ * reps_);
    ck(fishabm_get_state(ctx_, inds.coords.data(), inds.dir.data(), nullptr, inds.speed_factor.data(), inds.lag.data(),
                         inds.sample_rate.data(), inds.food_intake.data()));
}

void Backend::set_quality(const std::vector<int>& q, int rows, int cols) { ck(fishabm_set_quality(ctx_, q.data(), rows, cols, cols)); }

void Backend::get_quality(std::vector<int>& q, int replicate) {
    q.assign((size_t)s_.rows * s_.cols, 0);
    ck(fishabm_get_quality(ctx_, replicate, q.data(), s_.rows, s_.cols, s_.cols));
}

void initialize(Backend& b, individuals& inds, int n, int dim, double speed, const Structure&) {
    if (n != b.n()) throw Error(FISHABM_E_ARG, "initialize: n does not match the backend");
    b.ck(fishabm_initialize(b.ctx(), speed));
    inds.resize(n * b.replicates());
    inds.dim = dim; inds.speed = speed;
    b.pull(inds);
}

void get_environment(Backend& b, std::vector<int>& quality, const Structure& struc, const std::string& csv_path) {
    quality.assign((size_t)struc.rows * struc.cols, 0);
    int rows = 0, cols = 0;
    if (fishabm_load_quality_csv(csv_path.c_str(), quality.data(), struc.rows, struc.cols, struc.cols, &rows, &cols) != FISHABM_OK)
        throw Error(FISHABM_E_ARG, "get_environment: cannot read " + csv_path + " as a " + std::to_string(struc.rows) + "x" + std::to_string(struc.cols) + " grid");
    if (rows != struc.rows || cols != struc.cols) {
        // the reference tiles nothing: a smaller file leaves the rest of the grid empty; larger worlds tile the file
        std::vector<int> t((size_t)struc.rows * struc.cols);
        for (int r = 0; r < struc.rows; ++r)
            for (int c = 0; c < struc.cols; ++c) t[(size_t)r * struc.cols + c] = quality[(size_t)(r % rows) * struc.cols + (c % cols)];
        quality.swap(t);
    }
    b.set_quality(quality, struc.rows, struc.cols);
}

void move(Backend& b) { b.ck(fishabm_move(b.ctx())); }
void sample(Backend& b) { b.ck(fishabm_sample(b.ctx())); }
void step(Backend& b, int nsteps) { b.ck(fishabm_step(b.ctx(), nsteps)); }
void feed(Backend& b, int id) { int32_t i = id; b.ck(fishabm_feed(b.ctx(), &i, 1)); }

std::vector<int> in_zone(Backend& b, int fov, double r) {
    std::vector<int> out((size_t)b.n() * b.replicates());
    b.ck(fishabm_in_zone(b.ctx(), fov, r, out.data()));
    return out;
}
std::vector<double> social(Backend& b) {
    std::vector<double> out((size_t)b.n() * b.replicates());
    b.ck(fishabm_social(b.ctx(), out.data(), nullptr, nullptr, nullptr));
    return out;
}
std::vector<double> get_speed(Backend& b) {
    std::vector<double> out((size_t)b.n() * b.replicates());
    b.ck(fishabm_get_speed(b.ctx(), out.data()));
    return out;
}
long long sum_array(Backend& b) {
    std::vector<int64_t> s((size_t)b.replicates());
    b.ck(fishabm_sum_quality(b.ctx(), s.data()));
    long long t = 0;
    for (auto v : s) t += v;
    return t;
}
double averageturn(const std::vector<double>& v, bool comb_weight) { return fishabm_averageturn(v.data(), (int32_t)v.size(), comb_weight ? 1 : 0); }
double correctangle(double dir) { return fishabm_correctangle(dir); }
void correct_coords(double* xy, const Structure& struc) { fishabm_correct_coords(xy, struc.w, struc.h); }

}  // namespace fishabm_host
