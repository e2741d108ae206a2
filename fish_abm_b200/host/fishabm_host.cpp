#include "fishabm_host.hpp"

#include <algorithm>

namespace fishabm_host {

Backend::Backend(int n, const Structure& s, uint64_t seed, int device, int replicates, bool rolling, double speed_norm, uint32_t replicate0,
                 const std::vector<int>& slab_devices)
    : n_(n), reps_(replicates), s_(s) {
    const int nslabs = (int)slab_devices.size();
    if (nslabs > 1 && replicates != 1) throw Error(FISHABM_E_ARG, "a split school cannot be an ensemble");
    for (int r = 0; r < std::max(1, nslabs); ++r) {
        fishabm_params p;
        fishabm_params_default(&p);
        p.n = n; p.w = s.w; p.h = s.h; p.seed = seed; p.replicates = replicates; p.rolling = rolling ? 1 : 0;
        p.speed_norm = speed_norm; p.replicate0 = replicate0;
        p.device = nslabs > 1 ? slab_devices[r] : device;
        if (nslabs > 1) { p.rank = r; p.nranks = nslabs; }
        fishabm_ctx* c = nullptr;
        const int rc = fishabm_create(&p, &c);
        if (rc != FISHABM_OK) {
            const std::string msg = std::string("fishabm_create: ") + fishabm_last_error(nullptr);
            for (fishabm_ctx* o : ctxs_) fishabm_destroy(o);
            throw Error(rc, msg);
        }
        ctxs_.push_back(c);
    }
    // peer-memory transport only when every slab has its own GPU (kernels of one GPU cannot wait for each other); slabs that
    // share a GPU are driven through the caller-driven exchange (step())
    attached_ = nslabs > 1;
    for (int r = 0; r < nslabs; ++r)
        for (int q = r + 1; q < nslabs; ++q) if (slab_devices[r] == slab_devices[q]) attached_ = false;
    if (attached_) {  // every slab learns where its two neighbours' inboxes are
        std::vector<fishabm_slab_endpoint> ep(nslabs);
        for (int r = 0; r < nslabs; ++r) ck(fishabm_slab_get_endpoint(ctxs_[r], &ep[r]), ctxs_[r]);
        for (int r = 0; r < nslabs; ++r) ck(fishabm_slab_attach(ctxs_[r], &ep[(r + nslabs - 1) % nslabs], &ep[(r + 1) % nslabs]), ctxs_[r]);
    }
}

Backend::~Backend() { for (fishabm_ctx* c : ctxs_) fishabm_destroy(c); }

void Backend::ck(int rc, fishabm_ctx* c) const {
    if (rc != FISHABM_OK) throw Error(rc, fishabm_last_error(c ? c : ctxs_[0]));
}

void Backend::step(int nsteps) {
    if (ctxs_.size() == 1) { ck(fishabm_step(ctxs_[0], nsteps)); return; }
    if (!attached_) {  // caller-driven exchange: what leaves through a slab's left edge arrives at its left neighbour's right side
        const int k = (int)ctxs_.size();
        std::vector<void*> ol(k), orr(k), il(k), ir(k);
        int64_t bytes = 0;
        for (int z = 0; z < nsteps; ++z) {
            for (fishabm_ctx* c : ctxs_) ck(fishabm_slab_begin_step(c), c);
            for (int r = 0; r < k; ++r) ck(fishabm_slab_buffers(ctxs_[r], &ol[r], &orr[r], &il[r], &ir[r], &bytes), ctxs_[r]);
            for (int r = 0; r < k; ++r) {
                ck(fishabm_slab_copy(ir[(r + k - 1) % k], ol[r], bytes), ctxs_[r]);
                ck(fishabm_slab_copy(il[(r + 1) % k], orr[r], bytes), ctxs_[r]);
            }
            for (fishabm_ctx* c : ctxs_) ck(fishabm_slab_end_step(c), c);
        }
        return;
    }
    for (int done = 0; done < nsteps;) {  // chunks short enough for the launch queues: nobody can finish before all are enqueued
        const int k = std::min(16, nsteps - done);
        for (fishabm_ctx* c : ctxs_) ck(fishabm_step_enqueue(c, k), c);
        for (fishabm_ctx* c : ctxs_) ck(fishabm_wait(c), c);
        done += k;
    }
}

long long Backend::sum_quality() {
    long long t = 0;
    std::vector<int64_t> s((size_t)reps_);
    for (fishabm_ctx* c : ctxs_) {
        ck(fishabm_sum_quality(c, s.data()), c);
        for (auto v : s) t += v;
    }
    return t;
}

void Backend::push(const individuals& inds) {
    if (inds.n != n_ * reps_) throw Error(FISHABM_E_ARG, "individuals.n does not match the backend");
    each([&](fishabm_ctx* c) {
        return fishabm_set_state(c, inds.coords.data(), inds.dir.data(), inds.speed.data(), inds.speed_factor.data(), inds.lag.data(),
                                 inds.sample_rate.data(), inds.food_intake.data());
    });
}

void Backend::pull(individuals& inds) {
    if (inds.n != n_ * reps_) inds.resize(n_ * reps_);
    each([&](fishabm_ctx* c) {
        return fishabm_get_state(c, inds.coords.data(), inds.dir.data(), inds.speed.data(), inds.speed_factor.data(), inds.lag.data(),
                                 inds.sample_rate.data(), inds.food_intake.data());
    });
}

void Backend::set_quality(const std::vector<int>& q, int rows, int cols) {
    each([&](fishabm_ctx* c) { return fishabm_set_quality(c, q.data(), rows, cols, cols); });
}

void Backend::get_quality(std::vector<int>& q, int replicate) {
    q.assign((size_t)s_.rows * s_.cols, 0);
    each([&](fishabm_ctx* c) { return fishabm_get_quality(c, replicate, q.data(), s_.rows, s_.cols, s_.cols); });
}

void initialize(Backend& b, individuals& inds, int n, int dim, double speed, const Structure&) {
    if (n != b.n()) throw Error(FISHABM_E_ARG, "initialize: n does not match the backend");
    b.each([&](fishabm_ctx* c) { return fishabm_initialize(c, speed); });  // every slab composes the same school and keeps its columns
    inds.resize(n * b.replicates());
    inds.dim = dim;
    b.pull(inds);  // speed[i] = speed for every fish, as initialize() sets it (fish_mod.cpp:190)
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

static void whole_school_only(Backend& b, const char* what) {
    if (b.slabs() > 1) throw Error(FISHABM_E_STATE, std::string(what) + " separately is not available for a split school: use step()");
}
void move(Backend& b) { whole_school_only(b, "move()"); b.ck(fishabm_move(b.ctx())); }
void sample(Backend& b) { whole_school_only(b, "sample()"); b.ck(fishabm_sample(b.ctx())); }
void step(Backend& b, int nsteps) { b.step(nsteps); }
void feed(Backend& b, int id) { int32_t i = id; b.each([&](fishabm_ctx* c) { return fishabm_feed(c, &i, 1); }); }

std::vector<int> in_zone(Backend& b, int fov, double r) {
    std::vector<int> out((size_t)b.n() * b.replicates());
    b.each([&](fishabm_ctx* c) { return fishabm_in_zone(c, fov, r, out.data()); });
    return out;
}
std::vector<double> social(Backend& b) {
    std::vector<double> out((size_t)b.n() * b.replicates());
    b.each([&](fishabm_ctx* c) { return fishabm_social(c, out.data(), nullptr, nullptr, nullptr); });
    return out;
}
std::vector<double> get_speed(Backend& b) {
    std::vector<double> out((size_t)b.n() * b.replicates());
    b.each([&](fishabm_ctx* c) { return fishabm_get_speed(c, out.data()); });
    return out;
}
long long sum_array(Backend& b) { return b.sum_quality(); }
double averageturn(const std::vector<double>& v, bool comb_weight) { return fishabm_averageturn(v.data(), (int32_t)v.size(), comb_weight ? 1 : 0); }
double correctangle(double dir) { return fishabm_correctangle(dir); }
void correct_coords(double* xy, const Structure& struc) { fishabm_correct_coords(xy, struc.w, struc.h); }

}  // namespace fishabm_host
