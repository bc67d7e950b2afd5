// integration/rcpp_shim_util.h -- helpers shared by the reference-side binding files rcpp_shim.cpp and
// rcpp_shim_legacy.cpp: error forwarding, RNG bookkeeping, the cache of lna_problem handles.
// Everything sits in an anonymous namespace: each translation unit keeps its own handle cache.
#pragma once
#include <RcppArmadillo.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "lnaphylodyn_b200.h"

using namespace Rcpp;
using namespace arma;

namespace {

const char *kCholMsg = "Invalid input for cholesky decomposition., parametrization fails";  // LNA_functional.cpp:663

[[noreturn]] void fail_abi(const char *what) { Rcpp::stop(std::string(what) + ": " + lna_last_error()); }
#define LNA_CALL(expr, what)              \
    do {                                  \
        if ((expr) != 0) fail_abi(what);  \
    } while (0)

// ---- RNG bookkeeping ---------------------------------------------------------------------------------------------
// mark() after the draws the reference always makes; settle(n) leaves the generator as if exactly n more uniforms had
// been drawn after the mark.
#ifdef LNA_STANDIN_RCPP
struct RngMark {
    long at = 0;
    void mark() { at = shim_rng().used_unif; }
    void settle(int n) { shim_rng().used_unif = at + n; }
};
#else
struct RngMark {  // real R: snapshot .Random.seed, restore it, then advance by n uniforms
    Rcpp::RObject seed;
    void mark() {
        PutRNGstate();
        seed = Rcpp::clone(Rcpp::RObject(Rcpp::Environment::global_env()[".Random.seed"]));
        GetRNGstate();
    }
    void settle(int n) {
        PutRNGstate();
        Rcpp::Environment::global_env()[".Random.seed"] = seed;
        GetRNGstate();
        for (int i = 0; i < n; ++i) unif_rand();
    }
};
#endif

// ---- one cached lna_problem per (settings, genealogy) ----------------------------------------------------------------
void key_add(std::string &k, const void *p, size_t n) { k.append(static_cast<const char *>(p), n); }

struct InitView {  // the entries of coal_lik_init()'s list the likelihoods read, BY POSITION like the reference
    arma::vec C, D, y, gridrep;
    int ng = 0;
    bool present = false;
    InitView() {}
    explicit InitView(List init) {
        C = as<arma::vec>(init[2]);
        D = as<arma::vec>(init[3]);
        y = as<arma::vec>(init[4]);
        gridrep = as<arma::vec>(init[6]);
        ng = as<int>(init[9]);
        present = true;
    }
};

lna_problem *problem_for(const arma::vec &t, const arma::vec &x_r, const int (&xi)[4], int gridsize, double t_correct,
                         const InitView &ini, int model, int transX = LNA_TRANSX_STANDARD) {
    static std::map<std::string, lna_problem *> cache;
    std::string key;
    key_add(key, t.memptr(), sizeof(double) * t.n_elem);
    key_add(key, x_r.memptr(), sizeof(double) * x_r.n_elem);
    key_add(key, xi, sizeof xi);
    key_add(key, &gridsize, sizeof gridsize);
    key_add(key, &t_correct, sizeof t_correct);
    key_add(key, &model, sizeof model);
    key_add(key, &transX, sizeof transX);
    if (ini.present) {
        key_add(key, ini.C.memptr(), sizeof(double) * ini.C.n_elem);
        key_add(key, ini.D.memptr(), sizeof(double) * ini.D.n_elem);
        key_add(key, ini.y.memptr(), sizeof(double) * ini.y.n_elem);
        key_add(key, ini.gridrep.memptr(), sizeof(double) * ini.gridrep.n_elem);
    }
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    lna_cfg cfg;
    cfg.model = model;
    cfg.transX = transX;
    cfg.nch = xi[0];
    cfg.nparam = xi[1];
    cfg.iR0 = xi[2];
    cfg.iGamma = xi[3];
    cfg.gridsize = gridsize;
    cfg.nt = (int)t.n_elem;
    cfg.times = t.memptr();
    cfg.x_r = x_r.memptr();
    cfg.t_correct = t_correct;
    lna_init li;
    if (ini.present) {
        li.E = (int)ini.C.n_elem;
        li.ng = ini.ng;
        li.C = ini.C.memptr();
        li.D = ini.D.memptr();
        li.y = ini.y.memptr();
        li.gridrep = ini.gridrep.memptr();
    }
    lna_problem *pb = lna_problem_create(&cfg, ini.present ? &li : nullptr, /*device=*/0);
    if (!pb) fail_abi("lna_problem_create");
    if (cache.size() > 64) {
        for (auto &e : cache) lna_problem_destroy(e.second);
        cache.clear();
    }
    cache[key] = pb;
    return pb;
}

int model_id(const std::string &model) {
    if (model == "SIR") return LNA_MODEL_SIR;
    if (model == "SEIR") return LNA_MODEL_SEIR;
    Rcpp::stop("GPU path: model must be \"SIR\" or \"SEIR\"");
}
int transx_id(const std::string &transX) {
    if (transX == "standard") return LNA_TRANSX_STANDARD;
    if (transX == "log") return LNA_TRANSX_LOG;  // (accepted for model SIR; the library says so otherwise)
    Rcpp::stop("GPU path: transX must be \"standard\" or \"log\"");
}
// transformPtr (:96-102) knows "changepoint" only; anything else is a NULL external pointer, i.e. an R error at first use
void check_transP(const std::string &transP) {
    if (transP != "changepoint") Rcpp::stop("external pointer is not valid");
}
lna_problem *problem_xi(const arma::vec &t, const arma::vec &x_r, const arma::ivec &x_i, int gridsize, double t_correct,
                        const InitView &ini, const std::string &model, const std::string &transX) {
    const int xi[4] = {(int)x_i[0], (int)x_i[1], (int)x_i[2], (int)x_i[3]};
    return problem_for(t, x_r, xi, gridsize, t_correct, ini, model_id(model), transx_id(transX));
}
// problems for calls that carry only a coarse trajectory: its own time column is the grid, gridsize 1
lna_problem *traj_problem(const arma::mat &traj, double t_correct, const InitView &ini, int iS, int iI,
                          int transX = LNA_TRANSX_STANDARD) {
    const int p = (int)traj.n_cols - 1;
    arma::vec times = traj.col(0), x_r(1);
    x_r[0] = 1.0;
    const int xi[4] = {0, p == 2 ? 2 : 3, iS, iI};
    return problem_for(times, x_r, xi, 1, t_correct, ini, p == 2 ? LNA_MODEL_SIR : LNA_MODEL_SEIR, transX);
}
struct Dims {
    int p, k, nrow, npt;
    explicit Dims(lna_problem *pb) { lna_problem_dims(pb, &p, &k, &nrow, &npt); }
};
// the integrator reads param[0 .. nparam+nch-1]; trailing entries are optional (testthat 'Test SIR model')
arma::vec pad_param(const arma::vec &param, int npt) {
    if ((int)param.n_elem < npt - 1) Rcpp::stop("param is shorter than nparam + nch");
    arma::vec q(npt);
    for (int i = 0; i < npt && i < (int)param.n_elem; ++i) q[i] = param[i];
    return q;
}
List make_FT(const arma::cube &A, const arma::cube &S) {
    List ft;
    ft["A"] = A;
    ft["Sigma"] = S;
    return ft;
}
arma::vec draw_uniforms(int n) {
    arma::vec u(n);
    for (int i = 0; i < n; ++i) u[i] = R::runif(0, 1);
    return u;
}
std::vector<int32_t> to_i32(const arma::ivec &v) {
    std::vector<int32_t> r(v.n_elem);
    for (unsigned i = 0; i < v.n_elem; ++i) r[i] = (int32_t)v[i];
    return r;
}

}  // namespace
