/* oracle/ref_harness_util.h -- TEST INFRASTRUCTURE ONLY: buffer <-> arma / Rcpp conversions shared by ref_harness.cpp and
 * ref_harness_legacy.cpp (no arithmetic of the path lives here). */
#pragma once
#include "lna_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>

std::string &ref_err();  // the thread's last error (defined in ref_harness.cpp)
#define g_err ref_err()

namespace {

arma::vec V(const double *p, long n) {
    arma::vec v(n);
    if (n) std::memcpy(v.memptr(), p, sizeof(double) * (size_t)n);
    return v;
}
arma::mat M(const double *p, long r, long c) {
    arma::mat m(r, c);
    if (r * c) std::memcpy(m.memptr(), p, sizeof(double) * (size_t)(r * c));
    return m;
}
arma::cube CU(const double *p, long r, long c, long s) {
    arma::cube q(r, c, s);
    if (r * c * s) std::memcpy(q.memptr(), p, sizeof(double) * (size_t)(r * c * s));
    return q;
}
arma::ivec IV(const int *p, long n) {
    arma::ivec v(n);
    for (long i = 0; i < n; ++i) v[i] = p[i];
    return v;
}
void put(const arma::mat &m, double *dst) {
    if (dst && m.n_elem) std::memcpy(dst, m.memptr(), sizeof(double) * m.n_elem);
}
void put(const arma::cube &m, double *dst) {
    if (dst && m.n_elem) std::memcpy(dst, m.memptr(), sizeof(double) * m.n_elem);
}

struct Pack {  // the (x_r, x_i, transP, model, transX) argument pack of every reference call
    arma::vec x_r;
    arma::ivec x_i;
    std::string transP = "changepoint", model, transX;
    explicit Pack(const orc_cfg *c) {
        x_r = V(c->x_r, c->x_i[0] + 1);
        x_i = IV(c->x_i, 4);
        model = c->model == ORC_SIR ? "SIR" : c->model == ORC_SEIR ? "SEIR" : "SEIR2";
        transX = c->transX == ORC_LOG ? "log" : "standard";
    }
};
const char *transx_name(int t) { return t == ORC_LOG ? "log" : "standard"; }

// coal_lik_init()'s list, positions as the reference reads them: [2] C, [3] D, [4] y, [6] gridrep, [9] ng
List make_init(const orc_init *in) {
    List l(12);
    for (int i = 0; i < 12; ++i) l[i] = 0.0;
    l[2] = V(in->C, in->E);
    l[3] = V(in->D, in->E);
    l[4] = V(in->y, in->E);
    l[6] = V(in->gridrep, in->ng);
    l[9] = in->ng;
    return l;
}
List make_FT(const double *A, const double *S, int p, int k) {
    List ft;
    ft["A"] = CU(A, p, p, k);
    ft["Sigma"] = CU(S, p, p, k);
    return ft;
}
void put_FT(List ft, double *A, double *S) {
    put(as<arma::cube>(ft[0]), A);
    put(as<arma::cube>(ft[1]), S);
}
List make_priors(const double *pr8) {
    List l(4);
    for (int i = 0; i < 4; ++i) l[i] = V(pr8 + 2 * i, 2);
    return l;
}
void streams(const double *normals, long nn, const double *uniforms, long nu) {
    shim_rng_state &s = shim_rng();
    s.normals = normals;
    s.uniforms = uniforms;
    s.n_norm = normals ? nn : 0;
    s.n_unif = uniforms ? nu : 0;
    s.used_norm = s.used_unif = 0;
}

}  // namespace

// status for calls that return one; the reference signals a failed factorisation by std::invalid_argument
#define REF_TRY try {
#define REF_CATCH_STATUS                              \
    }                                                 \
    catch (const std::invalid_argument &e) {          \
        g_err = e.what();                             \
        return ORC_CHOL_FAIL;                         \
    }                                                 \
    catch (const std::exception &e) {                 \
        g_err = e.what();                             \
        return -1;                                    \
    }
#define REF_CATCH_VOID                                \
    }                                                 \
    catch (const std::exception &e) { g_err = e.what(); }
#define REF_CATCH_NAN                                 \
    }                                                 \
    catch (const std::exception &e) {                 \
        g_err = e.what();                             \
        return NAN;                                   \
    }

