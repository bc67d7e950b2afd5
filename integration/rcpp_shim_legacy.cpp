// integration/rcpp_shim_legacy.cpp -- the reference-side binding of the LEGACY copies of the pipeline.
//
// Replaces, with the reference's exact C++ signatures, the bodies of
//   src/LNA_NS_general.cpp : betaf, betafs, ODE_general_one, general_F, General_ODE_rk45, General_IntSigmaF,
//                            General_KOM_Filter, log_like_traj_general, Traj_sim_general, Traj_sim_ezG, ESlice_general
//   src/SIRS_period.cpp    : Traj_sim_SIRS, ESlice_SIRS          (its ODE2 / SIRS_KOM_Filter / log_like_trajSIRS: rcpp_shim.cpp)
//   src/SIR_phylodyn.cpp   : volz_loglik_nh
//   src/generalLNA.cpp     : class Foo and its RCPP_MODULE(mod_Foo) registration
// and forwards to the C ABI (include/lnaphylodyn_b200.h) with B = 1.  The integrator / moment functions are arithmetic
// twins of the current ones on the SIR model and go through rcpp_shim.cpp's exports with x_i = (nch, nparam, 0, 1).
// Random numbers are drawn in the reference's order (per repetition: randn(k*p) for the proposal, then u, theta_1 and one
// uniform per shrink) and the generator is left where the unmodified package would leave it (RngMark).
//
// Compiled and exercised like rcpp_shim.cpp: `make -C oracle dropin` includes this file into oracle/ref_harness_legacy.cpp
// on the stand-in headers; tests/test_gpu_dropin_legacy.py runs it (GPU) beside the reference's own bodies (CPU).
#include "rcpp_shim_util.h"

// exports of rcpp_shim.cpp (signatures of src/LNA_functional.h)
arma::vec betaTs(arma::vec param, arma::vec times, arma::vec x_r, arma::ivec x_i);
arma::mat ODE_rk45(arma::vec initial, arma::vec t, arma::vec param, arma::vec x_r, arma::ivec x_i, std::string transP,
                   std::string model, std::string transX);
List SigmaF(arma::mat Traj_par, arma::vec param, arma::vec x_r, arma::ivec x_i, std::string transP, std::string model,
            std::string transX);
List KF_param(arma::mat OdeTraj, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i, std::string transP,
              std::string model, std::string transX);
arma::mat Ode_Coarse_Slicer(arma::mat Ode_thin, int gridsize);

namespace {

// legacy packs: x_i = (nch, nparam), param = (R0, gamma, .., ch_1..ch_nch) -> the current layout (trailing hyper unused)
arma::ivec legacy_xi(const arma::ivec &x_i) {
    arma::ivec xi(4);
    xi[0] = x_i[0];
    xi[1] = x_i[1];
    xi[2] = 0;
    xi[3] = 1;
    return xi;
}
arma::vec legacy_param(const arma::vec &param, const arma::ivec &x_i) {
    const int n = (int)(x_i[0] + x_i[1]);
    if ((int)param.n_elem < n) Rcpp::stop("param is shorter than nparam + nch");
    arma::vec q(n + 1);
    for (int i = 0; i < n; ++i) q[i] = param[i];
    q[n] = 1.0;
    return q;
}
// the SIR effect matrix as a Foo network (3 x 2 column-major; the third reaction is empty)
const double kSirPre[6] = {1, 0, 0, 1, 1, 0}, kSirPost[6] = {0, 0, 0, 2, 0, 0};

lna_problem *any_problem() {
    arma::vec t01(2), x_r(1);
    t01[0] = 0.0;
    t01[1] = 1.0;
    x_r[0] = 1.0;
    const int xi[4] = {0, 2, 0, 1};
    return problem_for(t01, x_r, xi, 1, 0.0, InitView(), LNA_MODEL_SIR);
}
void foo_call(int mode, int n, const double *A_pre, const double *A_post, const double *par3, double N, const double *x,
              const double *t, double *out) {
    LNA_CALL(lna_foo(any_problem(), 1, mode, n, A_pre, A_post, par3, &N, x, t, out), "lna_foo");
}

// one repetition of the centred trajectory slice update (kind 0 / 2: ESlice_general, Foo::LNA_ESlice; kind 1: ESlice_SIRS)
arma::mat eslice_legacy_once(int kind, const arma::mat &f_cur, const arma::mat &OdeTraj, List FTs, const arma::vec &state,
                             List init, const arma::vec &betaN_or_params4, double t_correct, double lambda, bool volz) {
    const int k = (int)OdeTraj.n_rows - 1, p = (int)OdeTraj.n_cols - 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(init), 0, p == 2 ? 1 : 2);
    arma::cube A = as<arma::cube>(FTs[0]), S = as<arma::cube>(FTs[1]);
    arma::vec z = arma::randn(k * p, 1);  // the proposal's mvrnormArma / mvrnormArma2 draws, one randn(p, 1) per step
    RngMark rng;
    rng.mark();
    arma::vec un = draw_uniforms(22);
    arma::mat newTraj(k + 1, p + 1);
    int32_t ne = 0, nu = 0, st = 0;
    LNA_CALL(lna_eslice_legacy(pb, 1, kind, 1, volz ? 1 : 0, f_cur.memptr(), OdeTraj.memptr(), A.memptr(), S.memptr(),
                               state.memptr(), betaN_or_params4.memptr(), &lambda, z.memptr(), un.memptr(), newTraj.memptr(),
                               &ne, &nu, &st),
             "lna_eslice_legacy");
    if (st & LNA_ST_CHOL_FAIL) {
        rng.settle(0);  // arma::chol threw inside the proposal simulation, before any uniform was drawn
        throw std::runtime_error("chol(): decomposition failed");
    }
    rng.settle(nu);
    return newTraj;
}

}  // namespace

// ---- src/LNA_NS_general.cpp -------------------------------------------------------------------------------------------
arma::vec betafs(arma::vec ts, arma::vec param, arma::vec x_r, arma::ivec x_i) {  // :25-47
    return betaTs(legacy_param(param, x_i), ts, x_r, legacy_xi(x_i));
}
double betaf(double t, arma::vec param, arma::vec x_r, arma::ivec x_i) {  // :6-21
    arma::vec ts(1);
    ts[0] = t;
    return betafs(ts, param, x_r, x_i)[0];
}
arma::vec ODE_general_one(arma::vec states, arma::vec param, double t, arma::vec x_r, arma::ivec x_i) {  // :52-62
    const double par3[3] = {betaf(t, param, x_r, x_i), param[1], 0.0};
    arma::vec res(2);
    foo_call(5, 0, kSirPre, kSirPost, par3, 1.0, states.memptr(), nullptr, res.memptr());
    return res;
}
arma::mat general_F(arma::vec states, arma::vec param, double t, arma::vec x_r, arma::ivec x_i) {  // :66-72
    const double par3[3] = {betaf(t, param, x_r, x_i), param[1], 0.0};
    arma::mat M(2, 2);
    foo_call(6, 0, kSirPre, kSirPost, par3, 1.0, states.memptr(), nullptr, M.memptr());
    return M;
}
arma::mat General_ODE_rk45(arma::vec initial, arma::vec t, arma::vec param, arma::vec x_r, arma::ivec x_i) {  // :76-98
    return ODE_rk45(initial, t, legacy_param(param, x_i), x_r, legacy_xi(x_i), "changepoint", "SIR", "standard");
}
List General_IntSigmaF(arma::mat Traj_par, arma::vec param, arma::vec x_r, arma::ivec x_i) {  // :103-135
    List r = SigmaF(Traj_par, legacy_param(param, x_i), x_r, legacy_xi(x_i), "changepoint", "SIR", "standard");
    List Res;
    Res["expF"] = as<arma::mat>(r[0]);
    Res["Simga"] = as<arma::mat>(r[1]);  // (the reference's spelling)
    return Res;
}
List General_KOM_Filter(arma::mat OdeTraj, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i) {  // :139-159
    return KF_param(OdeTraj, legacy_param(param, x_i), gridsize, x_r, legacy_xi(x_i), "changepoint", "SIR", "standard");
}
double log_like_traj_general(arma::mat SdeTraj, arma::mat OdeTraj, List Filter, int gridsize, double t_correct) {  // :164-200
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, 1);
    arma::cube A = as<arma::cube>(Filter[0]), S = as<arma::cube>(Filter[1]);
    double ll = 0;
    LNA_CALL(lna_log_like_traj_general(pb, 1, SdeTraj.memptr(), OdeTraj.memptr(), A.memptr(), S.memptr(), &ll),
             "lna_log_like_traj_general");
    return ll;
}
List Traj_sim_general(arma::mat OdeTraj, List Filter, double t_correct) {  // :205-250
    const int k = (int)OdeTraj.n_rows - 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, 1);
    arma::cube A = as<arma::cube>(Filter[0]), S = as<arma::cube>(Filter[1]);
    arma::vec z = arma::randn(k * 2, 1);
    arma::mat traj(k + 1, 3);
    double ll = 0;
    LNA_CALL(lna_traj_sim_general(pb, 1, OdeTraj.memptr(), A.memptr(), S.memptr(), z.memptr(), traj.memptr(), &ll),
             "lna_traj_sim_general");
    List Res;
    Res["SimuTraj"] = traj;
    Res["loglike"] = ll;
    return Res;
}
List Traj_sim_ezG(arma::vec initial, arma::vec times, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i,
                  double t_correct) {  // :257-314
    arma::mat ode = General_ODE_rk45(initial, times, param, x_r, x_i);
    List kf = General_KOM_Filter(ode, param, gridsize, x_r, x_i);
    return Traj_sim_general(Ode_Coarse_Slicer(ode, gridsize), kf, t_correct);
}
arma::mat ESlice_general(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec betaN,
                         double t_correct, double lambda = 10, int reps = 1, int gridsize = 100, bool volz = false) {  // :320-394
    arma::mat cur = f_cur;
    for (int count = 0; count < reps; ++count)
        cur = eslice_legacy_once(0, cur, OdeTraj, FTs, state, init, betaN, t_correct, lambda, volz);
    return cur;
}

// ---- src/SIR_phylodyn.cpp ---------------------------------------------------------------------------------------------
double volz_loglik_nh(List init, arma::mat f1, arma::vec betaN, double t_correct, int gridsize) {  // :517-560
    lna_problem *pb = traj_problem(f1, t_correct, InitView(init), 0, 1);
    double ll = 0;
    LNA_CALL(lna_volz_loglik_nh(pb, 1, f1.memptr(), betaN.memptr(), &ll), "lna_volz_loglik_nh");
    return ll;
}

// ---- src/SIRS_period.cpp ----------------------------------------------------------------------------------------------
List Traj_sim_SIRS(arma::vec initial, arma::mat OdeTraj, List Filter, double t_correct) {  // :197-240
    const int k = (int)OdeTraj.n_rows - 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, 2);
    arma::cube A = as<arma::cube>(Filter[0]), S = as<arma::cube>(Filter[1]);
    arma::vec z = arma::randn(k * 3, 1);
    arma::mat traj(k + 1, 4);
    double ll = 0;
    int32_t st = 0;
    LNA_CALL(lna_traj_sim_sirs(pb, 1, initial.memptr(), OdeTraj.memptr(), A.memptr(), S.memptr(), z.memptr(), traj.memptr(),
                               &ll, &st),
             "lna_traj_sim_sirs");
    if (st & LNA_ST_CHOL_FAIL) throw std::runtime_error("chol(): decomposition failed");
    List Res;
    Res["SimuTraj"] = traj;
    Res["loglike"] = ll;
    return Res;
}
arma::mat ESlice_SIRS(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec params4,
                      double t_correct, double lambda = 10, int reps = 1, int gridsize = 100, bool volz = true) {  // :313-402
    arma::mat cur = f_cur;
    for (int count = 0; count < reps; ++count)
        cur = eslice_legacy_once(1, cur, OdeTraj, FTs, state, init, params4, t_correct, lambda, volz);
    return cur;
}

// ---- src/generalLNA.cpp: class Foo ------------------------------------------------------------------------------------
class Foo {
   private:
    arma::mat A_pre, A_post, A;

   public:
    arma::vec param;
    arma::ivec x_i;
    arma::vec x_r;
    int m, p;
    double N;

    Foo(arma::mat A_pre_, arma::mat A_post_, arma::vec param_, arma::ivec x_i_, arma::vec x_r_) {
        if (A_pre_.n_rows != 3 || A_pre_.n_cols != 2 || A_post_.n_rows != 3 || A_post_.n_cols != 2)
            Rcpp::stop("GPU path: Foo's hazards are written for 3 reactions on 2 species (generalLNA.cpp:76-90)");
        A_pre = A_pre_;
        A_post = A_post_;
        A = A_post - A_pre;
        param = param_;
        x_i = x_i_;
        x_r = x_r_;
        p = (int)A_pre.n_cols;
        m = (int)A_pre.n_rows;
        N = x_r[0];
    }
    arma::mat get_A_pre() { return A_pre; }
    arma::mat get_A_post() { return A_post; }
    arma::mat get_A() { return A; }
    arma::vec get_param() { return param; }
    void set_param(arma::vec new_param) { param = new_param; }
    void set_x_i(arma::ivec new_x_i) { x_i = new_x_i; }
    void set_x_r(arma::vec new_x_r) { x_r = new_x_r; }

    arma::vec transform_param(double t) {  // :50-55
        arma::vec one(2);
        one[0] = one[1] = 1.0;
        arma::vec h = rate_h(one, t), res = param;  // param_new(0) = param(0) * param(1) / N = h(0) at S = I = 1
        res[0] = h[0];
        return res;
    }
    arma::vec rate_h(arma::vec states, double t) {  // :76-83
        arma::vec h(3);
        foo_call(0, 0, A_pre.memptr(), A_post.memptr(), param.memptr(), N, states.memptr(), nullptr, h.memptr());
        return h;
    }
    arma::mat dh(arma::vec states, double t) {  // :84-90
        arma::mat res(3, 2);
        foo_call(1, 0, A_pre.memptr(), A_post.memptr(), param.memptr(), N, states.memptr(), nullptr, res.memptr());
        return res;
    }
    arma::vec F_vec(arma::vec states, double t) {  // :92-94: a 2 x 2 product returned through an arma::vec
        Rcpp::stop("Mat::init(): requested size is not compatible with column vector layout");
        return arma::vec();
    }
    arma::vec ODE_ones(arma::vec states, double t) {  // :96-98
        arma::vec res(2);
        foo_call(2, 0, A_pre.memptr(), A_post.memptr(), param.memptr(), N, states.memptr(), nullptr, res.memptr());
        return res;
    }
    arma::mat ODE_rk45(arma::vec initial, arma::vec t) {  // :100-122
        arma::mat out(t.n_elem, 3);
        foo_call(4, (int)t.n_elem, A_pre.memptr(), A_post.memptr(), param.memptr(), N, initial.memptr(), t.memptr(),
                 out.memptr());
        return out;
    }
    List IntSigmaF(arma::mat Traj_par) {  // :124-156 calls F_vec and multiplies by an empty local `A`: it throws as written
        F_vec(arma::vec(2), 0.0);
        return List();
    }
    List KOM_Filter_MX(arma::mat OdeTraj, int gridsize) { return IntSigmaF(OdeTraj); }  // :158-176
    double LNA_log_like_traj(arma::mat SdeTraj, arma::mat OdeTraj, List Filter, int gridsize, double t_correct) {  // :178-214
        return log_like_traj_general(SdeTraj, OdeTraj, Filter, gridsize, t_correct);
    }
    List LNA_Traj_sim(arma::mat OdeTraj, List Filter, double t_correct) {  // :217-262
        return Traj_sim_general(OdeTraj, Filter, t_correct);
    }
    List LNA_Traj_sim_ez(arma::vec initial, arma::vec times, int gridsize, double t_correct) {  // :264-294 (through KOM_Filter_MX)
        return IntSigmaF(arma::mat());
    }
    arma::mat LNA_ESlice(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec betaN,
                         double t_correct, double lambda = 10, int reps = 1, int gridsize = 100, bool volz = false) {  // :296-372
        arma::mat cur = f_cur;
        for (int count = 0; count < reps; ++count)
            cur = eslice_legacy_once(2, cur, OdeTraj, FTs, state, init, betaN, t_correct, lambda, volz);
        return cur;
    }
};

RCPP_MODULE(mod_Foo) {
    using namespace Rcpp;
    class_<Foo>("Foo")
        .constructor<arma::mat, arma::mat, arma::vec, arma::ivec, arma::vec>()
        .field("x_i", &Foo::x_i)
        .field("x_r", &Foo::x_r)
        .field("param", &Foo::param)
        .field("N", &Foo::N)
        .field("m", &Foo::m)
        .field("p", &Foo::p)
        .method("get_param", &Foo::get_param)
        .method("get_A", &Foo::get_A)
        .method("IntSigmaF", &Foo::IntSigmaF)
        .method("KOM_Filter_MX", &Foo::KOM_Filter_MX)
        .method("set_param", &Foo::set_param)
        .method("LNA_Traj_sim", &Foo::LNA_Traj_sim)
        .method("LNA_log_like_traj", &Foo::LNA_log_like_traj)
        .method("set_x_i", &Foo::set_x_i)
        .method("set_x_r", &Foo::set_x_r)
        .method("get_A_pre", &Foo::get_A_pre)
        .method("get_A_post", &Foo::get_A_post)
        .method("transform_param", &Foo::transform_param)
        .method("rate_h", &Foo::rate_h)
        .method("dh", &Foo::dh)
        .method("F_vec", &Foo::F_vec)
        .method("ODE_ones", &Foo::ODE_ones)
        .method("ODE_rk45", &Foo::ODE_rk45)
        .method("LNA_Traj_sim_ez", &Foo::LNA_Traj_sim_ez)
        .method("LNA_ESlice", &Foo::LNA_ESlice);
}
