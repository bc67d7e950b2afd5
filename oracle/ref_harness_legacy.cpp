/*
 * oracle/ref_harness_legacy.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points over the reference's LEGACY copies of the pipeline, compiled verbatim where they lie:
 *   src/LNA_NS_general.cpp (linked), src/generalLNA.cpp (class Foo lives in that .cpp, so the file is #included here with
 *   RCPP_MODULE turned into an ordinary function), ESlice_SIRS of src/SIRS_period.cpp (linked).
 * Same rule as ref_harness.cpp: every ref_X has the C signature of orc_X in lna_oracle.h and only converts buffers.
 * In a drop-in build (-DLNA_DROPIN) the same entry points run integration/rcpp_shim_legacy.cpp (the reference-side binding
 * over the C ABI) instead of the reference's bodies.
 */
#include "basic_util.h"
#include "SIR_phylodyn.h"

#define RCPP_MODULE(name) static void rcpp_module_##name()
#ifdef LNA_DROPIN
#include "../integration/rcpp_shim_legacy.cpp"
#else
#include "generalLNA.cpp"
#endif

#include "ref_harness_util.h"

#ifdef LNA_DROPIN
List General_MCMC_with_ESlice_chains(List MCMC_setting, arma::mat par, arma::cube OriginTraj, arma::ivec ESS_vec, int niter,
                                     int burn1, int thin, double seed, arma::ivec devices);
#endif

// exported by src/LNA_NS_general.cpp and src/SIRS_period.cpp, declared in no header of theirs
arma::vec betafs(arma::vec ts, arma::vec param, arma::vec x_r, arma::ivec x_i);
double betaf(double t, arma::vec param, arma::vec x_r, arma::ivec x_i);
arma::vec ODE_general_one(arma::vec states, arma::vec param, double t, arma::vec x_r, arma::ivec x_i);
arma::mat general_F(arma::vec states, arma::vec param, double t, arma::vec x_r, arma::ivec x_i);
arma::mat General_ODE_rk45(arma::vec initial, arma::vec t, arma::vec param, arma::vec x_r, arma::ivec x_i);
List General_IntSigmaF(arma::mat Traj_par, arma::vec param, arma::vec x_r, arma::ivec x_i);
List General_KOM_Filter(arma::mat OdeTraj, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i);
double log_like_traj_general(arma::mat SdeTraj, arma::mat OdeTraj, List Filter, int gridsize, double t_correct);
List Traj_sim_general(arma::mat OdeTraj, List Filter, double t_correct);
List Traj_sim_ezG(arma::vec initial, arma::vec times, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i,
                  double t_correct);
arma::mat ESlice_general(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec betaN,
                         double t_correct, double lambda, int reps, int gridsize, bool volz);
arma::mat ESlice_SIRS(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec params4,
                      double t_correct, double lambda, int reps, int gridsize, bool volz);

namespace {
Foo make_foo(const double *A_pre, const double *A_post, const double *param, double N) {
    arma::vec x_r(1);
    x_r[0] = N;
    arma::ivec x_i(1);
    x_i[0] = 0;
    return Foo(M(A_pre, 3, 2), M(A_post, 3, 2), V(param, 3), x_i, x_r);
}
}  // namespace

extern "C" {

// betafs, LNA_NS_general.cpp:25 (element 0 also through betaf, :6)
void ref_general_betafs(const double *ts, int m, const double *param, const double *x_r, const int *x_i2, double *betas) {
    REF_TRY
    const int np = x_i2[1] + x_i2[0];
    arma::vec b = betafs(V(ts, m), V(param, np), V(x_r, x_i2[0] + 1), IV(x_i2, 2));
    if (m > 0 && !(betaf(ts[0], V(param, np), V(x_r, x_i2[0] + 1), IV(x_i2, 2)) == b[0])) g_err = "betaf(t[0]) != betafs(t)[0]";
    put(b, betas);
    REF_CATCH_VOID
}
// ODE_general_one :52 (mode 0), general_F :66 (mode 1)
void ref_general_point(int mode, const double *states, const double *param, int nparam_total, double t, const double *x_r,
                       const int *x_i2, double *out) {
    REF_TRY
    if (mode == 0) put(ODE_general_one(V(states, 2), V(param, nparam_total), t, V(x_r, x_i2[0] + 1), IV(x_i2, 2)), out);
    else put(general_F(V(states, 2), V(param, nparam_total), t, V(x_r, x_i2[0] + 1), IV(x_i2, 2)), out);
    REF_CATCH_VOID
}
// General_ODE_rk45 :76
void ref_general_ode_rk45(const double *initial, const double *t, int n, const double *param, int nparam_total,
                          const double *x_r, const int *x_i2, double *OdeTraj) {
    REF_TRY
    put(General_ODE_rk45(V(initial, 2), V(t, n), V(param, nparam_total), V(x_r, x_i2[0] + 1), IV(x_i2, 2)), OdeTraj);
    REF_CATCH_VOID
}
// General_IntSigmaF :103
void ref_general_intsigmaf(const double *Traj_par, int nrow, const double *param, int nparam_total, const double *x_r,
                           const int *x_i2, double *expF, double *Sigma) {
    REF_TRY
    List r = General_IntSigmaF(M(Traj_par, nrow, 3), V(param, nparam_total), V(x_r, x_i2[0] + 1), IV(x_i2, 2));
    put(as<arma::mat>(r[0]), expF);
    put(as<arma::mat>(r[1]), Sigma);
    REF_CATCH_VOID
}
// General_KOM_Filter :139
void ref_general_kom_filter(const double *OdeTraj, int nrow, const double *param, int nparam_total, int gridsize,
                            const double *x_r, const int *x_i2, double *Acube, double *Scube) {
    REF_TRY
    put_FT(General_KOM_Filter(M(OdeTraj, nrow, 3), V(param, nparam_total), gridsize, V(x_r, x_i2[0] + 1), IV(x_i2, 2)), Acube,
           Scube);
    REF_CATCH_VOID
}
// mode 0: log_like_traj_general :164, 1: Traj_sim_general :205, 2: Foo::LNA_log_like_traj (generalLNA.cpp:178),
// 3: Foo::LNA_Traj_sim (:217)
int ref_legacy_centred(int mode, const double *X_or_normals, const double *OdeTraj, int nrow, const double *Acube,
                       const double *Scube, double t_correct, double *traj_out, double *loglik) {
    REF_TRY
    List ft = make_FT(Acube, Scube, 2, nrow - 1);
    const double eye[6] = {1, 0, 0, 0, 1, 0}, zero[6] = {0, 0, 0, 0, 0, 0}, par3[3] = {1, 1, 0};
    if (mode == 0) {
        *loglik = log_like_traj_general(M(X_or_normals, nrow, 3), M(OdeTraj, nrow, 3), ft, 1, t_correct);
    } else if (mode == 2) {
        *loglik = make_foo(zero, eye, par3, 1.0).LNA_log_like_traj(M(X_or_normals, nrow, 3), M(OdeTraj, nrow, 3), ft, 1, t_correct);
    } else {
        streams(X_or_normals, (long)(nrow - 1) * 2, nullptr, 0);
        List r = mode == 1 ? Traj_sim_general(M(OdeTraj, nrow, 3), ft, t_correct)
                           : make_foo(zero, eye, par3, 1.0).LNA_Traj_sim(M(OdeTraj, nrow, 3), ft, t_correct);
        put(as<arma::mat>(r[0]), traj_out);
        *loglik = as<double>(r[1]);
    }
    return ORC_OK;
    REF_CATCH_STATUS
}
// Traj_sim_ezG :257
int ref_traj_sim_ezG(const double *initial, const double *t, int nt, const double *param, int nparam_total, int gridsize,
                     const double *x_r, const int *x_i2, double t_correct, const double *normals, double *LNA_traj,
                     double *loglike) {
    REF_TRY
    streams(normals, (long)(nt / gridsize) * 2, nullptr, 0);
    List r = Traj_sim_ezG(V(initial, 2), V(t, nt), V(param, nparam_total), gridsize, V(x_r, x_i2[0] + 1), IV(x_i2, 2), t_correct);
    put(as<arma::mat>(r[0]), LNA_traj);
    *loglike = as<double>(r[1]);
    return ORC_OK;
    REF_CATCH_STATUS
}
// volz_loglik_nh, SIR_phylodyn.cpp:517 (f1 = a LOG trajectory)
double ref_volz_loglik_nh(const orc_init *init, const double *f1, int nrow, const double *betaN, double t_correct) {
    REF_TRY
    return volz_loglik_nh(make_init(init), M(f1, nrow, 3), V(betaN, nrow), t_correct, 1);
    REF_CATCH_NAN
}
// kind 0: ESlice_general (LNA_NS_general.cpp:320), 1: ESlice_SIRS (SIRS_period.cpp:313), 2: Foo::LNA_ESlice (generalLNA.cpp:296)
int ref_eslice_legacy(const orc_init *init, int kind, const double *f_cur, int nrow, int p, const double *OdeTraj,
                      const double *Acube, const double *Scube, const double *state, const double *betaN_or_params4,
                      double t_correct, double lambda, int reps, int volz, const double *normals, const double *uniforms,
                      int *n_unif, int *n_evals, double *newTraj) {
    *n_evals = -1;
    REF_TRY
    streams(normals, (long)reps * (nrow - 1) * p, uniforms, (long)reps * 22);
    List ft = make_FT(Acube, Scube, p, nrow - 1);
    const double eye[6] = {1, 0, 0, 0, 1, 0}, zero[6] = {0, 0, 0, 0, 0, 0}, par3[3] = {1, 1, 0};
    arma::mat r;
    if (kind == 0)
        r = ESlice_general(M(f_cur, nrow, p + 1), M(OdeTraj, nrow, p + 1), ft, V(state, p), make_init(init),
                           V(betaN_or_params4, nrow), t_correct, lambda, reps, 1, volz != 0);
    else if (kind == 1)
        r = ESlice_SIRS(M(f_cur, nrow, p + 1), M(OdeTraj, nrow, p + 1), ft, V(state, p), make_init(init), V(betaN_or_params4, 4),
                        t_correct, lambda, reps, 1, volz != 0);
    else
        r = make_foo(zero, eye, par3, 1.0).LNA_ESlice(M(f_cur, nrow, p + 1), M(OdeTraj, nrow, p + 1), ft, V(state, p),
                                                      make_init(init), V(betaN_or_params4, nrow), t_correct, lambda, reps, 1,
                                                      volz != 0);
    *n_unif = (int)shim_rng().used_unif;
    put(r, newTraj);
    return ORC_OK;
    REF_CATCH_STATUS
}
// Foo::rate_h (generalLNA.cpp:76, mode 0), dh (:84, 1), ODE_ones (:96, 2); mode 3 = F_vec (:92), whose arma::vec return type
// cannot hold the 2 x 2 product it computes: whatever the library does with that is reported, not computed here
void ref_foo_point(int mode, const double *A_pre, const double *A_post, const double *param, double N, const double *states,
                   double *out) {
    REF_TRY
    Foo f = make_foo(A_pre, A_post, param, N);
    if (mode == 0) put(f.rate_h(V(states, 2), 0.0), out);
    else if (mode == 1) put(f.dh(V(states, 2), 0.0), out);
    else if (mode == 2) put(f.ODE_ones(V(states, 2), 0.0), out);
    else {
        arma::mat r = f.F_vec(V(states, 2), 0.0);
        if (r.n_elem != 4) throw std::logic_error("Foo::F_vec: a 2 x 2 product does not fit its arma::vec return type");
        put(r, out);
    }
    REF_CATCH_VOID
}
// Foo::ODE_rk45 :100
void ref_foo_ode_rk45(const double *A_pre, const double *A_post, const double *param, double N, const double *initial,
                      const double *t, int n, double *OdeTraj) {
    REF_TRY
    put(make_foo(A_pre, A_post, param, N).ODE_rk45(V(initial, 2), V(t, n)), OdeTraj);
    REF_CATCH_VOID
}
// Foo::IntSigmaF (:124) / KOM_Filter_MX (:158): returns 0 if the reference's code ran, -1 (+ message) if it threw
int ref_foo_intsigmaf_runs(const double *A_pre, const double *A_post, const double *param, double N, const double *Traj_par,
                           int nrow) {
    REF_TRY
    make_foo(A_pre, A_post, param, N).IntSigmaF(M(Traj_par, nrow, 3));
    return 0;
    REF_CATCH_STATUS
}

#ifdef LNA_DROPIN
// General_MCMC_with_ESlice_chains (integration/rcpp_chains.cpp): the batched entry of the binding; no reference counterpart
// (the reference runs one chain per call), so this entry point exists in the drop-in build only
int ref_chains_run(const orc_cfg *c, const orc_init *init, const double *times, int nt, int gridsize, double t_correct,
                   const double *prior10, const double *proposal5, const int *ess7, long B, const double *par,
                   const double *eps, int eps_shared, int niter, int burn1, int thin, double seed, const int *devices, int ndev,
                   double *par_hist, double *l, double *l1, double *par_final, double *latent, double *coalLog) {
    REF_TRY
    const int nch = c->x_i[0], npar = 2 + c->nparam_total, k = (nt - 1) / gridsize, p = c->p;
    List setting, prior(5), proposal(5);
    for (int i = 0; i < 5; ++i) {
        prior[i] = V(prior10 + 2 * i, 2);
        proposal[i] = V(proposal5 + i, 1);
    }
    arma::vec xi(4);
    for (int i = 0; i < 4; ++i) xi[i] = c->x_i[i];
    setting["times"] = V(times, nt);
    setting["x_r"] = V(c->x_r, nch + 1);
    setting["x_i"] = xi;
    setting["t_correct"] = t_correct;
    setting["gridsize"] = gridsize;
    setting["prior"] = prior;
    setting["proposal"] = proposal;
    setting["Init"] = make_init(init);
    List r = General_MCMC_with_ESlice_chains(setting, M(par, npar, B), CU(eps, k, p, eps_shared ? 1 : B), IV(ess7, 7), niter, burn1,
                                             thin, seed, IV(devices, ndev));
    put(as<arma::cube>(r["par"]), par_hist);
    put(as<arma::mat>(r["l"]), l);
    put(as<arma::mat>(r["l1"]), l1);
    put(as<arma::mat>(r["par_final"]), par_final);
    put(as<arma::cube>(r["LatentTraj"]), latent);
    put(as<arma::mat>(r["coalLog"]), coalLog);
    return ORC_OK;
    REF_CATCH_STATUS
}
#endif

}  // extern "C"
