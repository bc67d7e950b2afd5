/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points over the REFERENCE'S OWN SOURCES, compiled verbatim from /root/reference/src/
 * (LNA_functional.cpp, basic_util.cpp, SIR_LNA.cpp, SIRS_period.cpp, SIR_phylodyn.cpp) against the
 * stand-in headers of oracle/shim/ (R / Rcpp / RcppArmadillo are absent from the image).  The result,
 * oracle/_ref/liblna_ref.so, is "the reference code as written" and is used
 *   (1) to pin the C oracle (oracle/lna_oracle.c) function by function, including the Metropolis and
 *       elliptical-slice loops with injected random streams (tests/test_ref_vs_oracle.py),
 *   (2) to generate tests/golden/ref_*.npz (tests/golden/make_ref_golden.py), and
 *   (3) as the timed CPU baseline of bench.py (`cpu_baseline.kind = "reference"`).
 *
 * Every ref_X below has the SAME C signature as orc_X in oracle/lna_oracle.h and does nothing but convert
 * the raw buffers to the by-value arma / Rcpp arguments of the reference function named in its comment,
 * call it, and copy the result out.  No arithmetic of the path lives here.
 */
#include "LNA_functional.h"  // the reference's header (-I/root/reference/src)
#include "lna_oracle.h"      // orc_cfg / orc_init / status bits

#include <cfloat>
#include <cstring>

// exported by the reference sources but declared in no header of theirs
double volz_loglik_nh3(List init, arma::mat f1, arma::vec betaN, double t_correct, arma::ivec index, std::string transX);
double ODE_rk45_stop(arma::vec initial, arma::vec t, arma::vec param, arma::vec x_r, arma::ivec x_i, std::string transP,
                     std::string model, std::string transX, double tol);
arma::mat Ode_Coarse_Slicer(arma::mat Ode_thin, int gridsize);
List New_Param_List(arma::vec param, arma::vec initial, int gridsize, arma::vec t, arma::vec x_r, arma::ivec x_i,
                    std::string transP, std::string model, std::string transX);
List Update_Param(arma::vec param, arma::vec initial, arma::vec t, arma::mat OriginTraj, arma::vec x_r, arma::ivec x_i,
                  List init, int gridsize, double coal_log, double prior_proposal_offset, double t_correct,
                  std::string transP, std::string model, std::string transX, bool volz);
arma::vec Param_Slice_update(arma::vec param, arma::vec x_r, arma::ivec x_i, double theta, arma::vec newChs, double rho);
arma::vec Param_Slice_update_all2(arma::vec par_old, arma::vec x_r, arma::ivec x_i, double theta, arma::vec newChs,
                                  arma::ivec ESS_vec, List priorList);
List ESlice_change_points(arma::vec param, arma::vec initial, arma::vec t, arma::mat OriginTraj, arma::vec x_r,
                          arma::ivec x_i, List init, int gridsize, double coal_log, double t_correct, std::string transP,
                          std::string model, std::string transX, bool volz);
List ESlice_par_General(arma::vec par_old, arma::vec t, arma::mat OriginTraj, List priorList, arma::vec x_r,
                        arma::ivec x_i, List init, int gridsize, arma::ivec ESS_vec, double coal_log, double t_correct,
                        std::string transP, std::string model, std::string transX, bool volz);
arma::mat ODE2(arma::vec initial, arma::vec t, arma::vec param, double period);
List SIRS_KOM_Filter(arma::mat OdeTraj, arma::vec param, int gridsize, double period);
double log_like_trajSIRS(arma::mat SdeTraj, arma::mat OdeTraj, List Filter, int gridsize, double t_correct);
List Traj_sim_SIRS(arma::vec initial, arma::mat OdeTraj, List Filter, double t_correct);

#include "ref_harness_util.h"

std::string &ref_err() {
    static thread_local std::string e;
    return e;
}

extern "C" {

const char *ref_last_error(void) { return g_err.c_str(); }
void ref_clear_error(void) { g_err.clear(); }

// betaTs, LNA_functional.cpp:31
void ref_betaTs(const double *param, const double *times, int m, const double *x_r, const int *x_i, double *betas) {
    REF_TRY
    const int np = x_i[1] + x_i[0] + 1;
    put(betaTs(V(param, np), V(times, m), V(x_r, x_i[0] + 1), IV(x_i, 4)), betas);
    REF_CATCH_VOID
}
// param_transform, :63
void ref_param_transform(double t, const double *param, int nparam_total, const double *x_r, const int *x_i,
                         double *param2) {
    REF_TRY
    put(param_transform(t, V(param, nparam_total), V(x_r, x_i[0] + 1), IV(x_i, 4)), param2);
    REF_CATCH_VOID
}
// ODE_rk45, :455
void ref_ode_rk45(const orc_cfg *c, const double *initial, const double *t, int n, const double *param, double *OdeTraj) {
    REF_TRY
    Pack k(c);
    put(ODE_rk45(V(initial, c->p), V(t, n), V(param, c->nparam_total), k.x_r, k.x_i, k.transP, k.model, k.transX), OdeTraj);
    REF_CATCH_VOID
}
// ODE_rk45_stop, :495
double ref_ode_rk45_stop(const orc_cfg *c, const double *initial, const double *t, int n, const double *param, double tol) {
    REF_TRY
    Pack k(c);
    return ODE_rk45_stop(V(initial, c->p), V(t, n), V(param, c->nparam_total), k.x_r, k.x_i, k.transP, k.model, k.transX, tol);
    REF_CATCH_NAN
}
// SigmaF, :536 (Traj_par: nrow rows of a column-major matrix with leading dimension ld)
void ref_sigmaF(const orc_cfg *c, const double *Traj_par, int nrow, int ld, const double *param, double *expF,
                double *Sigma) {
    REF_TRY
    Pack k(c);
    arma::mat T(nrow, c->p + 1);
    for (int j = 0; j <= c->p; ++j)
        for (int i = 0; i < nrow; ++i) T(i, j) = Traj_par[i + (size_t)j * ld];
    List r = SigmaF(T, V(param, c->nparam_total), k.x_r, k.x_i, k.transP, k.model, k.transX);
    put(as<arma::mat>(r[0]), expF);
    put(as<arma::mat>(r[1]), Sigma);
    REF_CATCH_VOID
}
// KF_param, :603
void ref_kf_param(const orc_cfg *c, const double *OdeTraj, int nrow, const double *param, int gridsize, double *Acube,
                  double *Scube) {
    REF_TRY
    Pack k(c);
    put_FT(KF_param(M(OdeTraj, nrow, c->p + 1), V(param, c->nparam_total), gridsize, k.x_r, k.x_i, k.transP, k.model, k.transX),
           Acube, Scube);
    REF_CATCH_VOID
}
// KF_param_chol, :638
int ref_kf_param_chol(const orc_cfg *c, const double *OdeTraj, int nrow, const double *param, int gridsize, double *Acube,
                      double *Lcube) {
    REF_TRY
    Pack k(c);
    put_FT(KF_param_chol(M(OdeTraj, nrow, c->p + 1), V(param, c->nparam_total), gridsize, k.x_r, k.x_i, k.transP, k.model,
                         k.transX),
           Acube, Lcube);
    return ORC_OK;
    REF_CATCH_STATUS
}
// Ode_Coarse_Slicer, :1179
void ref_ode_coarse_slicer(const double *Ode_thin, int nrow, int ncol, int gridsize, double *Ode_coarse) {
    REF_TRY
    put(Ode_Coarse_Slicer(M(Ode_thin, nrow, ncol), gridsize), Ode_coarse);
    REF_CATCH_VOID
}
// New_Param_List, :1191
int ref_new_param_list(const orc_cfg *c, const double *param, const double *initial, int gridsize, const double *t, int nt,
                       double *Acube, double *Lcube, double *Ode_coarse, double *betaN) {
    REF_TRY
    Pack k(c);
    List r = New_Param_List(V(param, c->nparam_total), V(initial, c->p), gridsize, V(t, nt), k.x_r, k.x_i, k.transP, k.model,
                            k.transX);
    put_FT(as<List>(r[0]), Acube, Lcube);
    put(as<arma::mat>(r[1]), Ode_coarse);
    put(as<arma::vec>(r[2]), betaN);
    return ORC_OK;
    REF_CATCH_STATUS
}
// TransformTraj, :877
void ref_transform_traj(const double *OdeTraj, int nrow, int p, const double *OriginLatent, int k, const double *Acube,
                        const double *Lcube, double *LNA_traj) {
    REF_TRY
    put(TransformTraj(M(OdeTraj, nrow, p + 1), M(OriginLatent, k, p), make_FT(Acube, Lcube, p, k)), LNA_traj);
    REF_CATCH_VOID
}
// Traj_sim_general_noncentral, :824 (normals: p per step, in step order)
void ref_traj_sim_noncentral(const double *OdeTraj, int nrow, int p, const double *Acube, const double *Lcube,
                             double t_correct, const double *normals, double *LNA_traj, double *OriginLatent,
                             double *logMultiNorm, double *logOrigin) {
    REF_TRY
    streams(normals, (long)(nrow - 1) * p, nullptr, 0);
    List r = Traj_sim_general_noncentral(M(OdeTraj, nrow, p + 1), make_FT(Acube, Lcube, p, nrow - 1), t_correct);
    put(as<arma::mat>(r[0]), LNA_traj);
    put(as<arma::mat>(r[1]), OriginLatent);
    *logMultiNorm = as<double>(r[2]);
    *logOrigin = as<double>(r[3]);
    REF_CATCH_VOID
}
// log_like_traj_general2, :675
double ref_log_like_traj_general2(const double *SdeTraj, const double *OdeTraj, int nrow, int p, const double *Acube,
                                  const double *Scube, double t_correct) {
    REF_TRY
    return log_like_traj_general2(M(SdeTraj, nrow, p + 1), M(OdeTraj, nrow, p + 1), make_FT(Acube, Scube, p, nrow - 1), 1,
                                  t_correct);
    REF_CATCH_NAN
}
// volz_loglik_nh2, :1119
double ref_volz_loglik_nh2(const orc_init *init, const double *f1, int nrow, int p, const double *betaN, double t_correct,
                           const int *index, int transX) {
    REF_TRY
    return volz_loglik_nh2(make_init(init), M(f1, nrow, p + 1), V(betaN, nrow), t_correct, IV(index, 2), transx_name(transX));
    REF_CATCH_NAN
}
// volz_loglik_nh3, :1058
double ref_volz_loglik_nh3(const orc_init *init, const double *f1, int nrow, int p, const double *betaN, double t_correct,
                           const int *index, int transX) {
    REF_TRY
    return volz_loglik_nh3(make_init(init), M(f1, nrow, p + 1), V(betaN, nrow), t_correct, IV(index, 2), transx_name(transX));
    REF_CATCH_NAN
}
// coal_loglik3, :1016
double ref_coal_loglik3(const orc_init *init, const double *f1, int nrow, int p, double t_correct, double lambda, int Index,
                        int transX) {
    REF_TRY
    return coal_loglik3(make_init(init), M(f1, nrow, p + 1), t_correct, lambda, Index, transx_name(transX));
    REF_CATCH_NAN
}
// coal_loglik, SIR_phylodyn.cpp:436
double ref_coal_loglik(const orc_init *init, const double *f1, int nrow, int p, double t_correct, double lambda) {
    REF_TRY
    return coal_loglik(make_init(init), M(f1, nrow, p + 1), t_correct, lambda, 1, "standard");
    REF_CATCH_NAN
}

// the body shared by Update_Param (:1217-1226) and the slice loops: New_Param_List -> TransformTraj -> volz_loglik_nh2
int ref_full_loglik(const orc_cfg *c, const orc_init *init, const double *param, const double *initial, const double *t,
                    int nt, int gridsize, const double *OriginTraj, double t_correct, double *Acube, double *Lcube,
                    double *Ode_coarse, double *betaN, double *LatentTraj, double *coalLog) {
    *coalLog = ORC_SENTINEL;
    REF_TRY
    Pack k(c);
    const int kk = (nt - 1) / gridsize;
    List r = New_Param_List(V(param, c->nparam_total), V(initial, c->p), gridsize, V(t, nt), k.x_r, k.x_i, k.transP, k.model,
                            k.transX);
    List ft = as<List>(r[0]);
    arma::mat ode = as<arma::mat>(r[1]);
    arma::vec bN = as<arma::vec>(r[2]);
    arma::mat traj = TransformTraj(ode, M(OriginTraj, kk, c->p), ft);
    *coalLog = volz_loglik_nh2(make_init(init), traj, bN, t_correct, k.x_i.subvec(2, 3), k.transX);
    put_FT(ft, Acube, Lcube);
    put(ode, Ode_coarse);
    put(bN, betaN);
    put(traj, LatentTraj);
    return ORC_OK;
    REF_CATCH_STATUS
}

// Update_Param, :1212 (one uniform)
int ref_update_param(const orc_cfg *c, const orc_init *init, const double *param, const double *initial, const double *t,
                     int nt, const double *OriginTraj, int gridsize, double coal_log, double prior_proposal_offset,
                     double t_correct, double unif, int *accept, double *Acube, double *Lcube, double *Ode_coarse,
                     double *betaN, double *LatentTraj, double *coalLog_new) {
    *accept = 0;
    *coalLog_new = NAN;  // the reference does not report the proposal's likelihood on rejection
    REF_TRY
    Pack k(c);
    streams(nullptr, 0, &unif, 1);
    List r = Update_Param(V(param, c->nparam_total), V(initial, c->p), V(t, nt), M(OriginTraj, (nt - 1) / gridsize, c->p), k.x_r,
                          k.x_i, make_init(init), gridsize, coal_log, prior_proposal_offset, t_correct, k.transP, k.model,
                          k.transX, true);
    if (as<bool>(r["accept"])) {
        *accept = 1;
        put_FT(as<List>(r["FT"]), Acube, Lcube);
        put(as<arma::mat>(r["Ode"]), Ode_coarse);
        put(as<arma::vec>(r["betaN"]), betaN);
        *coalLog_new = as<double>(r["coalLog"]);
        put(as<arma::mat>(r["LatentTraj"]), LatentTraj);
    }
    return ORC_OK;
    REF_CATCH_STATUS
}

// B FULL evaluations through Update_Param itself (coal_log = -DBL_MAX: always accepted), one core
int ref_full_loglik_batch(const orc_cfg *c, const orc_init *init, long B, const double *param, const double *initial,
                          const double *t, int nt, int gridsize, const double *OriginTraj, double t_correct, double *coalLog,
                          int *status) {
    Pack k(c);
    List ini = make_init(init);
    arma::vec tt = V(t, nt);
    const int kk = (nt - 1) / gridsize;
    const double half = 0.5;
    for (long b = 0; b < B; ++b) {
        status[b] = ORC_OK;
        coalLog[b] = ORC_SENTINEL;
        try {
            streams(nullptr, 0, &half, 1);
            List r = Update_Param(V(param + b * c->nparam_total, c->nparam_total), V(initial + b * c->p, c->p), tt,
                                  M(OriginTraj + (size_t)b * kk * c->p, kk, c->p), k.x_r, k.x_i, ini, gridsize, -DBL_MAX, 0.0,
                                  t_correct, k.transP, k.model, k.transX, true);
            coalLog[b] = as<double>(r["coalLog"]);
        } catch (const std::invalid_argument &) {
            status[b] = ORC_CHOL_FAIL;
        } catch (const std::exception &e) {
            g_err = e.what();
            return -1;
        }
    }
    return 0;
}

// Param_Slice_update, :1246
void ref_param_slice_update(const double *param, int nparam_total, const int *x_i, double theta, const double *newChs,
                            double *param_new) {
    REF_TRY
    arma::vec x_r(x_i[0] + 1);
    put(Param_Slice_update(V(param, nparam_total), x_r, IV(x_i, 4), theta, V(newChs, x_i[0]), 1.0), param_new);
    REF_CATCH_VOID
}
// Param_Slice_update_all2, :1268
void ref_param_slice_update_all2(const double *par_old, int npar, const int *x_i, double theta, const double *newChs,
                                 const int *ESS_vec, const double *pr8, double *par_new) {
    REF_TRY
    arma::vec x_r(x_i[0] + 1);
    put(Param_Slice_update_all2(V(par_old, npar), x_r, IV(x_i, 4), theta, V(newChs, npar - 1), IV(ESS_vec, 7), make_priors(pr8)),
        par_new);
    REF_CATCH_VOID
}

// ESlice_par_General, :1532.  n_evals is not observable from outside the reference function: -1.
int ref_eslice_par_general(const orc_cfg *c, const orc_init *init, const double *par_old, int npar, const double *t, int nt,
                           const double *OriginTraj, const double *pr8, int gridsize, const int *ESS_vec, double coal_log,
                           double t_correct, const double *normals, const double *uniforms, int *n_norm, int *n_unif,
                           int *n_evals, double *par_new, double *Acube, double *Lcube, double *Ode_coarse, double *betaN,
                           double *LatentTraj, double *OriginTraj_new, double *CoalLog, double *logOrigin) {
    *n_evals = -1;
    REF_TRY
    Pack k(c);
    const int kk = (nt - 1) / gridsize;
    streams(normals, (long)(npar - 1) + (long)kk * c->p, uniforms, 22);
    List r = ESlice_par_General(V(par_old, npar), V(t, nt), M(OriginTraj, kk, c->p), make_priors(pr8), k.x_r, k.x_i,
                                make_init(init), gridsize, IV(ESS_vec, 7), coal_log, t_correct, k.transP, k.model, k.transX,
                                true);
    *n_norm = (int)shim_rng().used_norm;
    *n_unif = (int)shim_rng().used_unif;
    put(as<arma::vec>(r["betaN"]), betaN);
    put_FT(as<List>(r["FT"]), Acube, Lcube);
    put(as<arma::mat>(r["OdeTraj"]), Ode_coarse);
    put(as<arma::vec>(r["par"]), par_new);
    put(as<arma::mat>(r["LatentTraj"]), LatentTraj);
    *CoalLog = as<double>(r["CoalLog"]);
    put(as<arma::mat>(r["OriginTraj"]), OriginTraj_new);
    *logOrigin = as<double>(r["logOrigin"]);
    return ORC_OK;
    REF_CATCH_STATUS
}

// ESlice_general_NC, :1690
int ref_eslice_general_nc(const orc_init *init, const double *f_cur, int k, int p, const double *OdeTraj, const double *Acube,
                          const double *Lcube, const double *betaN, double t_correct, double lambda, double coal_log, int volz,
                          const int *Index, int transX, const double *normals, const double *uniforms, int *n_norm,
                          int *n_unif, int *n_evals, double *LatentTraj, double *OriginTraj_new, double *logOrigin,
                          double *CoalLog) {
    *n_evals = -1;
    REF_TRY
    streams(normals, (long)k * p, uniforms, 23);
    arma::vec state(p);
    List r = ESlice_general_NC(M(f_cur, k, p), M(OdeTraj, k + 1, p + 1), make_FT(Acube, Lcube, p, k), state, make_init(init),
                               V(betaN, k + 1), t_correct, lambda, coal_log, 1, volz != 0, Index[1] == 2 ? "SEIR" : "SIR",
                               transx_name(transX));
    *n_norm = (int)shim_rng().used_norm;
    *n_unif = (int)shim_rng().used_unif;
    put(as<arma::mat>(r["LatentTraj"]), LatentTraj);
    put(as<arma::mat>(r["OriginTraj"]), OriginTraj_new);
    *logOrigin = as<double>(r["logOrigin"]);
    *CoalLog = as<double>(r["CoalLog"]);
    return ORC_OK;
    REF_CATCH_STATUS
}

// ESlice_change_points, :1317
int ref_eslice_change_points(const orc_cfg *c, const orc_init *init, const double *param, const double *initial,
                             const double *t, int nt, const double *OriginTraj, int gridsize, double coal_log,
                             double t_correct, const double *normals, const double *uniforms, int *n_norm, int *n_unif,
                             int *n_evals, double *param_new, double *Acube, double *Lcube, double *Ode_coarse, double *betaN,
                             double *LatentTraj, double *CoalLog, double *LogChProb) {
    *n_evals = -1;
    REF_TRY
    Pack k(c);
    streams(normals, c->x_i[0], uniforms, 23);
    List r = ESlice_change_points(V(param, c->nparam_total), V(initial, c->p), V(t, nt), M(OriginTraj, (nt - 1) / gridsize, c->p),
                                  k.x_r, k.x_i, make_init(init), gridsize, coal_log, t_correct, k.transP, k.model, k.transX,
                                  true);
    *n_norm = (int)shim_rng().used_norm;
    *n_unif = (int)shim_rng().used_unif;
    put(as<arma::vec>(r["betaN"]), betaN);
    put_FT(as<List>(r["FT"]), Acube, Lcube);
    put(as<arma::mat>(r["OdeTraj"]), Ode_coarse);
    put(as<arma::vec>(r["param"]), param_new);
    put(as<arma::mat>(r["LatentTraj"]), LatentTraj);
    *CoalLog = as<double>(r["CoalLog"]);
    *LogChProb = as<double>(r["LogChProb"]);
    return ORC_OK;
    REF_CATCH_STATUS
}

// Traj_sim_general3, :771 (normals: p per step)
int ref_traj_sim_general3(const double *OdeTraj, int nrow, int p, const double *Acube, const double *Scube, double t_correct,
                          const double *normals, double *LNA_traj, double *loglike) {
    REF_TRY
    streams(normals, (long)(nrow - 1) * p, nullptr, 0);
    List r = Traj_sim_general3(M(OdeTraj, nrow, p + 1), make_FT(Acube, Scube, p, nrow - 1), t_correct);
    put(as<arma::mat>(r[0]), LNA_traj);
    *loglike = as<double>(r[1]);
    return ORC_OK;
    REF_CATCH_STATUS
}
// Traj_sim_ezG2, :907
int ref_traj_sim_ezG2(const orc_cfg *c, const double *initial, const double *t, int nt, const double *param, int gridsize,
                      double t_correct, const double *normals, double *LNA_traj, double *loglike) {
    REF_TRY
    Pack k(c);
    streams(normals, (long)(nt / gridsize) * c->p, nullptr, 0);
    List r = Traj_sim_ezG2(V(initial, c->p), V(t, nt), V(param, c->nparam_total), gridsize, k.x_r, k.x_i, t_correct, k.transP,
                           k.model, k.transX);
    put(as<arma::mat>(r[0]), LNA_traj);
    *loglike = as<double>(r[1]);
    return ORC_OK;
    REF_CATCH_STATUS
}
// ESlice_general2, :1783
int ref_eslice_general2(const orc_init *init, const double *f_cur, int nrow, int p, const double *OdeTraj, const double *Acube,
                        const double *Scube, const double *betaN, double t_correct, const int *Index, int transX,
                        const double *normals, const double *uniforms, int *n_unif, int *n_evals, double *newTraj) {
    *n_evals = -1;
    REF_TRY
    streams(normals, (long)(nrow - 1) * p, uniforms, 23);
    arma::vec state(p);
    arma::mat r = ESlice_general2(M(f_cur, nrow, p + 1), M(OdeTraj, nrow, p + 1), make_FT(Acube, Scube, p, nrow - 1), state,
                                  make_init(init), V(betaN, nrow), t_correct, 10.0, 1, 1, true, Index[1] == 2 ? "SEIR" : "SIR",
                                  transx_name(transX));
    *n_unif = (int)shim_rng().used_unif;
    put(r, newTraj);
    return ORC_OK;
    REF_CATCH_STATUS
}

// ODE2, SIRS_period.cpp:174
void ref_sirs_ode2(const double *initial, const double *t, int n, const double *param, double period, double *OdeTraj) {
    REF_TRY
    put(ODE2(V(initial, 3), V(t, n), V(param, 4), period), OdeTraj);
    REF_CATCH_VOID
}
// SIRS_KOM_Filter, SIRS_period.cpp:148
void ref_sirs_kom_filter(const double *OdeTraj, int nrow, const double *param, int gridsize, double period, double *Acube,
                         double *Scube) {
    REF_TRY
    put_FT(SIRS_KOM_Filter(M(OdeTraj, nrow, 4), V(param, 4), gridsize, period), Acube, Scube);
    REF_CATCH_VOID
}
// mode 0: log_like_trajSIRS, SIRS_period.cpp:82; mode 1: Traj_sim_SIRS, :197 (normals: 3 per step)
int ref_sirs_centred(int mode, const double *X_or_normals, const double *initial, const double *OdeTraj, int nrow,
                     const double *Acube, const double *Scube, double t_correct, double *traj_out, double *loglik) {
    REF_TRY
    List ft = make_FT(Acube, Scube, 3, nrow - 1);
    if (mode == 0) {
        *loglik = log_like_trajSIRS(M(X_or_normals, nrow, 4), M(OdeTraj, nrow, 4), ft, 1, t_correct);
    } else {
        streams(X_or_normals, (long)(nrow - 1) * 3, nullptr, 0);
        List r = Traj_sim_SIRS(V(initial, 3), M(OdeTraj, nrow, 4), ft, t_correct);
        put(as<arma::mat>(r[0]), traj_out);
        *loglik = as<double>(r[1]);
    }
    return ORC_OK;
    REF_CATCH_STATUS
}

#ifndef LNA_DROPIN  // basic_util.cpp stays the reference's own file in a drop-in build
// basic_util.cpp:5 / :16
void ref_inv2(const double *a, double *res) {
    REF_TRY
    put(inv2(M(a, 2, 2)), res);
    REF_CATCH_VOID
}
void ref_chols(const double *S, double *L) {
    REF_TRY
    put(chols(M(S, 2, 2)), L);
    REF_CATCH_VOID
}
#endif

}  // extern "C"
