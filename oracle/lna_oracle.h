/*
 * lna_oracle.h -- CPU ORACLE for the LNA + coalescent likelihood / elliptical-slice hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a scalar, single-threaded, dependency-free C restatement of
 * the reference algorithm (MingweiWilliamTang/LNAphyloDyn, src/LNA_functional.cpp and friends).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call it, and only as the checker / the timed CPU baseline -- never from the product path
 * (lnaphylodyn_b200/), which must fail loudly when its CUDA library is missing.
 *
 * Parity pinning, two independent anchors:
 *  (1) tests/test_oracle_golden.py checks this oracle against the golden vectors the reference ships in
 *      its knitr caches (tests/golden/{changepoint_sim,constant_sim}.npz, made by tests/golden/make_golden.py):
 *      ODE trajectory, FT$A, FT$Sigma(=L), LatentTraj, betaN, coalLog of the final chain state and
 *      2x104 per-iteration (par, Trajectory)->coalLog triplets.
 *  (2) tests/test_ref_vs_oracle.py checks it, function by function, against THE REFERENCE'S OWN SOURCES compiled
 *      verbatim (oracle/_ref/liblna_ref.so = /root/reference/src/LNA_functional.cpp, basic_util.cpp, SIR_LNA.cpp,
 *      SIR_log_LNA.cpp, SIRS_period.cpp, SIR_phylodyn.cpp against the stand-in Armadillo/Rcpp headers of
 *      oracle/shim/; recipe `make -C oracle ref`; harness oracle/ref_harness.cpp).  That covers what no cache
 *      covers: Update_Param decisions and the shrink sequences of ESlice_par_General / ESlice_general_NC /
 *      ESlice_change_points and of whole driver iterations, on injected streams.  The same library produced
 *      tests/golden/ref_as_written.npz, which the GPU tests use where /root/reference does not exist.
 * The random GENERATORS themselves (arma::randn over R's RNG, R::runif) are third-party and unpinned: every
 * sampler takes PRE-DRAWN normals / uniforms and consumes them in the reference's order (SURVEY.md Appendix C;
 * the consumption order IS pinned by (2)).
 *
 * The reference's R package build (R CMD INSTALL: R + Rcpp + RcppArmadillo + LAPACK) cannot run here; only its
 * C++ sources are compiled, as they are, on the stand-in headers.
 *
 * Conventions: all matrices are COLUMN-MAJOR doubles exactly as Armadillo lays them out;
 * trajectories are nrow x (p+1) with column 0 = time; cubes are p x p x k (slice-major).
 * All `file:line` citations are into /root/reference/.
 */
#ifndef LNA_ORACLE_H_
#define LNA_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SIR = 0, ORC_SEIR = 1, ORC_SEIR2 = 2 };
enum { ORC_STANDARD = 0, ORC_LOG = 1 };

/* Replaces the reference's (model, transX, x_r, x_i) argument packs (LNA_functional.cpp:33-35). */
typedef struct {
    int model;          /* ORC_SIR / ORC_SEIR / ORC_SEIR2                      */
    int transX;         /* ORC_STANDARD / ORC_LOG                              */
    int p;              /* number of compartments (SIR 2, SEIR 3, SEIR2 2)     */
    int nparam_total;   /* length of `param` = nparam + nch + 1 (hyper)        */
    const double *x_r;  /* (N, cht_1 .. cht_nch)                               */
    int x_i[4];         /* (nch, nparam, index R0, index gamma)                */
} orc_cfg;

/* Positional view of coal_lik_init()'s list (R/coal_simulation.R:204-277). */
typedef struct {
    int E;                  /* number of intervals = length(C) = sum(gridrep)  */
    int ng;                 /* Init$ng                                          */
    const double *C, *D, *y; /* length E                                        */
    const double *gridrep;  /* length ng (stored as double in R)               */
} orc_init;

#define ORC_SENTINEL (-10000000.0)

/* status bits (shared with the C ABI of the product) */
enum { ORC_OK = 0, ORC_CHOL_FAIL = 1, ORC_NEG_STATE = 2, ORC_NAN = 4, ORC_CAP_HIT = 8 };

void orc_betaTs(const double *param, const double *times, int m, const double *x_r, const int *x_i,
                double *betas);
void orc_param_transform(double t, const double *param, int nparam_total, const double *x_r,
                         const int *x_i, double *param2);
void orc_ode_one(const orc_cfg *c, const double *states, const double *param, double t, double *res);
void orc_F(const orc_cfg *c, const double *states, const double *thetas, double *F /* p x p */);
int  orc_h(const orc_cfg *c, const double *states, const double *thetas, double *h);
void orc_ode_rk45(const orc_cfg *c, const double *initial, const double *t, int n, const double *param,
                  double *OdeTraj /* n x (p+1) */);
double orc_ode_rk45_stop(const orc_cfg *c, const double *initial, const double *t, int n,
                         const double *param, double tol);
void orc_sigmaF(const orc_cfg *c, const double *Traj_par, int nrow, int ld, const double *param,
                double *expF, double *Sigma);
void orc_kf_param(const orc_cfg *c, const double *OdeTraj, int nrow, const double *param, int gridsize,
                  double *Acube, double *Scube);
int  orc_kf_param_chol(const orc_cfg *c, const double *OdeTraj, int nrow, const double *param, int gridsize,
                       double *Acube, double *Lcube);
int  orc_chol_lower(const double *S, int p, double jitter, double *L);
void orc_ode_coarse_slicer(const double *Ode_thin, int nrow, int ncol, int gridsize, double *Ode_coarse);
int  orc_new_param_list(const orc_cfg *c, const double *param, const double *initial, int gridsize,
                        const double *t, int nt, double *Acube, double *Lcube, double *Ode_coarse,
                        double *betaN);
void orc_transform_traj(const double *OdeTraj, int nrow, int p, const double *OriginLatent, int k,
                        const double *Acube, const double *Lcube, double *LNA_traj);
void orc_traj_sim_noncentral(const double *OdeTraj, int nrow, int p, const double *Acube,
                             const double *Lcube, double t_correct, const double *normals,
                             double *LNA_traj, double *OriginLatent, double *logMultiNorm,
                             double *logOrigin);
double orc_log_like_traj_general2(const double *SdeTraj, const double *OdeTraj, int nrow, int p,
                                  const double *Acube, const double *Scube, double t_correct);
double orc_volz_loglik_nh2(const orc_init *init, const double *f1, int nrow, int p, const double *betaN,
                           double t_correct, const int *index, int transX);
double orc_volz_loglik_nh3(const orc_init *init, const double *f1, int nrow, int p, const double *betaN,
                           double t_correct, const int *index, int transX);
double orc_coal_loglik3(const orc_init *init, const double *f1, int nrow, int p, double t_correct,
                        double lambda, int Index, int transX);
double orc_coal_loglik(const orc_init *init, const double *f1, int nrow, int p, double t_correct,
                       double lambda);

/* One FULL evaluation: New_Param_List -> TransformTraj -> volz_loglik_nh2 (the body shared by
 * Update_Param :1217-1226 and the ESlice loops).  Returns status bits; *coalLog is set. */
int orc_full_loglik(const orc_cfg *c, const orc_init *init, const double *param, const double *initial,
                    const double *t, int nt, int gridsize, const double *OriginTraj, double t_correct,
                    double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj,
                    double *coalLog);

int orc_update_param(const orc_cfg *c, const orc_init *init, const double *param, const double *initial,
                     const double *t, int nt, const double *OriginTraj, int gridsize, double coal_log,
                     double prior_proposal_offset, double t_correct, double unif, int *accept,
                     double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj,
                     double *coalLog_new);

void orc_param_slice_update(const double *param, int nparam_total, const int *x_i, double theta,
                            const double *newChs, double *param_new);
void orc_param_slice_update_all2(const double *par_old, int npar, const int *x_i, double theta,
                                 const double *newChs, const int *ESS_vec, const double *pr8,
                                 double *par_new);

/* Elliptical-slice loops.  `normals` / `uniforms` are pre-drawn streams, consumed in the
 * reference's order; *n_norm / *n_unif return how many were consumed, *n_evals the number of
 * likelihood evaluations, *status the OR of ORC_* bits (ORC_CAP_HIT when the 20-shrink cap hit). */
int orc_eslice_par_general(const orc_cfg *c, const orc_init *init, const double *par_old, int npar,
                           const double *t, int nt, const double *OriginTraj, const double *pr8,
                           int gridsize, const int *ESS_vec, double coal_log, double t_correct,
                           const double *normals, const double *uniforms, int *n_norm, int *n_unif,
                           int *n_evals, double *par_new, double *Acube, double *Lcube,
                           double *Ode_coarse, double *betaN, double *LatentTraj,
                           double *OriginTraj_new, double *CoalLog, double *logOrigin);

int orc_eslice_general_nc(const orc_init *init, const double *f_cur, int k, int p, const double *OdeTraj,
                          const double *Acube, const double *Lcube, const double *betaN, double t_correct,
                          double lambda, double coal_log, int volz, const int *Index, int transX,
                          const double *normals, const double *uniforms, int *n_norm, int *n_unif,
                          int *n_evals, double *LatentTraj, double *OriginTraj_new, double *logOrigin,
                          double *CoalLog);

int orc_eslice_change_points(const orc_cfg *c, const orc_init *init, const double *param,
                             const double *initial, const double *t, int nt, const double *OriginTraj,
                             int gridsize, double coal_log, double t_correct, const double *normals,
                             const double *uniforms, int *n_norm, int *n_unif, int *n_evals,
                             double *param_new, double *Acube, double *Lcube, double *Ode_coarse,
                             double *betaN, double *LatentTraj, double *CoalLog, double *LogChProb);

/* centred parametrisation */
int orc_traj_sim_general3(const double *OdeTraj, int nrow, int p, const double *Acube, const double *Scube,
                          double t_correct, const double *normals, double *LNA_traj, double *loglike);
int orc_traj_sim_ezG2(const orc_cfg *c, const double *initial, const double *t, int nt, const double *param,
                      int gridsize, double t_correct, const double *normals, double *LNA_traj, double *loglike);
int orc_eslice_general2(const orc_init *init, const double *f_cur, int nrow, int p, const double *OdeTraj,
                        const double *Acube, const double *Scube, const double *betaN, double t_correct,
                        const int *Index, int transX, const double *normals, const double *uniforms, int *n_unif,
                        int *n_evals, double *newTraj);

/* seasonal SIRS model variant (src/SIRS_period.cpp): ODE2, SIRS_KOM_Filter, log_like_trajSIRS / Traj_sim_SIRS */
void orc_sirs_ode2(const double *initial, const double *t, int n, const double *param, double period,
                   double *OdeTraj);
void orc_sirs_kom_filter(const double *OdeTraj, int nrow, const double *param, int gridsize, double period,
                         double *Acube, double *Scube);
int orc_sirs_centred(int mode, const double *X_or_normals, const double *initial, const double *OdeTraj, int nrow,
                     const double *Acube, const double *Scube, double t_correct, double *traj_out, double *loglik);

/* legacy copies: src/LNA_NS_general.cpp, class Foo of src/generalLNA.cpp, ESlice_SIRS; x_i2 = (nch, nparam) */
void orc_general_betafs(const double *ts, int m, const double *param, const double *x_r, const int *x_i2, double *betas);
void orc_general_point(int mode, const double *states, const double *param, int nparam_total, double t, const double *x_r,
                       const int *x_i2, double *out);
void orc_general_ode_rk45(const double *initial, const double *t, int n, const double *param, int nparam_total,
                          const double *x_r, const int *x_i2, double *OdeTraj);
void orc_general_intsigmaf(const double *Traj_par, int nrow, const double *param, int nparam_total, const double *x_r,
                           const int *x_i2, double *expF, double *Sigma);
void orc_general_kom_filter(const double *OdeTraj, int nrow, const double *param, int nparam_total, int gridsize,
                            const double *x_r, const int *x_i2, double *Acube, double *Scube);
int orc_legacy_centred(int mode, const double *X_or_normals, const double *OdeTraj, int nrow, const double *Acube,
                       const double *Scube, double t_correct, double *traj_out, double *loglik);
int orc_traj_sim_ezG(const double *initial, const double *t, int nt, const double *param, int nparam_total, int gridsize,
                     const double *x_r, const int *x_i2, double t_correct, const double *normals, double *LNA_traj,
                     double *loglike);
double orc_volz_loglik_nh(const orc_init *init, const double *f1, int nrow, const double *betaN, double t_correct);
int orc_eslice_legacy(const orc_init *init, int kind, const double *f_cur, int nrow, int p, const double *OdeTraj,
                      const double *Acube, const double *Scube, const double *state, const double *betaN_or_params4,
                      double t_correct, double lambda, int reps, int volz, const double *normals, const double *uniforms,
                      int *n_unif, int *n_evals, double *newTraj);
void orc_foo_point(int mode, const double *A_pre, const double *A_post, const double *param, double N, const double *states,
                   double *out);
void orc_foo_ode_rk45(const double *A_pre, const double *A_post, const double *param, double N, const double *initial,
                      const double *t, int n, double *OdeTraj);

/* data simulator: volz_thin_sir, R/coal_simulation.R:11-70, on a pre-drawn stream indexed by proposal number */
int orc_volz_thin_sir(const double *t, const double *S, const double *I, const double *betaN, int n, double t_correct,
                      const double *samp_times, const int *n_sampled, int ns, double lower, const double *ua,
                      const double *ub, long n_avail, double *coal_times, int *lineages, int max_coal, int *n_coal,
                      long *n_draws);

/* B independent FULL evaluations on one core (item-major inputs like the product's C ABI). */
int orc_full_loglik_batch(const orc_cfg *c, const orc_init *init, long B, const double *param,
                          const double *initial, const double *t, int nt, int gridsize,
                          const double *OriginTraj, double t_correct, double *coalLog, int *status);

/* basic_util.cpp:23-41 */
void orc_inv2(const double *a, double *res);
void orc_chols(const double *S, double *L);

#ifdef __cplusplus
}
#endif
#endif
