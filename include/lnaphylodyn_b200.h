/*
 * lnaphylodyn_b200.h -- C ABI of the B200-native LNA + coalescent likelihood / elliptical-slice path.
 *
 * This is the drop-in boundary (DESIGN.md section "Boundary"): the reference's Rcpp glue
 * (src/RcppExports.cpp, R/RcppExports.R) would forward its `.Call("_LNAPhyloDyn_<fn>")` entry
 * points for this path to the functions below (INTEGRATION.md shows the shim).  Plain C: pointers,
 * sizes and two small structs; no C++/torch types.  All `file:line` citations are into the
 * reference tree (MingweiWilliamTang/LNAphyloDyn).
 *
 * Conventions
 *  - Every entry point is BATCHED: `B` independent items (chains / proposals / parameter draws).
 *    B = 1 reproduces the single-call Rcpp semantics.  Item b of an array with per-item length n
 *    lives at `ptr + b*n`; inside one item the layout is exactly Armadillo's (column-major
 *    matrices, p x p x k cubes slice-major, trajectories nrow x (p+1) with column 0 = time).
 *  - Functions without suffix take HOST pointers (copies happen inside); `_dev` variants take
 *    DEVICE pointers on the problem's device and are asynchronous on the problem's stream.
 *  - Random numbers are never drawn here: samplers take PRE-DRAWN standard normals / uniforms and
 *    consume them in the reference's order, so chains are replayable.
 *  - Return value: 0 on success, -1 on error (message via lna_last_error()).  Numerical failure is
 *    reported per item in `status` (LNA_ST_* bits) and, like the reference, in-band as -1e7.
 *  - There is NO CPU fallback: without a CUDA device every compute call fails with -1.
 */
#ifndef LNAPHYLODYN_B200_H_
#define LNAPHYLODYN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNA_ABI_VERSION 1

/* model strings of the reference ("SIR","SEIR": LNA_functional.cpp:390-443) become enums */
/* LNA_MODEL_SIRS_PERIOD: the seasonal SIRS of src/SIRS_period.cpp (param = (beta, gamma, mu, alpha), nch = 0,
 * nparam = 4, x_r[0] = period): ODE2, SIRS_KOM_Filter and log_like_trajSIRS only -- the reference has no
 * non-centred / sampler path for it. */
enum { LNA_MODEL_SIR = 0, LNA_MODEL_SEIR = 1, LNA_MODEL_SIRS_PERIOD = 2 };
/* transX: "standard" everywhere; "log" (the reference's legacy log-scale LNA) for model SIR in the
 * integrator / LNA-moment / likelihood entry points (lna_ode_rk45, lna_kf_param, lna_new_param_list, lna_full_loglik,
 * lna_transform_traj, lna_volz_loglik_nh2/3, lna_coal_loglik); the Metropolis / slice operators are standard-scale only,
 * like every driver of the reference that calls them. */
enum { LNA_TRANSX_STANDARD = 0, LNA_TRANSX_LOG = 1 };

/* per-item status bits (reference failure channels) */
enum {
    LNA_ST_OK = 0,
    LNA_ST_CHOL_FAIL = 1, /* arma::chol threw in KF_param_chol (LNA_functional.cpp:660-664)      */
    LNA_ST_NEG_STATE = 2, /* state < 0 in rows 0..ng-1 -> -1e7 (LNA_functional.cpp:1130-1132)     */
    LNA_ST_NAN = 4,       /* NaN among the per-event terms -> -1e7 (LNA_functional.cpp:1172-1174) */
    LNA_ST_CAP_HIT = 8    /* 20-shrink cap -> reverted to the current state (:1609-1622,1741-1747) */
};

#define LNA_SENTINEL (-10000000.0)

/* Replaces (model, transP, transX, x_r, x_i, t, gridsize, t_correct): the arguments every
 * hot-path export repeats (e.g. New_Param_List, LNA_functional.cpp:1191-1193). */
typedef struct lna_cfg {
    int model;          /* LNA_MODEL_*                                                        */
    int transX;         /* LNA_TRANSX_STANDARD | LNA_TRANSX_LOG                               */
    int nch;            /* x_i[0]: number of R0 change points                                 */
    int nparam;         /* x_i[1]: number of model parameters before the change ratios        */
    int iR0;            /* x_i[2]: index of R0 in `param`; also the S column for volz         */
    int iGamma;         /* x_i[3]: index of gamma in `param`; also the I column for volz      */
    int gridsize;       /* fine steps per LNA interval                                        */
    int nt;             /* length(times)                                                      */
    const double *times;/* fine time grid, length nt (uniform in every reference config)      */
    const double *x_r;  /* (N, cht_1..cht_nch), length nch+1                                  */
    double t_correct;   /* forward time of the most recent sample                             */
} lna_cfg;

/* Positional view of coal_lik_init()'s list (R/coal_simulation.R:204-277): the entries the
 * coalescent likelihoods read (init[2],[3],[4],[6],[9]; LNA_functional.cpp:1124,1146,1169). */
typedef struct lna_init {
    int E;                 /* length(C) = length(D) = length(y) = sum(gridrep)             */
    int ng;                /* Init$ng                                                         */
    const double *C;       /* Init$C  = l(l-1)/2 per interval                                */
    const double *D;       /* Init$D  = interval lengths                                     */
    const double *y;       /* Init$y  = 1 if the interval ends in a coalescence              */
    const double *gridrep; /* Init$gridrep (R numeric), length ng                            */
} lna_init;

typedef struct lna_problem lna_problem; /* opaque: device copy of cfg + genealogy, stream, scratch */

/* ---- runtime ---------------------------------------------------------------------------------- */
int lna_abi_version(void);
const char *lna_last_error(void);
int lna_device_count(void);
/* contiguous shard [*begin,*end) of B items owned by `rank` of `nranks` */
int lna_shard(int64_t B, int rank, int nranks, int64_t *begin, int64_t *end);

/* Build the per-genealogy device state.  `init` may be NULL for ODE/LNA-only use.  The per-cell
 * sums  sum_{e in cell} C_e D_e  and  sum y_e  are reduced on the device here (segmented reduction
 * over the sorted events) so that each likelihood evaluation is O(ng) instead of O(E). */
lna_problem *lna_problem_create(const lna_cfg *cfg, const lna_init *init, int device);
void lna_problem_destroy(lna_problem *pb);
int lna_problem_sync(lna_problem *pb);               /* cudaStreamSynchronize on the problem stream */
void *lna_problem_stream(lna_problem *pb);           /* cudaStream_t */
int lna_problem_set_stream(lna_problem *pb, void *cuda_stream);
/* dimensions derived from cfg: p, k = (nt-1)/gridsize, nrow = nt/gridsize + 1, npt = nparam+nch+1 */
int lna_problem_dims(const lna_problem *pb, int *p, int *k, int *nrow, int *npt);
/* number of kernel launches issued through this problem so far (bench.py "gpu_launches") */
int64_t lna_problem_launch_count(const lna_problem *pb);
/* per-cell sums computed at create time (host copies), each of length ng-1 */
int lna_problem_cell_sums(const lna_problem *pb, double *cellCD, double *cellY, double *cellCnt);

/* ---- helpers of the integrator --------------------------------------------------------------- */
/* betaTs (LNA_functional.cpp:31-58): beta at `m` sorted times, per item.  betas: B x m. */
int lna_betaTs(lna_problem *pb, int64_t B, const double *param, const double *times, int m, double *betas);
/* param_transform (:63-86): theta at time t[b] per item; param2: B x npt. */
int lna_param_transform(lna_problem *pb, int64_t B, const double *t, const double *param, double *param2);
/* ODE_rk45 (:455-491): fine mean trajectory, OdeTraj: B x (nt x (p+1)). */
int lna_ode_rk45(lna_problem *pb, int64_t B, const double *initial, const double *param, double *OdeTraj);
/* KF_param_chol / KF_param (:638-670 / :603-634) from a given fine trajectory.
 * Acube, Scube: B x (p*p*k).  chol != 0: Scube holds L = chol(Sigma+1e-8 I)^T and status gets
 * LNA_ST_CHOL_FAIL where the reference would throw; chol == 0: Scube holds Sigma. */
int lna_kf_param(lna_problem *pb, int64_t B, const double *OdeTraj, const double *param, int chol,
                 double *Acube, double *Scube, int32_t *status);
/* Ode_Coarse_Slicer (:1179-1188). Ode_coarse: B x (nrow x (p+1)). */
int lna_ode_coarse_slicer(lna_problem *pb, int64_t B, const double *OdeTraj, double *Ode_coarse);
/* lna_full_loglik (likelihood only) without the synchronisation at the end: the outputs are valid after lna_problem_sync(),
 * and every host buffer passed in must stay alive and unmodified until then (pinned memory, or the copies serialise).
 * Consecutive calls alternate between two device staging sets, so the host-to-device copies of call n+1 overlap the last
 * kernel of call n -- what a caller streaming many batches through host memory wants.  No reference counterpart. */
int lna_full_loglik_async(lna_problem *pb, int64_t B, const double *param, const double *initial, const double *OriginTraj,
                          double *coalLog, int32_t *status);

/* New_Param_List (:1191-1208): ODE -> FT (A, L) -> coarse ODE -> betaN, fused in one kernel. */
int lna_new_param_list(lna_problem *pb, int64_t B, const double *param, const double *initial,
                       double *Acube, double *Lcube, double *Ode_coarse, double *betaN, int32_t *status);

/* ---- non-centred trajectory ------------------------------------------------------------------ */
/* TransformTraj (:877-900).  OriginLatent: B x (k x p); LNA_traj: B x (nrow x (p+1)). */
int lna_transform_traj(lna_problem *pb, int64_t B, const double *Ode_coarse, const double *OriginLatent,
                       const double *Acube, const double *Lcube, double *LNA_traj);
/* Traj_sim_general_noncentral (:824-872) with supplied normals (B x k*p, step-major: step i uses
 * normals[i*p..i*p+p-1]).  Also returns OriginLatent (k x p col-major), logMultiNorm, logOrigin. */
int lna_traj_sim_noncentral(lna_problem *pb, int64_t B, const double *Ode_coarse, const double *Acube,
                            const double *Lcube, const double *normals, double *LNA_traj,
                            double *OriginLatent, double *logMultiNorm, double *logOrigin);

/* ---- coalescent likelihoods ------------------------------------------------------------------ */
/* volz_loglik_nh2 (:1119-1176) on given trajectories (B x nrow x (p+1)) and betaN (B x nrow). */
int lna_volz_loglik_nh2(lna_problem *pb, int64_t B, const double *traj, const double *betaN,
                        double *loglik, int32_t *status);
/* volz_loglik_nh3 (:1058-1113): the same without the NaN guard (a NaN sum is returned as is). */
int lna_volz_loglik_nh3(lna_problem *pb, int64_t B, const double *traj, const double *betaN,
                        double *loglik, int32_t *status);
/* coal_loglik3 (:1016-1054) / coal_loglik (SIR_phylodyn.cpp:436-468): Kingman with scale lambda[b].
 * take_log != 0: column `col` (1-based state column of traj) holds counts (coal_loglik3, transX
 * "standard"); take_log == 0: it already holds log-counts (coal_loglik uses col = 2). */
int lna_coal_loglik(lna_problem *pb, int64_t B, const double *traj, const double *lambda, int col,
                    int take_log, double *loglik);

/* ---- the hot path: one FULL evaluation per item ---------------------------------------------- */
/* (param, initial, OriginTraj) -> coalLog = volz_loglik_nh2(TransformTraj(New_Param_List(...)))
 * i.e. the body of Update_Param (:1217-1226) and of every ESlice iteration, fused in one kernel,
 * one thread per item, all state in registers.  Optional outputs may be NULL. */
int lna_full_loglik(lna_problem *pb, int64_t B, const double *param, const double *initial,
                    const double *OriginTraj, double *coalLog, int32_t *status,
                    double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj);
int lna_full_loglik_dev(lna_problem *pb, int64_t B, const double *param, const double *initial,
                        const double *OriginTraj, double *coalLog, int32_t *status,
                        double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj);

/* Chain-shared form of the hot call: the `proposals` consecutive evaluations of a chain -- the body of the slice loop,
 * ESlice_par_General :1588-1640, where every candidate of a chain is evaluated on the chain's OriginTraj and, unless
 * ESS_vec[1] is set, its initial state -- read ONE initial state and ONE OriginTraj per chain.  param: (n_chains *
 * proposals) x npt; initial: n_chains x p; OriginTraj: n_chains x (k x p); coalLog / status: n_chains * proposals.
 * Same kernel and same bits as lna_full_loglik on the expanded inputs; a step ships 336 + (16 + 640) / proposals bytes per
 * evaluation instead of 992 (SIR, nch = 39).  `_dev`: device pointers, asynchronous on the problem's stream; `_async`: host
 * pointers, results valid after lna_problem_sync() (copies of call n+1 overlap the kernel of call n). */
int lna_full_loglik_shared(lna_problem *pb, int64_t n_chains, int proposals, const double *param, const double *initial,
                           const double *OriginTraj, double *coalLog, int32_t *status);
int lna_full_loglik_shared_dev(lna_problem *pb, int64_t n_chains, int proposals, const double *param, const double *initial,
                               const double *OriginTraj, double *coalLog, int32_t *status);
int lna_full_loglik_shared_async(lna_problem *pb, int64_t n_chains, int proposals, const double *param,
                                 const double *initial, const double *OriginTraj, double *coalLog, int32_t *status);

/* Seasonal SIRS (src/SIRS_period.cpp), fused: per draw
 *     log_like_trajSIRS(SdeTraj, coarse rows of ODE2(initial, t, param, period), SIRS_KOM_Filter(ODE2(...), param, gridsize,
 *                       period), gridsize, t_correct)                                   (:174-193, :147-168, :50-79, :82-118)
 * in ONE pass per draw, registers only -- the likelihood sweep over parameter draws of BASELINE configs[3].  Problem created
 * with LNA_MODEL_SIRS_PERIOD (x_r[0] = period).  param: B x 4 (beta, gamma, mu, alpha); initial: B x 3, or 1 x 3 with
 * initial_shared; SdeTraj: nrow x 4 column-major per item -- one per draw (sde_group = 1), one per sde_group consecutive draws,
 * or ONE for the whole sweep (sde_group = 0); NULL = no density (then loglik must be NULL).  Optional outputs: Acube, Scube
 * (B x 3*3*k: SIRS_KOM_Filter's A and Sigma), Ode_coarse (B x nrow x 4).  status: LNA_ST_NAN where the density is NaN. */
int lna_sirs_loglik(lna_problem *pb, int64_t B, const double *param, const double *initial, int initial_shared,
                    const double *SdeTraj, int sde_group, double *loglik, int32_t *status, double *Acube, double *Scube,
                    double *Ode_coarse);
int lna_sirs_loglik_dev(lna_problem *pb, int64_t B, const double *param, const double *initial, int initial_shared,
                        const double *SdeTraj, int sde_group, double *loglik, int32_t *status, double *Acube, double *Scube,
                        double *Ode_coarse);

/* Update_Param (:1212-1243): FULL evaluation + MH test  log(unif[b]) < coalLog_new - coal_log[b] +
 * prior_proposal_offset[b].  accept[b] in {0,1}; outputs are written for every item (the caller
 * keeps them only where accept[b] == 1, like the R wrapper LNA_chpt_slice.R:254-262). */
int lna_update_param(lna_problem *pb, int64_t B, const double *param, const double *initial,
                     const double *OriginTraj, const double *coal_log, const double *prior_proposal_offset,
                     const double *unif, int32_t *accept, double *coalLog_new, int32_t *status,
                     double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj);

/* ---- elliptical slice sampling --------------------------------------------------------------- */
/* Param_Slice_update (:1246-1253) and Param_Slice_update_all2 (:1268-1314), per item. */
int lna_param_slice_update(lna_problem *pb, int64_t B, const double *param, const double *theta,
                           const double *newChs, double *param_new);
int lna_param_slice_update_all2(lna_problem *pb, int64_t B, const double *par_old, const double *theta,
                                const double *newChs, const int32_t *ESS_vec7, const double *prior8,
                                double *par_new);

/* Number of uniforms an ESlice call may consume: u, theta_1 and one per shrink. */
#define LNA_ESLICE_MAX_UNIF 22

/* ESlice_par_General (:1532-1688), SIR (p = 2) only like the reference.  par_old: B x npar
 * (npar = 2 + npt); normals: B x (npar-1); uniforms: B x LNA_ESLICE_MAX_UNIF (u, theta draws).
 * prior8 = (priorList[[1]][1:2], ..., priorList[[4]][1:2]); ESS_vec7[6] = 1: see lna_eslice_par_general_ex.
 * The candidates theta_1..theta_20 depend only on the uniforms, so they are evaluated
 * speculatively, `spec_width` (1,2,4,8,16,32) per in-kernel wave, and the first one with loglike > logy
 * is taken: identical to the sequential loop.  spec_width = 0 picks a phase plan from the batch size
 * (lanes per chain that fill the machine, with the still-unresolved chains compacted between phases);
 * a negative value -w forces the plan the automatic policy uses when w lanes per chain fill the machine.  n_evals[b] = evaluations the
 * sequential loop would have made (1 + shrinks). */
int lna_eslice_par_general(lna_problem *pb, int64_t B, const double *par_old, const double *OriginTraj,
                           const double *prior8, const int32_t *ESS_vec7, const double *coal_log,
                           const double *normals, const double *uniforms, int spec_width,
                           double *par_new, double *Acube, double *Lcube, double *Ode_coarse,
                           double *betaN, double *LatentTraj, double *CoalLog, double *logOrigin,
                           int32_t *n_evals, int32_t *status);

/* The same with the slot of the reference's result list the call above leaves out: OriginTraj_new (B x k*p; may be NULL).
 * It equals OriginTraj unless ESS_vec7[6] = 1, the self-mixing branch (:1576-1578, 1660-1662): there every candidate whose
 * New_Param_List succeeds replaces OriginTraj_new by OriginTraj cos(theta) + OriginTraj_new sin(theta), a running scalar
 * multiple of OriginTraj; the candidates of a chain are then evaluated one per wave in order (the multiple depends on which
 * earlier candidates failed their factorisation), and the k*p extra normals the reference draws and discards are simply
 * not taken.  The R caller (update_Par_ESlice_combine, R/LNA_chpt_slice.R:456-472) never stores OriginTraj_new, and
 * neither do the resident chains. */
int lna_eslice_par_general_ex(lna_problem *pb, int64_t B, const double *par_old, const double *OriginTraj,
                              const double *prior8, const int32_t *ESS_vec7, const double *coal_log,
                              const double *normals, const double *uniforms, int spec_width,
                              double *par_new, double *Acube, double *Lcube, double *Ode_coarse,
                              double *betaN, double *LatentTraj, double *CoalLog, double *logOrigin,
                              int32_t *n_evals, int32_t *status, double *OriginTraj_new);

/* ESlice_general_NC (:1690-1778), volz likelihood.  f_cur: B x (k x p); normals: B x (k*p)
 * (column-major k x p like arma::randn(k,p)); uniforms: B x LNA_ESLICE_MAX_UNIF.
 * coal_log[b] == -99999999 recomputes the current likelihood like the reference. */
int lna_eslice_general_nc(lna_problem *pb, int64_t B, const double *f_cur, const double *Ode_coarse,
                          const double *Acube, const double *Lcube, const double *betaN,
                          const double *coal_log, const double *normals, const double *uniforms,
                          double *LatentTraj, double *OriginTraj_new, double *logOrigin, double *CoalLog,
                          int32_t *n_evals, int32_t *status);

/* ESlice_general_NC with volz = FALSE (:1717-1721, 1733-1760): the slice loop on the Kingman likelihood
 * coal_loglik3(init, LogTraj(newTraj), t_correct, lambda, Index(1)) -- LogTraj (SIR_LNA.cpp:40-46) logs the trajectory and
 * coal_loglik3 logs it again, so the value is NaN as soon as I < 1, and a NaN ENDS the loop and is returned (status gets
 * LNA_ST_NAN); the 20-shrink cap recomputes with volz_loglik_nh2 like the reference.  lambda: B.  coal_log must be the
 * caller's current likelihood (-99999999 makes the reference treat the increment matrix as a trajectory: refused). */
int lna_eslice_general_nc_kingman(lna_problem *pb, int64_t B, const double *f_cur, const double *Ode_coarse,
                                  const double *Acube, const double *Lcube, const double *betaN,
                                  const double *coal_log, const double *lambda, const double *normals,
                                  const double *uniforms, double *LatentTraj, double *OriginTraj_new, double *logOrigin,
                                  double *CoalLog, int32_t *n_evals, int32_t *status);

/* ESlice_change_points (:1317-1409).  param: B x npt; normals: B x nch; uniforms: B x 22. */
int lna_eslice_change_points(lna_problem *pb, int64_t B, const double *param, const double *initial,
                             const double *OriginTraj, const double *coal_log, const double *normals,
                             const double *uniforms, int spec_width, double *param_new, double *Acube,
                             double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj,
                             double *CoalLog, double *LogChProb, int32_t *n_evals, int32_t *status);

/* ---- centred parametrisation (legacy entry points) ------------------------------ */
/* log_like_traj_general2 (:675-722): Gaussian transition log-density of SdeTraj around OdeTraj given
 * KF_param's (A, Sigma); rows with time <= t_correct. */
int lna_log_like_traj_general2(lna_problem *pb, int64_t B, const double *SdeTraj, const double *OdeTraj,
                               const double *Acube, const double *Scube, double *loglik);
/* log_like_trajSIRS (SIRS_period.cpp:82-118): the same density for the rank-2 covariance of the seasonal SIRS
 * (PseudoInverse :22-48, DegenerateDet3 :17-19). */
int lna_log_like_traj_sirs(lna_problem *pb, int64_t B, const double *SdeTraj, const double *OdeTraj,
                           const double *Acube, const double *Scube, double *loglik);
/* Traj_sim_SIRS (SIRS_period.cpp:197-240): centred simulation of the seasonal SIRS from `initial` (B x 3) with supplied
 * normals (B x k*3, step-major); noise = mvrnormArma2(Sigma) = chol(Sigma + 1e-13 I_3)' z (basic_util.cpp:72-81), density via
 * PseudoInverse / DegenerateDet3.  Sigma has rank 2, so whether its jittered factorisation exists hangs on the last rounding:
 * the outcome is reported per item (status LNA_ST_CHOL_FAIL = the reference's chol() exception; that item's outputs after
 * the failing interval are meaningless). */
int lna_traj_sim_sirs(lna_problem *pb, int64_t B, const double *initial, const double *OdeTraj, const double *Acube,
                      const double *Scube, const double *normals, double *LNA_traj, double *loglike, int32_t *status);
/* ---- legacy copies of the pipeline: src/LNA_NS_general.cpp, class Foo of src/generalLNA.cpp, ESlice_SIRS ----
 * The integrator / moment functions of LNA_NS_general.cpp are arithmetic twins of the current ones on the SIR model
 * (betafs :25-47 = lna_betaTs; General_ODE_rk45 :76-98 = lna_ode_rk45; General_IntSigmaF :103-135 / General_KOM_Filter
 * :139-159 = lna_kf_param(chol = 0)) and are served by those entry points with x_i = (nch, nparam, 0, 1); the functions
 * below are the ones whose arithmetic differs. */
/* log_like_traj_general (LNA_NS_general.cpp:164-200) = Foo::LNA_log_like_traj (generalLNA.cpp:178-214): the Gaussian
 * transition density of a centred trajectory through inv2(Sigma) (basic_util.cpp:5-14); p = 2. */
int lna_log_like_traj_general(lna_problem *pb, int64_t B, const double *SdeTraj, const double *OdeTraj, const double *Acube,
                              const double *Scube, double *loglik);
/* Traj_sim_general (LNA_NS_general.cpp:205-250) = Foo::LNA_Traj_sim (generalLNA.cpp:217-262): centred simulation with
 * supplied normals (B x k*2, step-major), noise = chols(Sigma) z, density through inv2.  Traj_sim_ezG (:257-314) =
 * lna_ode_rk45 + lna_kf_param(chol = 0) + lna_ode_coarse_slicer + this. */
int lna_traj_sim_general(lna_problem *pb, int64_t B, const double *OdeTraj, const double *Acube, const double *Scube,
                         const double *normals, double *LNA_traj, double *loglike);
/* volz_loglik_nh (SIR_phylodyn.cpp:517-560): f1_log = B x nrow x 3 LOG trajectories (LogTraj, SIR_LNA.cpp:40-46). */
int lna_volz_loglik_nh(lna_problem *pb, int64_t B, const double *f1_log, const double *betaN, double *loglik);
/* The centred trajectory slice update of the legacy drivers, `reps` repetitions per call:
 *   kind 0 ESlice_general (LNA_NS_general.cpp:320-394), kind 2 Foo::LNA_ESlice (generalLNA.cpp:296-372; same text): p = 2,
 *          proposal Traj_sim_general, betaN_or_params4 = betaN (B x nrow), at the 20-shrink cap newTraj = f_cur;
 *   kind 1 ESlice_SIRS (SIRS_period.cpp:313-402): p = 3, proposal Traj_sim_SIRS(state, .), betaN_or_params4 = params4
 *          (B x 4: beta, A, period, N; betat = betaDyn(.) * N), at the cap the LAST candidate stays (:377-380); a failed
 *          chol(Sigma + 1e-13 I) of a proposal = the reference's exception: status LNA_ST_CHOL_FAIL, outputs meaningless.
 * volz != 0: volz_loglik_nh(LogTraj(.), betaN); volz == 0: coal_loglik(LogTraj(.), lambda[b]) (SIR_phylodyn.cpp:436-468).
 * normals: reps x B x k*p (repetition-major, each item step-major); uniforms: B x reps*22, consumed contiguously
 * (u, theta_1, one per shrink; per repetition); n_unif returns how many were consumed. */
int lna_eslice_legacy(lna_problem *pb, int64_t B, int kind, int reps, int volz, const double *f_cur, const double *OdeTraj,
                      const double *Acube, const double *Scube, const double *state, const double *betaN_or_params4,
                      const double *lambda, const double *normals, const double *uniforms, double *newTraj,
                      int32_t *n_evals, int32_t *n_unif, int32_t *status);
/* class Foo (generalLNA.cpp:9-372): reaction network with the hard-wired hazards h = (param0 param1 / N S I, param1 I,
 * param2 N) and effect matrix A_post - A_pre (B x 6: 3 reactions x 2 species, column-major).  mode 0 rate_h :76-83 (out B x
 * 3), 1 dh :84-90 (B x 6), 2 ODE_ones :96-98 (B x 2), 3 the product A' dh that F_vec :92-94 forms (B x 4), 4 ODE_rk45
 * :100-122 (B x n*3 column-major trajectories; t = the n grid times, states_or_initial = the initial states); modes 5 / 6 =
 * 2 / 3 with the infection rate given directly in param[0]: ODE_general_one / general_F (LNA_NS_general.cpp:52-72) on the SIR
 * effect matrix with param[0] = betaf(t) (lna_betaTs).
 * (Foo::IntSigmaF / KOM_Filter_MX / LNA_Traj_sim_ez cannot run as written -- F_vec returns a 2 x 2 product through an
 * arma::vec and IntSigmaF multiplies by an empty local `A` (:130, 146) -- and have no entry point.) */
int lna_foo(lna_problem *pb, int64_t B, int mode, int n, const double *A_pre, const double *A_post, const double *param,
            const double *N, const double *states_or_initial, const double *t, double *out);
/* ---- data simulator ----
 * volz_thin_sir (R/coal_simulation.R:11-70): B genealogies' coalescent times by thinning against the Volz rate on SIR
 * trajectories given on a fine grid of n points (t, S, I, betaN: item b at + b * traj_stride; traj_stride 0 = one
 * trajectory for all items), sampling design (samp_times, n_sampled)[ns] shared by all items.  One warp per genealogy, 32
 * consecutive proposals per round; R's rexp / runif are a counter-based stream (Philox4x32-10, keyed by seed, indexed by
 * (item_offset + b, proposal number)), replayable with lna_thin_draws, and the result is bit-identical to the sequential
 * loop consuming that stream.  Outputs per item: coal_times / lineages (B x max_coal, zero-padded), n_coal, n_draws
 * (proposals consumed), status (1 = max_draws [0 = default cap] or max_coal was hit). */
int lna_volz_thin_sir(lna_problem *pb, int64_t B, int n, int64_t traj_stride, const double *t, const double *S, const double *I,
                      const double *betaN, int ns, const double *samp_times, const int32_t *n_sampled, double t_correct,
                      double lower, uint64_t seed, int64_t item_offset, uint64_t max_draws, int max_coal, double *coal_times,
                      int32_t *lineages, int32_t *n_coal, uint64_t *n_draws, int32_t *status);
int lna_thin_draws(lna_problem *pb, uint64_t seed, uint64_t item, uint64_t first, int64_t count, double *ua, double *ub);
/* Traj_sim_general3 (:771-817): centred LNA simulation with supplied normals (B x k*p, step-major);
 * noise = mvrnormArma(Sigma) (basic_util.cpp:44-58).  Traj_sim_ezG2 (:907-975) = lna_ode_rk45 +
 * lna_kf_param(chol = 0) + lna_ode_coarse_slicer + this. */
int lna_traj_sim_general3(lna_problem *pb, int64_t B, const double *OdeTraj, const double *Acube,
                          const double *Scube, const double *normals, double *LNA_traj, double *loglike,
                          int32_t *status);
/* ESlice_general2 (:1783-1857), volz likelihood: centred elliptical slice on the trajectory itself.
 * f_cur, newTraj: B x (nrow x (p+1)); normals: B x k*p (drawn by the inner Traj_sim_general3);
 * uniforms: B x LNA_ESLICE_MAX_UNIF. */
int lna_eslice_general2(lna_problem *pb, int64_t B, const double *f_cur, const double *OdeTraj,
                        const double *Acube, const double *Scube, const double *betaN, const double *normals,
                        const double *uniforms, double *newTraj, int32_t *n_evals, int32_t *status);

/* ---- resident chain batches: the MCMC driver loop on the device -------------------------------- */
/* State of B chains kept in HBM between iterations (par, OriginTraj, FT, coarse ODE, betaN,
 * LatentTraj, coalLog, logOrigin): the batched equivalent of the R list MCMC_obj
 * (R/general_change_point.R:558-561). */
typedef struct lna_chains lna_chains;

/* Settings of General_MCMC_with_ESlice's iteration body (R/LNA_chpt_slice.R:775-847) as used by
 * the Changpoint_sim / Ebola vignettes: MH on log(gamma) + the no-op hyper MH,
 * parameter ESS, trajectory ESS. */
typedef struct lna_mcmc_opts {
    double prior8[8];     /* (pop_pr[1:2], R0_pr, prior[[3]], hyper_pr) as passed to ESlice_par_General */
    double gamma_prior[2];/* MCMC_setting$prior[[priorId]] for the gamma MH step                     */
    double gamma_prop;    /* MCMC_setting$proposal[[proposeId]]: half-width of the uniform RW        */
    int32_t ESS_vec[7];
    int32_t update_traj;  /* 1: run updateTraj_general_NC (i >= burn1); 0: skip (i < burn1)          */
    int32_t hyper_noop_mh;/* 1: also run the no-op hyper MH entry (parIdlist d)         */
    int32_t spec_width;   /* speculative candidates per wave for the parameter ESS (0 = auto)        */
    /* General entry list of Update_Param_each_Norm (R/LNA_chpt_slice.R:169-265), used instead of the two fields above when
     * n_mh > 0 (e.g. the Ebola vignette's parIdlist = (1, 3, nch + 5)): entries run in order, one Update_Param each.
     * mh_kind 0: uniform random walk of half-width mh_prop on par[mh_index] (0 = S0, 1 = I0, 2 = R0, 3 = gamma), on the log
     *            scale if mh_is_log, prior term dnorm(raw; mh_prior) (:211-217);
     *         1: the hyper entry with hyper = F: no proposal, ch <- exp(log(ch) * hyper / hyper), offset 0;
     *         2: the hyper entry with hyper = T: hyper' = hyper * exp(u), u ~ U(-mh_prop, mh_prop), prior term
     *            dlnorm(hyper'; mh_prior) - u - dlnorm(hyper; mh_prior), ch <- exp(log(ch) * hyper / hyper') (:219-241). */
    int32_t n_mh;
    int32_t mh_kind[4], mh_index[4], mh_is_log[4];
    double mh_prior[4][2];
    double mh_prop[4];
    /* ESS_vec2 of General_MCMC_with_ESlice (R/LNA_chpt_slice.R:826-833): when use_ess2 is set, iterations with update_traj run
     * a SECOND update_Par_ESlice_combine with ESS_vec2 INSTEAD of the trajectory slice update; it consumes the first npar-1
     * normals and the 22 uniforms of the trajectory update's slots. */
    int32_t use_ess2;
    int32_t ESS_vec2[7];
} lna_mcmc_opts;

/* uniforms per chain-iteration: gamma RW draw, MH accept, hyper MH accept, ESS-par (22), ESS-traj (22) */
#define LNA_ITER_UNIF (3 + 2 * LNA_ESLICE_MAX_UNIF)
/* with the general entry list (n_mh > 0): (walk, accept) per entry in slots [2e, 2e+1], ESS-par at 8, ESS-traj at 30 */
#define LNA_ITER_UNIF_GENERAL (2 * 4 + 2 * LNA_ESLICE_MAX_UNIF)

lna_chains *lna_chains_create(lna_problem *pb, int64_t B);
void lna_chains_destroy(lna_chains *ch);
/* Initialise from (par, OriginTraj) per chain: New_Param_List + TransformTraj + volz_loglik_nh2
 * (MCMC_initialize_general, R/general_change_point.R:507-536).  Host pointers. */
int lna_chains_init(lna_chains *ch, const double *par, const double *OriginTraj);
/* One MCMC iteration for all chains.  normals: B x (npar-1 + k*p); uniforms: B x LNA_ITER_UNIF.
 * `_dev`: streams already in HBM.  Host variant copies them in first. */
int lna_chains_step(lna_chains *ch, const lna_mcmc_opts *opts, const double *normals, const double *uniforms);
int lna_chains_step_dev(lna_chains *ch, const lna_mcmc_opts *opts, const double *normals, const double *uniforms);
/* Replayable on-device random streams (Philox4x32-10 keyed by seed and global chain id, counter =
 * iteration and slot; normals by Box-Muller).  lna_chains_step_rng draws this iteration's streams in HBM
 * and runs lna_chains_step_dev on them: nothing but the summaries ever crosses PCIe.
 * lna_draw_streams returns the very same streams to the host (B x n_norm, B x n_unif) so that a chain can
 * be replayed through any other implementation draw for draw.  chain_offset = first global chain id of
 * this batch (lna_shard), which makes the streams independent of how chains are sharded over GPUs. */
int lna_chains_step_rng(lna_chains *ch, const lna_mcmc_opts *opts, uint64_t seed, uint64_t iteration,
                        int64_t chain_offset);
int lna_draw_streams(lna_problem *pb, int64_t B, int64_t chain_offset, int n_norm, int n_unif, uint64_t seed,
                     uint64_t iteration, double *normals, double *uniforms);

/* Settings of General_MCMC2's iteration body (R/LNA_chpt_slice.R:576-691; the constant_sim vignette):
 * Update_Param_each entries (random walk on one parameter each, R/LNA_chpt_slice.R:266-365),
 * update_hyper (R/general_change_point.R:281-313), ESlice_change_points, trajectory ESS. */
typedef struct lna_mcmc2_opts {
    int32_t n_mh;              /* number of Update_Param_each entries (<= 4)                        */
    int32_t mh_index[4];       /* 0-based index into par: 1 = I0, 2 = R0, 3 = gamma                 */
    int32_t mh_is_log[4];      /* isLoglist: the walk is on log(par)                                */
    int32_t mh_prior_kind[4];  /* 0: dnorm(mean, sd) on the (log) value; 1: dunif(lo, hi)           */
    double mh_prior[4][2];
    double mh_prop[4];         /* half-width of the uniform random walk                             */
    int32_t update_hyper;      /* updateVec[5] == 0.5 and i >= burn1                                */
    double hyper_prop;         /* proposal$hyper_prop                                               */
    double hyper_prior[2];     /* dgamma(shape, rate)                                               */
    int32_t update_chpt;       /* updateVec[6]                                                      */
    int32_t update_traj;       /* updateVec[7] and i >= burn1                                       */
    int32_t spec_width;        /* 0 = auto                                                          */
    /* stage i >= warmup2 (R/LNA_chpt_slice.R:655-663): the scalar entries and the hyper step are replaced by ONE joint
     * random-walk Metropolis step, update_Param_joint(method = "admcmc") (:9-88), over the listed parameters;
     * the proposal is raw' = raw + T z with the chain's transform set by lna_chains_set_joint_transform */
    int32_t joint;             /* 1: joint step instead of mh_* / update_hyper                       */
    int32_t joint_d;           /* number of jointly proposed parameters (1..4)                       */
    int32_t joint_index[4];    /* 1 = I0, 2 = R0, 3 = gamma, 4 = hyper (last, if present)            */
    int32_t joint_is_log[4];   /* the walk is on log(par)                                            */
    int32_t joint_prior_kind[4]; /* 0 dnorm(mean, sd), 1 dunif(lo, hi), 2 dgamma(shape, rate) on the raw value */
    double joint_prior[4][2];
} lna_mcmc2_opts;

/* Proposal transforms of the joint step: T is B x d x d row-major (host), raw' = raw + T z with z ~ N(0, I_d), i.e.
 * T = V diag(sqrt(max(ev, 0))) of the chain's proposal covariance PCOV = V diag(ev) V' -- what MASS::mvrnorm(1, raw, PCOV)
 * computes (R/LNA_chpt_slice.R:39).  The adaptation of PCOV from the chain's history (:608-630) is host logic
 * (lnaphylodyn_b200/mcmc.py).  With `joint` set an iteration takes d more normals per chain, appended to the stream:
 * normals are B x (nch + k*p + 4), the joint ones at column nch + k*p; its accept test uses uniform 1. */
int lna_chains_set_joint_transform(lna_chains *ch, int d, const double *T);

/* uniforms per General_MCMC2 iteration: 4 MH slots x (walk, accept), hyper (walk, accept), 22 + 22 */
#define LNA_ITER2_UNIF (2 * 4 + 2 + 2 * LNA_ESLICE_MAX_UNIF)

/* One General_MCMC2 iteration.  normals: B x (nch + k*p); uniforms: B x LNA_ITER2_UNIF laid out as
 * [2e, 2e+1] for MH entry e, [8, 9] hyper, [10..31] change-point ESS, [32..53] trajectory ESS. */
int lna_chains_step_mcmc2(lna_chains *ch, const lna_mcmc2_opts *opts, const double *normals, const double *uniforms);
int lna_chains_step_mcmc2_dev(lna_chains *ch, const lna_mcmc2_opts *opts, const double *normals,
                              const double *uniforms);
/* The same on streams drawn on the device: exactly the numbers lna_draw_streams(pb, B, chain_offset, nch + k*p (+ 4 with
 * opts->joint), LNA_ITER2_UNIF, seed, iteration, ...) hands to the host. */
int lna_chains_step_mcmc2_rng(lna_chains *ch, const lna_mcmc2_opts *opts, uint64_t seed, uint64_t iteration,
                              int64_t chain_offset);

/* Stream groups.  A batch of >= 2048 chains runs as G groups on G streams (chains of different groups never interact).
 * By default every step ends by ordering the problem's stream after all groups, so work the caller enqueues there
 * afterwards sees the step's results.  With free-running groups (on = 1) that join is deferred: the groups run ahead of each
 * other ACROSS iterations, so one group's latency-bound stages overlap another group's throughput-bound ones; results are
 * the same bits.  The caller then observes the chains through lna_chains_get / lna_chains_history_get (which synchronise),
 * or calls lna_chains_join (orders the problem's stream after everything enqueued so far: needed before reading
 * lna_chains_summary_dev on that stream or recording a timing event there) / lna_chains_sync (host waits).  Buffers passed to
 * the `_dev` step variants must stay valid until then.  No reference counterpart (the reference is single-threaded). */
int lna_chains_set_free_running(lna_chains *ch, int on);
int lna_chains_join(lna_chains *ch);
int lna_chains_sync(lna_chains *ch);
int lna_chains_groups(const lna_chains *ch);

/* Read the state back (any pointer may be NULL).  counters: B x 4 int32 = (n_full_evals,
 * n_cheap_evals, n_mh_accept, status OR). */
int lna_chains_get(lna_chains *ch, double *par, double *OriginTraj, double *LatentTraj, double *coalLog,
                   double *logOrigin, int32_t *counters);
/* The drivers' per-iteration records (params[i,] = par, l[i] = logOrigin, l1[i] = coalLog; R/LNA_chpt_slice.R:685-690, 842-845)
 * kept on the device: begin(niter) allocates, record(i) stores the current state as iteration i (0-based; three
 * device-to-device copies per stream group, no host round trip), get() returns par as B x niter x npar and l, l1 as B x niter
 * (host buffers, any may be NULL).  2000 iterations of 4096 chains are 2.9 GB. */
int lna_chains_history_begin(lna_chains *ch, int64_t niter);
int lna_chains_history_record(lna_chains *ch, int64_t i);
int lna_chains_history_get(lna_chains *ch, double *par, double *l, double *l1);
/* Device pointer to the B x 4 summary block (coalLog, logOrigin, n_full_evals, status) that is
 * all-gathered across GPUs; valid until the next step. */
const double *lna_chains_summary_dev(lna_chains *ch);

/* ---- multi-GPU ---------------------------------------------------------------------------------------
 * Chains shard contiguously over GPUs (lna_shard); nothing is exchanged inside an iteration.  The one exchange is an
 * all-gather of the per-chain summary rows (coalLog, logOrigin, n_full_evals, status) over NCCL (NVLink 5 / NVSwitch):
 * ONE ncclAllGather on preallocated buffers for equal shards, one grouped launch of broadcasts for ragged ones -- no size
 * exchange, no host synchronisation.  NCCL is bound at run time (dlopen libnccl.so.2, LNA_NCCL_LIB overrides).
 *
 * One process per GPU (torchrun / MPI): rank 0 calls lna_comm_unique_id, the 128 bytes are distributed by the launcher's
 * own means (torch.distributed broadcast, MPI_Bcast, a file), every rank calls lna_comm_create. */
typedef struct lna_comm lna_comm;
#define LNA_COMM_ID_BYTES 128
int lna_comm_available(void);                 /* NCCL version code (> 0) if libnccl could be bound, else 0 (see lna_last_error) */
int lna_comm_unique_id(void *id128);
lna_comm *lna_comm_create(int nranks, int rank, const void *id128, int device);
void lna_comm_destroy(lna_comm *c);
int lna_comm_rank(const lna_comm *c);
int lna_comm_size(const lna_comm *c);
/* rank r owns counts[r] doubles at offset sum(counts[0..r-1]) of recv_dev; asynchronous on `stream` (cudaStream_t) */
int lna_comm_allgather(lna_comm *c, const double *send_dev, double *recv_dev, const int64_t *counts, void *stream);
/* the summary rows of all B_total chains of the job, in global chain order, into out_dev (B_total x 4, this rank's device);
 * this rank's batch must be its lna_shard of B_total.  Asynchronous on the problem's stream. */
int lna_chains_allgather_summaries(lna_chains *ch, lna_comm *c, int64_t B_total, double *out_dev);

/* One process driving several GPUs (an R session behind the Rcpp layer): B_total chains sharded over `devices`, one problem
 * and one resident batch per device.  Every step only enqueues work, so one host thread keeps all devices busy; the Philox
 * streams are keyed by the global chain index, so results do not depend on the number of devices. */
typedef struct lna_multi lna_multi;
lna_multi *lna_multi_create(const lna_cfg *cfg, const lna_init *init, int ndev, const int *devices, int64_t B_total);
void lna_multi_destroy(lna_multi *m);
int lna_multi_devices(const lna_multi *m);
lna_chains *lna_multi_chains(lna_multi *m, int i);    /* the resident batch / problem of device slot i (borrowed) */
lna_problem *lna_multi_problem(lna_multi *m, int i);
int lna_multi_init(lna_multi *m, const double *par, const double *OriginTraj);             /* B_total x npar, B_total x k*p */
int lna_multi_set_free_running(lna_multi *m, int on);
int lna_multi_step_rng(lna_multi *m, const lna_mcmc_opts *opts, uint64_t seed, uint64_t iteration, int64_t chain_offset);
int lna_multi_step_mcmc2_rng(lna_multi *m, const lna_mcmc2_opts *opts, uint64_t seed, uint64_t iteration,
                             int64_t chain_offset);
int lna_multi_set_joint_transform(lna_multi *m, int d, const double *T);
int lna_multi_sync(lna_multi *m);
int lna_multi_get(lna_multi *m, double *par, double *OriginTraj, double *LatentTraj, double *coalLog, double *logOrigin,
                  int32_t *counters);
int lna_multi_history_begin(lna_multi *m, int64_t niter);
int lna_multi_history_record(lna_multi *m, int64_t i);
int lna_multi_history_get(lna_multi *m, int64_t niter, double *par, double *l, double *l1);
/* every device ends up with the B_total x 4 summary rows of all chains (NCCL over NVLink when use_nccl != 0; device-to-host
 * copies otherwise, e.g. with one device listed twice); out_host (may be NULL) receives a host copy */
int lna_multi_gather_summaries(lna_multi *m, int use_nccl, double *out_host);
const double *lna_multi_gathered_dev(lna_multi *m, int i);

/* Measurement: FULL evaluations the chains' stage kernels have EXECUTED since creation, speculative ones included (compare
 * with the consumed count, counters[0] of lna_chains_get), and the number of warp-evaluations they occupied (a warp costs the
 * FP64 pipe the same whatever its lane mask).  Synchronises. */
int lna_chains_exec_counts(lna_chains *ch, uint64_t *evals, uint64_t *warp_evals);

/* ---- measurement helpers --------------------------------------------------------------------- */
/* Debug timeline: with LNA_TIMELINE=1 in the environment every kernel launch is followed by a CUDA event on its stream;
 * this writes "stream,kernel,t_end_us" rows (one GPU-global time axis) to `path` and clears the record.  Use with
 * LNA_GRAPH=0 (launches replayed from a CUDA graph are not marked). */
int lna_debug_timeline_dump(const char *path);
/* Independent-DFMA microbenchmark: returns achieved FP64 FLOP/s of the device (2 flops per FMA),
 * used as the measured denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 entry). */
double lna_measure_fp64_peak(int device, int iters);

#ifdef __cplusplus
}
#endif
#endif /* LNAPHYLODYN_B200_H_ */
