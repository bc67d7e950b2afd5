// integration/rcpp_shim.cpp -- the reference-side binding.
//
// Drop this file into the reference package's src/ IN PLACE OF the bodies of the hot-path exports of
// src/LNA_functional.cpp (and ODE2 / SIRS_KOM_Filter / log_like_trajSIRS of src/SIRS_period.cpp, coal_loglik of
// src/SIR_phylodyn.cpp): every function below has the reference's exact C++ signature (src/LNA_functional.h:15-120,
// R/RcppExports.R:27-181), so src/RcppExports.cpp, R/RcppExports.R and all R code stay as they are, and forwards to the
// C ABI of include/lnaphylodyn_b200.h with B = 1.  Armadillo objects are column-major doubles, exactly the item layout
// of the C ABI, so memptr() goes straight through.
//
// Random numbers: the C ABI takes pre-drawn streams.  Each sampler therefore draws what the reference's loop COULD
// consume, in the reference's order, calls the device, and then rewinds the generator to what
// the reference WOULD have consumed (1 + number of slice angles actually used), so that the R session's RNG state
// after the call is the state the unmodified package would leave (RngMark below).
//
// R and Rcpp are not in the build image.  To keep this file honest it IS compiled and exercised here: the test build
// (make -C oracle dropin) compiles it against the stand-in headers of oracle/shim/ -- the same ones the reference's own
// sources compile against for oracle/_ref -- together with oracle/ref_harness.cpp, and tests/test_gpu_dropin.py runs the
// same harness entry points against this file (on the GPU) and against the reference's own bodies (on the CPU).
#include "rcpp_shim_util.h"

// ---- a1: betaTs :31-58, param_transform :63-86 -------------------------------------------------------------------------
arma::vec betaTs(arma::vec param, arma::vec times, arma::vec x_r, arma::ivec x_i) {
    arma::vec t01(2);
    t01[1] = 1.0;
    const int xi[4] = {(int)x_i[0], std::max((int)x_i[1], 2), (int)x_i[2], (int)x_i[3]};
    lna_problem *pb = problem_for(t01, x_r, xi, 1, 0.0, InitView(), LNA_MODEL_SIR);
    Dims d(pb);
    arma::vec q(d.npt), betas(times.n_elem);
    for (int i = 0; i < d.npt && i < (int)param.n_elem; ++i) q[i] = param[i];
    LNA_CALL(lna_betaTs(pb, 1, q.memptr(), times.memptr(), (int)times.n_elem, betas.memptr()), "lna_betaTs");
    return betas;
}

arma::vec param_transform(double t, arma::vec param, arma::vec x_r, arma::ivec x_i) {
    arma::vec t01(2);
    t01[1] = 1.0;
    const int xi[4] = {(int)x_i[0], std::max((int)x_i[1], 2), (int)x_i[2], (int)x_i[3]};
    lna_problem *pb = problem_for(t01, x_r, xi, 1, 0.0, InitView(), LNA_MODEL_SIR);
    Dims d(pb);
    arma::vec q(d.npt), out(d.npt);
    for (int i = 0; i < d.npt && i < (int)param.n_elem; ++i) q[i] = param[i];
    LNA_CALL(lna_param_transform(pb, 1, &t, q.memptr(), out.memptr()), "lna_param_transform");
    arma::vec param2 = param;
    param2[0] = out[0];
    return param2;
}

// ---- a3: ODE_rk45 :455-491, ODE_rk45_stop :495-530 ----------------------------------------------------------------------
arma::mat ODE_rk45(arma::vec initial, arma::vec t, arma::vec param, arma::vec x_r, arma::ivec x_i, std::string transP,
                   std::string model, std::string transX) {
    check_transP(transP);
    lna_problem *pb = problem_xi(t, x_r, x_i, 1, 0.0, InitView(), model, transX);
    Dims d(pb);
    arma::vec q = pad_param(param, d.npt);
    arma::mat OdeTraj(t.n_elem, d.p + 1);
    LNA_CALL(lna_ode_rk45(pb, 1, initial.memptr(), q.memptr(), OdeTraj.memptr()), "lna_ode_rk45");
    return OdeTraj;
}

double ODE_rk45_stop(arma::vec initial, arma::vec t, arma::vec param, arma::vec x_r, arma::ivec x_i, std::string transP,
                     std::string model, std::string transX, double tol) {
    arma::mat traj = ODE_rk45(initial, t, param, x_r, x_i, transP, model, transX);  // trajectory on the device, scan here
    const int col = 1 + (int)x_i[3], n = (int)t.n_elem;
    for (int i = 1; i < n; ++i)
        if (traj(i, col) < tol) return t[i];
    return t[n - 1] + 0.1;
}

// ---- a5/a6: SigmaF :536-599, KF_param :603-634, KF_param_chol :638-670 --------------------------------------------------
static List kf_param_impl(const arma::mat &OdeTraj, const arma::vec &param, int gridsize, const arma::vec &x_r,
                          const arma::ivec &x_i, const std::string &model, const std::string &transX, int chol) {
    arma::vec times = OdeTraj.col(0);
    lna_problem *pb = problem_xi(times, x_r, x_i, gridsize, 0.0, InitView(), model, transX);
    Dims d(pb);
    arma::vec q = pad_param(param, d.npt);
    arma::cube A(d.p, d.p, d.k), S(d.p, d.p, d.k);
    int32_t st = 0;
    LNA_CALL(lna_kf_param(pb, 1, OdeTraj.memptr(), q.memptr(), chol, A.memptr(), S.memptr(), &st), "lna_kf_param");
    if (chol && (st & LNA_ST_CHOL_FAIL)) throw std::invalid_argument(kCholMsg);
    return make_FT(A, S);
}
List KF_param(arma::mat OdeTraj, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i, std::string transP,
              std::string model, std::string transX) {
    check_transP(transP);
    return kf_param_impl(OdeTraj, param, gridsize, x_r, x_i, model, transX, 0);
}
List KF_param_chol(arma::mat OdeTraj, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i, std::string transP,
                   std::string model, std::string transX) {
    check_transP(transP);
    return kf_param_impl(OdeTraj, param, gridsize, x_r, x_i, model, transX, 1);
}
List SigmaF(arma::mat Traj_par, arma::vec param, arma::vec x_r, arma::ivec x_i, std::string transP, std::string model,
            std::string transX) {
    check_transP(transP);
    // one LNA interval = KF_param with gridsize = nrow(Traj_par); a closing row (never read) completes the grid
    const int g = (int)Traj_par.n_rows, nc = (int)Traj_par.n_cols;
    arma::mat T(g + 1, nc);
    for (int j = 0; j < nc; ++j) {
        for (int i = 0; i < g; ++i) T(i, j) = Traj_par(i, j);
        T(g, j) = Traj_par(g - 1, j);
    }
    T(g, 0) = Traj_par(g - 1, 0) + (Traj_par(1, 0) - Traj_par(0, 0));
    List kf = kf_param_impl(T, param, g, x_r, x_i, model, transX, 0);
    arma::cube A = as<arma::cube>(kf[0]), S = as<arma::cube>(kf[1]);
    List Res;
    Res["expF"] = arma::mat(A.slice(0));
    Res["Sigma"] = arma::mat(S.slice(0));
    return Res;
}

// ---- a7: Ode_Coarse_Slicer :1179-1188, New_Param_List :1191-1208 ---------------------------------------------------------
arma::mat Ode_Coarse_Slicer(arma::mat Ode_thin, int gridsize) {
    const int p = (int)Ode_thin.n_cols - 1;
    arma::vec times = Ode_thin.col(0), x_r(1);
    x_r[0] = 1.0;
    const int xi[4] = {0, p == 2 ? 2 : 3, 0, 1};
    lna_problem *pb = problem_for(times, x_r, xi, gridsize, 0.0, InitView(), p == 2 ? LNA_MODEL_SIR : LNA_MODEL_SEIR);
    Dims d(pb);
    arma::mat out(d.nrow, p + 1);
    LNA_CALL(lna_ode_coarse_slicer(pb, 1, Ode_thin.memptr(), out.memptr()), "lna_ode_coarse_slicer");
    return out;
}

List New_Param_List(arma::vec param, arma::vec initial, int gridsize, arma::vec t, arma::vec x_r, arma::ivec x_i,
                    std::string transP, std::string model, std::string transX) {
    check_transP(transP);
    lna_problem *pb = problem_xi(t, x_r, x_i, gridsize, 0.0, InitView(), model, transX);
    Dims d(pb);
    arma::vec q = pad_param(param, d.npt), betaN(d.nrow);
    arma::cube A(d.p, d.p, d.k), L(d.p, d.p, d.k);
    arma::mat Ode(d.nrow, d.p + 1);
    int32_t st = 0;
    LNA_CALL(lna_new_param_list(pb, 1, q.memptr(), initial.memptr(), A.memptr(), L.memptr(), Ode.memptr(), betaN.memptr(), &st),
             "lna_new_param_list");
    if (st & LNA_ST_CHOL_FAIL) throw std::invalid_argument(kCholMsg);  // the exception the reference throws (:663)
    List result;
    result["FT"] = make_FT(A, L);
    result["Ode"] = Ode;
    result["betaN"] = betaN;
    return result;
}

// ---- a8: TransformTraj :877-900, Traj_sim_general_noncentral :824-872, Traj_sim_ezG_NC :980-1004 ------------------------
arma::mat TransformTraj(arma::mat OdeTraj, arma::mat OriginLatent, List Filter_NC) {
    lna_problem *pb = traj_problem(OdeTraj, 0.0, InitView(), 0, 1);
    arma::cube A = as<arma::cube>(Filter_NC[0]), L = as<arma::cube>(Filter_NC[1]);
    arma::mat out(OdeTraj.n_rows, OdeTraj.n_cols);
    LNA_CALL(lna_transform_traj(pb, 1, OdeTraj.memptr(), OriginLatent.memptr(), A.memptr(), L.memptr(), out.memptr()),
             "lna_transform_traj");
    return out;
}

List Traj_sim_general_noncentral(arma::mat OdeTraj, List Filter_NC, double t_correct) {
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, 1);
    const int k = (int)OdeTraj.n_rows - 1, p = (int)OdeTraj.n_cols - 1;
    arma::cube A = as<arma::cube>(Filter_NC[0]), L = as<arma::cube>(Filter_NC[1]);
    arma::vec z = arma::randn(k * p, 1);  // the reference draws randn(p,1) per step, in step order (:850)
    arma::mat traj(k + 1, p + 1), origin(k, p);
    double lm = 0, lo = 0;
    LNA_CALL(lna_traj_sim_noncentral(pb, 1, OdeTraj.memptr(), A.memptr(), L.memptr(), z.memptr(), traj.memptr(),
                                     origin.memptr(), &lm, &lo),
             "lna_traj_sim_noncentral");
    List Res;
    Res["SimuTraj"] = traj;
    Res["OriginTraj"] = origin;
    Res["logMultiNorm"] = lm;
    Res["logOrigin"] = lo;
    return Res;
}

List Traj_sim_ezG_NC(arma::vec initial, arma::vec times, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i,
                     double t_correct, std::string transP, std::string model, std::string transX) {
    check_transP(transP);
    List npl = New_Param_List(param, initial, gridsize, times, x_r, x_i, transP, model, transX);
    return Traj_sim_general_noncentral(as<arma::mat>(npl[1]), as<List>(npl[0]), t_correct);
}

// ---- a9: volz_loglik_nh2 :1119-1176, volz_loglik_nh3 :1058-1113, coal_loglik3 :1016-1054, coal_loglik -----------------
static double volz_impl(List init, const arma::mat &f1, const arma::vec &betaN, double t_correct, const arma::ivec &index,
                        const std::string &transX, bool nan_guard) {
    lna_problem *pb = traj_problem(f1, t_correct, InitView(init), (int)index[0], (int)index[1], transx_id(transX));
    double ll = 0;
    int32_t st = 0;
    if (nan_guard) LNA_CALL(lna_volz_loglik_nh2(pb, 1, f1.memptr(), betaN.memptr(), &ll, &st), "lna_volz_loglik_nh2");
    else LNA_CALL(lna_volz_loglik_nh3(pb, 1, f1.memptr(), betaN.memptr(), &ll, &st), "lna_volz_loglik_nh3");
    return ll;  // the -1e7 sentinels come back in-band, like the reference
}
double volz_loglik_nh2(List init, arma::mat f1, arma::vec betaN, double t_correct, arma::ivec index, std::string transX) {
    return volz_impl(init, f1, betaN, t_correct, index, transX, true);
}
double volz_loglik_nh3(List init, arma::mat f1, arma::vec betaN, double t_correct, arma::ivec index, std::string transX) {
    return volz_impl(init, f1, betaN, t_correct, index, transX, false);
}
static double kingman_impl(List init, const arma::mat &f1, double t_correct, double lambda, int col, int take_log) {
    lna_problem *pb = traj_problem(f1, t_correct, InitView(init), 0, 1);
    double ll = 0;
    LNA_CALL(lna_coal_loglik(pb, 1, f1.memptr(), &lambda, col, take_log, &ll), "lna_coal_loglik");
    return ll;
}
double coal_loglik3(List init, arma::mat f1, double t_correct, double lambda, int Index, std::string transX) {
    return kingman_impl(init, f1, t_correct, lambda, Index + 1, transX == "standard" ? 1 : 0);
}
double coal_loglik(List init, arma::mat f1, double t_correct, double lambda, int gridsize, std::string transX) {
    return kingman_impl(init, f1, t_correct, lambda, 2, 0);  // SIR_phylodyn.cpp:436-468: f1[,3] already holds log I
}

// ---- a10: Update_Param :1212-1243 ------------------------------------------------------------------------------------------
List Update_Param(arma::vec param, arma::vec initial, arma::vec t, arma::mat OriginTraj, arma::vec x_r, arma::ivec x_i,
                  List init, int gridsize, double coal_log, double prior_proposal_offset, double t_correct,
                  std::string transP, std::string model, std::string transX, bool volz) {
    check_transP(transP);
    lna_problem *pb = problem_xi(t, x_r, x_i, gridsize, t_correct, InitView(init), model, transX);
    Dims d(pb);
    arma::vec q = pad_param(param, d.npt), betaN(d.nrow);
    arma::cube A(d.p, d.p, d.k), L(d.p, d.p, d.k);
    arma::mat Ode(d.nrow, d.p + 1), Lat(d.nrow, d.p + 1);
    const double u = R::runif(0, 1);  // the one draw of the reference (:1230)
    int32_t acc = 0, st = 0;
    double cl = 0;
    LNA_CALL(lna_update_param(pb, 1, q.memptr(), initial.memptr(), OriginTraj.memptr(), &coal_log, &prior_proposal_offset, &u,
                              &acc, &cl, &st, A.memptr(), L.memptr(), Ode.memptr(), betaN.memptr(), Lat.memptr()),
             "lna_update_param");
    if (st & LNA_ST_CHOL_FAIL) throw std::invalid_argument(kCholMsg);
    List Result;
    if (acc) {
        Result["accept"] = true;
        Result["FT"] = make_FT(A, L);
        Result["Ode"] = Ode;
        Result["betaN"] = betaN;
        Result["coalLog"] = cl;
        Result["LatentTraj"] = Lat;
    } else {
        Result["accept"] = false;
    }
    return Result;
}

// ---- a11: Param_Slice_update :1246-1253, Param_Slice_update_all2 :1268-1314 ----------------------------------------------
static lna_problem *slice_problem(const arma::ivec &x_i) {
    arma::vec t01(2), x_r((int)x_i[0] + 1);
    t01[1] = 1.0;
    x_r[0] = 1.0;
    for (int j = 0; j < (int)x_i[0]; ++j) x_r[j + 1] = 1.0 + j;
    const int xi[4] = {(int)x_i[0], (int)x_i[1], (int)x_i[2], (int)x_i[3]};
    return problem_for(t01, x_r, xi, 1, 0.0, InitView(), LNA_MODEL_SIR);
}
arma::vec Param_Slice_update(arma::vec param, arma::vec x_r, arma::ivec x_i, double theta, arma::vec newChs, double rho) {
    lna_problem *pb = slice_problem(x_i);
    arma::vec out(param.n_elem);
    LNA_CALL(lna_param_slice_update(pb, 1, param.memptr(), &theta, newChs.memptr(), out.memptr()), "lna_param_slice_update");
    return out;
}
arma::vec Param_Slice_update_all2(arma::vec par_old, arma::vec x_r, arma::ivec x_i, double theta, arma::vec newChs,
                                  arma::ivec ESS_vec, List priorList) {
    lna_problem *pb = slice_problem(x_i);
    double pr[8];
    for (int i = 0; i < 4; ++i) {
        arma::vec v = as<arma::vec>(priorList[i]);
        pr[2 * i] = v[0];
        pr[2 * i + 1] = v[1];
    }
    std::vector<int32_t> ess = to_i32(ESS_vec);
    arma::vec out(par_old.n_elem);
    LNA_CALL(lna_param_slice_update_all2(pb, 1, par_old.memptr(), &theta, newChs.memptr(), ess.data(), pr, out.memptr()),
             "lna_param_slice_update_all2");
    return out;
}

// ---- a12: ESlice_par_General :1532-1688 ---------------------------------------------------------------------------------
List ESlice_par_General(arma::vec par_old, arma::vec t, arma::mat OriginTraj, List priorList, arma::vec x_r, arma::ivec x_i,
                        List init, int gridsize, arma::ivec ESS_vec, double coal_log, double t_correct, std::string transP,
                        std::string model, std::string transX, bool volz) {
    check_transP(transP);
    if (ESS_vec[6] != 0) Rcpp::stop("GPU path: ESS_vec[7] (the self-mixing OriginTraj branch, :1660-1662) is not built");
    lna_problem *pb = problem_xi(t, x_r, x_i, gridsize, t_correct, InitView(init), model, transX);
    Dims d(pb);
    const int npar = (int)par_old.n_elem;
    arma::vec nu = arma::randn(npar - 1, 1);  // :1573
    RngMark rng;
    rng.mark();
    arma::vec un = draw_uniforms(LNA_ESLICE_MAX_UNIF);  // u (:1583), then one per slice angle (:1625, :1635)
    double pr[8];
    for (int i = 0; i < 4; ++i) {
        arma::vec v = as<arma::vec>(priorList[i]);
        pr[2 * i] = v[0];
        pr[2 * i + 1] = v[1];
    }
    std::vector<int32_t> ess = to_i32(ESS_vec);
    arma::vec par_new(npar), betaN(d.nrow);
    arma::cube A(d.p, d.p, d.k), L(d.p, d.p, d.k);
    arma::mat Ode(d.nrow, d.p + 1), Lat(d.nrow, d.p + 1);
    double cl = 0, lo = 0;
    int32_t ne = 0, st = 0;
    LNA_CALL(lna_eslice_par_general(pb, 1, par_old.memptr(), OriginTraj.memptr(), pr, ess.data(), &coal_log, nu.memptr(),
                                    un.memptr(), /*spec_width=*/0, par_new.memptr(), A.memptr(), L.memptr(), Ode.memptr(),
                                    betaN.memptr(), Lat.memptr(), &cl, &lo, &ne, &st),
             "lna_eslice_par_general");
    rng.settle(1 + ne - ((st & LNA_ST_CAP_HIT) ? 1 : 0));  // the cap's re-evaluation draws no angle (:1609-1622)
    if (st & LNA_ST_CHOL_FAIL) throw std::invalid_argument(kCholMsg);  // only the cap's recompute can throw (:1613)
    List result;
    result["betaN"] = betaN;
    result["FT"] = make_FT(A, L);
    result["OdeTraj"] = Ode;
    result["par"] = par_new;
    result["LatentTraj"] = Lat;
    result["CoalLog"] = cl;
    result["OriginTraj"] = OriginTraj;
    result["logOrigin"] = lo;
    return result;
}

// ---- a13: ESlice_general_NC :1690-1778 ----------------------------------------------------------------------------------
List ESlice_general_NC(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec betaN,
                       double t_correct, double lambda, double coal_log, int gridsize, bool volz, std::string model,
                       std::string transX) {
    if (!volz) Rcpp::stop("GPU path: ESlice_general_NC(volz = FALSE), the Kingman branch, is not built");
    const int iI = model == "SEIR" ? 2 : 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(init), 0, iI);
    const int k = (int)f_cur.n_rows, p = (int)f_cur.n_cols;
    arma::cube A = as<arma::cube>(FTs[0]), L = as<arma::cube>(FTs[1]);
    arma::mat v_traj = arma::randn(k, p);  // :1706
    RngMark rng;
    rng.mark();
    arma::vec un = draw_uniforms(LNA_ESLICE_MAX_UNIF);
    arma::mat Lat(k + 1, p + 1), f_prime(k, p);
    double lo = 0, cl = 0;
    int32_t ne = 0, st = 0;
    LNA_CALL(lna_eslice_general_nc(pb, 1, f_cur.memptr(), OdeTraj.memptr(), A.memptr(), L.memptr(), betaN.memptr(), &coal_log,
                                   v_traj.memptr(), un.memptr(), Lat.memptr(), f_prime.memptr(), &lo, &cl, &ne, &st),
             "lna_eslice_general_nc");
    // evaluations that draw no angle: the recomputation of the current likelihood (coal_log == -99999999) and the cap
    rng.settle(1 + ne - (coal_log == -99999999 ? 1 : 0) - ((st & LNA_ST_CAP_HIT) ? 1 : 0));
    List Res;
    Res["LatentTraj"] = Lat;
    Res["OriginTraj"] = f_prime;
    Res["logOrigin"] = lo;
    Res["CoalLog"] = cl;
    return Res;
}

// ---- a14: ESlice_change_points :1317-1409 ---------------------------------------------------------------------------------
List ESlice_change_points(arma::vec param, arma::vec initial, arma::vec t, arma::mat OriginTraj, arma::vec x_r,
                          arma::ivec x_i, List init, int gridsize, double coal_log, double t_correct, std::string transP,
                          std::string model, std::string transX, bool volz) {
    check_transP(transP);
    lna_problem *pb = problem_xi(t, x_r, x_i, gridsize, t_correct, InitView(init), model, transX);
    Dims d(pb);
    arma::vec un(LNA_ESLICE_MAX_UNIF);
    un[0] = R::runif(0, 1);  // u (:1331)
    un[1] = R::runif(0, 1);  // theta_1 (:1340)
    arma::vec nu = arma::randn((int)x_i[0], 1);  // :1344
    RngMark rng;
    rng.mark();
    for (int i = 2; i < LNA_ESLICE_MAX_UNIF; ++i) un[i] = R::runif(0, 1);
    arma::vec q = pad_param(param, d.npt), param_new(d.npt), betaN(d.nrow);
    arma::cube A(d.p, d.p, d.k), L(d.p, d.p, d.k);
    arma::mat Ode(d.nrow, d.p + 1), Lat(d.nrow, d.p + 1);
    double cl = 0, lcp = 0;
    int32_t ne = 0, st = 0;
    LNA_CALL(lna_eslice_change_points(pb, 1, q.memptr(), initial.memptr(), OriginTraj.memptr(), &coal_log, nu.memptr(),
                                      un.memptr(), 0, param_new.memptr(), A.memptr(), L.memptr(), Ode.memptr(), betaN.memptr(),
                                      Lat.memptr(), &cl, &lcp, &ne, &st),
             "lna_eslice_change_points");
    rng.settle(ne - 1 - ((st & LNA_ST_CAP_HIT) ? 1 : 0));  // angles after the first one
    if (st & LNA_ST_CHOL_FAIL) throw std::invalid_argument(kCholMsg);  // this loop has no try/catch (:1347)
    List result;
    result["betaN"] = betaN;
    result["FT"] = make_FT(A, L);
    result["OdeTraj"] = Ode;
    result["param"] = param_new;
    result["LatentTraj"] = Lat;
    result["CoalLog"] = cl;
    result["LogChProb"] = lcp;
    return result;
}

// ---- f3: centred parametrisation ------------------------------------------------------------------------------------------
double log_like_traj_general2(arma::mat SdeTraj, arma::mat OdeTraj, List Filter, int gridsize, double t_correct) {
    const int p = (int)OdeTraj.n_cols - 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, p == 2 ? 1 : 2);
    arma::cube A = as<arma::cube>(Filter[0]), S = as<arma::cube>(Filter[1]);
    double ll = 0;
    LNA_CALL(lna_log_like_traj_general2(pb, 1, SdeTraj.memptr(), OdeTraj.memptr(), A.memptr(), S.memptr(), &ll),
             "lna_log_like_traj_general2");
    return ll;
}
double log_like_traj_general_adjust(arma::mat SdeTraj, arma::mat OdeTraj, List Filter_NC, int gridsize, double t_correct) {
    return 0;  // the reference's body is commented out (:746-767)
}
List Traj_sim_general3(arma::mat OdeTraj, List Filter, double t_correct) {
    const int k = (int)OdeTraj.n_rows - 1, p = (int)OdeTraj.n_cols - 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, p == 2 ? 1 : 2);
    arma::cube A = as<arma::cube>(Filter[0]), S = as<arma::cube>(Filter[1]);
    arma::vec z = arma::randn(k * p, 1);  // mvrnormArma draws randn(p,1) per step (basic_util.cpp:44-46)
    arma::mat traj(k + 1, p + 1);
    double ll = 0;
    int32_t st = 0;
    LNA_CALL(lna_traj_sim_general3(pb, 1, OdeTraj.memptr(), A.memptr(), S.memptr(), z.memptr(), traj.memptr(), &ll, &st),
             "lna_traj_sim_general3");
    if (st & LNA_ST_CHOL_FAIL) throw std::runtime_error("chol(): decomposition failed");
    List Res;
    Res["SimuTraj"] = traj;
    Res["loglike"] = ll;
    return Res;
}
List Traj_sim_ezG2(arma::vec initial, arma::vec times, arma::vec param, int gridsize, arma::vec x_r, arma::ivec x_i,
                   double t_correct, std::string transP, std::string model, std::string transX) {
    check_transP(transP);
    arma::mat ode = ODE_rk45(initial, times, param, x_r, x_i, transP, model, transX);
    List kf = KF_param(ode, param, gridsize, x_r, x_i, transP, model, transX);
    return Traj_sim_general3(Ode_Coarse_Slicer(ode, gridsize), kf, t_correct);
}
arma::mat ESlice_general2(arma::mat f_cur, arma::mat OdeTraj, List FTs, arma::vec state, List init, arma::vec betaN,
                          double t_correct, double lambda, int reps, int gridsize, bool volz, std::string model,
                          std::string transX) {
    if (!volz) Rcpp::stop("GPU path: ESlice_general2(volz = FALSE), the Kingman branch, is not built");
    const int k = (int)OdeTraj.n_rows - 1, p = (int)OdeTraj.n_cols - 1;
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(init), 0, model == "SEIR" ? 2 : 1);
    arma::cube A = as<arma::cube>(FTs[0]), S = as<arma::cube>(FTs[1]);
    arma::vec z = arma::randn(k * p, 1);  // the inner Traj_sim_general3 (:1801)
    RngMark rng;
    rng.mark();
    arma::vec un = draw_uniforms(LNA_ESLICE_MAX_UNIF);
    arma::mat newTraj(k + 1, p + 1);
    int32_t ne = 0, st = 0;
    LNA_CALL(lna_eslice_general2(pb, 1, f_cur.memptr(), OdeTraj.memptr(), A.memptr(), S.memptr(), betaN.memptr(), z.memptr(),
                                 un.memptr(), newTraj.memptr(), &ne, &st),
             "lna_eslice_general2");
    rng.settle(ne);  // u + one angle per candidate; the first evaluation (of f_cur, :1810) draws none
    return newTraj;
}

// ---- row 13: seasonal SIRS (src/SIRS_period.cpp) ------------------------------------------------------------------------------
static lna_problem *sirs_problem(const arma::vec &times, double period, int gridsize) {
    arma::vec x_r(1);
    x_r[0] = period;
    const int xi[4] = {0, 4, 0, 1};
    return problem_for(times, x_r, xi, gridsize, 0.0, InitView(), LNA_MODEL_SIRS_PERIOD);
}
arma::mat ODE2(arma::vec initial, arma::vec t, arma::vec param, double period) {
    lna_problem *pb = sirs_problem(t, period, 1);
    arma::mat out(t.n_elem, 4);
    LNA_CALL(lna_ode_rk45(pb, 1, initial.memptr(), param.memptr(), out.memptr()), "lna_ode_rk45");
    return out;
}
List SIRS_KOM_Filter(arma::mat OdeTraj, arma::vec param, int gridsize, double period) {
    arma::vec times = OdeTraj.col(0);
    lna_problem *pb = sirs_problem(times, period, gridsize);
    Dims d(pb);
    arma::cube A(3, 3, d.k), S(3, 3, d.k);
    int32_t st = 0;
    LNA_CALL(lna_kf_param(pb, 1, OdeTraj.memptr(), param.memptr(), 0, A.memptr(), S.memptr(), &st), "lna_kf_param");
    return make_FT(A, S);
}
double log_like_trajSIRS(arma::mat SdeTraj, arma::mat OdeTraj, List Filter, int gridsize, double t_correct) {
    lna_problem *pb = traj_problem(OdeTraj, t_correct, InitView(), 0, 2);  // any 3-compartment problem on this grid
    arma::cube A = as<arma::cube>(Filter[0]), S = as<arma::cube>(Filter[1]);
    double ll = 0;
    LNA_CALL(lna_log_like_traj_sirs(pb, 1, SdeTraj.memptr(), OdeTraj.memptr(), A.memptr(), S.memptr(), &ll),
             "lna_log_like_traj_sirs");
    return ll;
}
