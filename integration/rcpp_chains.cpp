// integration/rcpp_chains.cpp -- MANY chains from one R session.
//
// The reference's drivers run one chain per call (R/LNA_chpt_slice.R:723-849).  This export is the batched entry an R
// user reaches through the same Rcpp layer: n chains resident on the device(s), the loop body of General_MCMC_with_ESlice
// (Update_Param_each_Norm entries c, d; update_Par_ESlice_combine; updateTraj_general_NC from iteration burn1 on) run for
// all of them per iteration, the per-iteration records (params[i, ], l[i], l1[i]) kept in HBM and fetched once.
//
//   res <- General_MCMC_with_ESlice_chains(MCMC_setting,          # from MCMC_setup_general(), unchanged
//                                          par, OriginTraj,       # npar x n_chains, k x p x n_chains (or k x p: shared):
//                                                                 #   the initial states, e.g. n x MCMC_initialize_general()
//                                          ESS_vec, niter, burn1, thin, seed, devices = c(0L, 1L, ...))
//   res$par[i, , c]   = params[i, ] of chain c;  res$l[i, c], res$l1[i, c];  res$LatentTraj[, , c], res$OriginTraj[, , c]: final
//
// Random numbers: the iteration's normals / uniforms are drawn on the device (Philox4x32-10 keyed by `seed`, the GLOBAL chain
// index and the iteration), so results depend neither on the number of devices nor on the R session's RNG state;
// lna_draw_streams hands the same numbers to the host for a draw-for-draw replay through the single-chain exports.
//
// Compiled and run here like the other binding files: `make -C oracle dropin` builds it against the stand-in headers and
// tests/test_gpu_dropin_chains.py drives it through oracle/ref_harness_legacy.cpp (ref_chains_run).
#include "rcpp_shim_util.h"

namespace {

struct MultiGuard {
    lna_multi *m = nullptr;
    ~MultiGuard() {
        if (m) lna_multi_destroy(m);
    }
};

arma::vec list_vec(List l, const char *name) { return as<arma::vec>(l[name]); }

}  // namespace

// [[Rcpp::export()]]
List General_MCMC_with_ESlice_chains(List MCMC_setting, arma::mat par, arma::cube OriginTraj, arma::ivec ESS_vec, int niter,
                                     int burn1, int thin, double seed, arma::ivec devices) {
    // ---- MCMC_setting, as MCMC_setup_general() builds it (R/general_change_point.R:416-435) ----
    arma::vec times = list_vec(MCMC_setting, "times"), x_r = list_vec(MCMC_setting, "x_r");
    arma::vec x_id = list_vec(MCMC_setting, "x_i");
    const double t_correct = as<double>(MCMC_setting["t_correct"]);
    const int gridsize = as<int>(MCMC_setting["gridsize"]);
    List prior = as<List>(MCMC_setting["prior"]), proposal = as<List>(MCMC_setting["proposal"]);
    List init_list = as<List>(MCMC_setting["Init"]);
    InitView ini(init_list);
    if (niter < 1 || thin < 1 || par.n_cols < 1) Rcpp::stop("niter, thin and the number of chains must be positive");
    const int64_t B = (int64_t)par.n_cols;
    lna_cfg cfg;
    cfg.model = LNA_MODEL_SIR;
    cfg.transX = LNA_TRANSX_STANDARD;
    cfg.nch = (int)x_id[0];
    cfg.nparam = (int)x_id[1];
    cfg.iR0 = (int)x_id[2];
    cfg.iGamma = (int)x_id[3];
    cfg.gridsize = gridsize;
    cfg.nt = (int)times.n_elem;
    cfg.times = times.memptr();
    cfg.x_r = x_r.memptr();
    cfg.t_correct = t_correct;
    lna_init li;
    li.E = (int)ini.C.n_elem;
    li.ng = ini.ng;
    li.C = ini.C.memptr();
    li.D = ini.D.memptr();
    li.y = ini.y.memptr();
    li.gridrep = ini.gridrep.memptr();
    std::vector<int> dev(devices.n_elem ? devices.n_elem : 1, 0);
    for (unsigned i = 0; i < devices.n_elem; ++i) dev[i] = (int)devices[i];
    MultiGuard g;
    g.m = lna_multi_create(&cfg, &li, (int)dev.size(), dev.data(), B);
    if (!g.m) fail_abi("lna_multi_create");
    Dims d(lna_multi_problem(g.m, 0));
    const int npar = 2 + d.npt, kp = d.k * d.p;
    if ((int)par.n_rows != npar) Rcpp::stop("par must be npar x n_chains (npar = 2 + nparam + nch + 1)");
    if ((int)(OriginTraj.n_rows * OriginTraj.n_cols) != kp || (OriginTraj.n_slices != 1 && (int64_t)OriginTraj.n_slices != B))
        Rcpp::stop("OriginTraj must be k x p (shared) or k x p x n_chains");
    // item-major host buffers: par is already (npar x B column-major) = B items of npar; OriginTraj slices are k x p column-major
    std::vector<double> eps((size_t)B * kp);
    for (int64_t c = 0; c < B; ++c)
        std::memcpy(eps.data() + (size_t)c * kp, OriginTraj.memptr() + (OriginTraj.n_slices == 1 ? 0 : (size_t)c * kp),
                    sizeof(double) * kp);
    LNA_CALL(lna_multi_init(g.m, par.memptr(), eps.data()), "lna_multi_init");

    // ---- options of the two stages (R/LNA_chpt_slice.R:743, 773-774, 791-792): prior / proposal lists BY POSITION ----
    auto pos = [](List l, int i) { return as<arma::vec>(l[i]); };
    lna_mcmc_opts o_full;
    std::memset(&o_full, 0, sizeof o_full);
    {
        arma::vec pop = pos(prior, 0), r0 = pos(prior, 1), gam = pos(prior, 2), hyp = pos(prior, 4), gprop = pos(proposal, 2);
        const double pr8[8] = {pop[0], pop[1], r0[0], r0[1], gam[0], gam[1], hyp[0], hyp[1]};
        for (int i = 0; i < 8; ++i) o_full.prior8[i] = pr8[i];
        o_full.gamma_prior[0] = gam[0];
        o_full.gamma_prior[1] = gam[1];
        o_full.gamma_prop = gprop[0];
    }
    if (ESS_vec.n_elem != 7) Rcpp::stop("ESS_vec must have 7 entries");
    for (int i = 0; i < 7; ++i) o_full.ESS_vec[i] = (int32_t)ESS_vec[i];
    o_full.update_traj = 1;
    o_full.hyper_noop_mh = 1;
    lna_mcmc_opts o_burn = o_full;  // i < burn1: ESS_temp (ESS_vec[5] <- 0), no trajectory update
    o_burn.ESS_vec[4] = 0;
    o_burn.update_traj = 0;

    LNA_CALL(lna_multi_history_begin(g.m, niter), "lna_multi_history_begin");
    const uint64_t s64 = (uint64_t)seed;
    for (int i = 1; i <= niter; ++i) {  // R's 1-based iteration counter is the Philox iteration
        LNA_CALL(lna_multi_step_rng(g.m, i < burn1 ? &o_burn : &o_full, s64, (uint64_t)i, 0), "lna_multi_step_rng");
        LNA_CALL(lna_multi_history_record(g.m, i - 1), "lna_multi_history_record");
    }
    // B x niter x npar (chain-major) -> the R arrays: par[i, j, c], l[i, c], l1[i, c]
    std::vector<double> hp((size_t)B * niter * npar), hl((size_t)B * niter), hl1((size_t)B * niter);
    LNA_CALL(lna_multi_history_get(g.m, niter, hp.data(), hl.data(), hl1.data()), "lna_multi_history_get");
    arma::cube Rpar(niter, npar, B);
    arma::mat Rl(niter, B), Rl1(niter, B);
    for (int64_t c = 0; c < B; ++c)
        for (int i = 0; i < niter; ++i) {
            for (int j = 0; j < npar; ++j) Rpar(i, j, c) = hp[((size_t)c * niter + i) * npar + j];
            Rl(i, c) = hl[(size_t)c * niter + i];
            Rl1(i, c) = hl1[(size_t)c * niter + i];
        }
    // final state (the MCMC_obj entries a continuation needs)
    arma::mat fpar(npar, B);
    arma::cube feps(d.k, d.p, B), flat(d.nrow, d.p + 1, B);
    arma::vec fcl(B), flo(B);
    std::vector<int32_t> cnt((size_t)B * 4);
    LNA_CALL(lna_multi_get(g.m, fpar.memptr(), feps.memptr(), flat.memptr(), fcl.memptr(), flo.memptr(), cnt.data()),
             "lna_multi_get");
    arma::mat counters(4, B);
    for (int64_t c = 0; c < B; ++c)
        for (int q = 0; q < 4; ++q) counters(q, c) = cnt[(size_t)c * 4 + q];
    List res;
    res["par"] = Rpar;
    res["l"] = Rl;
    res["l1"] = Rl1;
    res["par_final"] = fpar;
    res["OriginTraj"] = feps;
    res["LatentTraj"] = flat;
    res["coalLog"] = fcl;
    res["logOrigin"] = flo;
    res["counters"] = counters;  // rows: FULL evaluations, CHEAP evaluations, Metropolis accepts, status OR
    res["thin"] = thin;          // (every iteration is recorded; subsample rows seq(thin, niter, by = thin) like the R driver)
    return res;
}
