// lna_capi_legacy.inl -- C ABI of the reference's legacy copies of the pipeline; included by
// lna_capi.cu.  Kernels: lna_legacy.cuh (+ k_centred with the inv2 density).
// ------------------------------------------------------------------------------------------------

/* log_like_traj_general (LNA_NS_general.cpp:164-200) = Foo::LNA_log_like_traj (generalLNA.cpp:178-214) */
extern "C" int lna_log_like_traj_general(lna_problem *pb, int64_t B, const double *SdeTraj, const double *OdeTraj,
                                         const double *Acube, const double *Scube, double *loglik) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (P.p != 2) return fail("lna_log_like_traj_general: the legacy copies are SIR-only (p = 2)");
    const size_t nB = (size_t)B, per = (size_t)P.nrow * 3, cube = 4 * (size_t)P.k;
    Stage s(pb);
    const double *dX = s.in(SdeTraj, nB * per), *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube),
                 *dS = s.in(Scube, nB * cube);
    double *dll = s.tmp<double>(nB);
    if (!s.ok || !dX || !dO || !dA || !dS) return fail("lna_log_like_traj_general: bad arguments / staging failed");
    if (centred_launch(pb, B, 0 | 8, dX, dO, dA, dS, nullptr, dll, nullptr)) return -1;
    s.back(loglik, dll, nB);
    return s.finish("lna_log_like_traj_general");
}

/* Traj_sim_general (LNA_NS_general.cpp:205-250) = Foo::LNA_Traj_sim (generalLNA.cpp:217-262) */
extern "C" int lna_traj_sim_general(lna_problem *pb, int64_t B, const double *OdeTraj, const double *Acube,
                                    const double *Scube, const double *normals, double *LNA_traj, double *loglike) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (P.p != 2) return fail("lna_traj_sim_general: the legacy copies are SIR-only (p = 2)");
    const size_t nB = (size_t)B, per = (size_t)P.nrow * 3, cube = 4 * (size_t)P.k;
    Stage s(pb);
    const double *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube), *dS = s.in(Scube, nB * cube),
                 *dz = s.in(normals, nB * P.k * 2);
    double *dT = s.tmp<double>(nB * per), *dll = s.tmp<double>(nB);
    if (!s.ok || !dO || !dA || !dS || !dz) return fail("lna_traj_sim_general: bad arguments / staging failed");
    if (centred_launch(pb, B, 1 | 8, dz, dO, dA, dS, dT, dll, nullptr)) return -1;
    s.back(LNA_traj, dT, nB * per);
    s.back(loglike, dll, nB);
    return s.finish("lna_traj_sim_general");
}

/* volz_loglik_nh (SIR_phylodyn.cpp:517-560) on LOG trajectories (B x nrow x 3) */
extern "C" int lna_volz_loglik_nh(lna_problem *pb, int64_t B, const double *f1_log, const double *betaN, double *loglik) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (!P.has_init) return fail("lna_volz_loglik_nh: problem was created without a genealogy");
    const size_t nB = (size_t)B;
    Stage s(pb);
    const double *dF = s.in(f1_log, nB * P.nrow * 3), *db = s.in(betaN, nB * P.nrow);
    double *dll = s.tmp<double>(nB);
    if (!s.ok || !dF || !db) return fail("lna_volz_loglik_nh: bad arguments / staging failed");
    k_volz_nh_log<<<nblk(B, 64), 64, 0, pb->stream>>>(P, B, dF, db, dll);
    if (launch_check(pb, "k_volz_nh_log")) return -1;
    s.back(loglik, dll, nB);
    return s.finish("lna_volz_loglik_nh");
}

/* ESlice_general (LNA_NS_general.cpp:320-394; kind 0), ESlice_SIRS (SIRS_period.cpp:313-402; kind 1), Foo::LNA_ESlice
 * (generalLNA.cpp:296-372; kind 2 = kind 0).  The problem's compartment count decides p (2 / 3). */
extern "C" int lna_eslice_legacy(lna_problem *pb, int64_t B, int kind, int reps, int volz, const double *f_cur,
                                 const double *OdeTraj, const double *Acube, const double *Scube, const double *state,
                                 const double *betaN_or_params4, const double *lambda, const double *normals,
                                 const double *uniforms, double *newTraj, int32_t *n_evals, int32_t *n_unif, int32_t *status) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (!P.has_init) return fail("lna_eslice_legacy: problem was created without a genealogy");
    if (kind < 0 || kind > 2 || reps < 1) return fail("lna_eslice_legacy: kind must be 0, 1 or 2 and reps >= 1");
    const int p = kind == 1 ? 3 : 2;
    if (P.p != p) return fail("lna_eslice_legacy: kind %d needs a %d-compartment problem", kind, p);
    if (!f_cur || !OdeTraj || !Acube || !Scube || !betaN_or_params4 || !normals || !uniforms || (kind == 1 && !state) ||
        (!volz && !lambda))
        return fail("lna_eslice_legacy: null input");
    const size_t nB = (size_t)B, per = (size_t)P.nrow * (p + 1), cube = (size_t)p * p * P.k, nrow = P.nrow, R = (size_t)reps;
    Stage s(pb);
    const double *dF = s.in(f_cur, nB * per), *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube),
                 *dS = s.in(Scube, nB * cube), *dz = s.in(normals, R * nB * P.k * p), *du = s.in(uniforms, nB * R * 22),
                 *dI = kind == 1 ? s.in(state, nB * 3) : nullptr, *dlam = volz ? nullptr : s.in(lambda, nB),
                 *dq = s.in(betaN_or_params4, kind == 1 ? nB * 4 : nB * nrow);
    double *dV = s.tmp<double>(R * nB * per), *dN = s.tmp<double>(nB * per), *dC = s.tmp<double>(nB * per),
           *dll = s.tmp<double>(nB), *db = kind == 1 ? s.tmp<double>(nB * nrow) : nullptr;
    int32_t *dne = s.tmp<int32_t>(nB), *dnu = s.tmp<int32_t>(nB), *dst = s.tmp<int32_t>(nB), *dvs = s.tmp<int32_t>(R * nB);
    if (!s.ok) return fail("lna_eslice_legacy: device staging failed");
    // the proposals of every repetition: Traj_sim_general / Traj_sim_SIRS do not depend on f_cur (normals repetition-major)
    for (size_t r = 0; r < R; ++r)
        if (centred_launch(pb, B, kind == 1 ? 3 : (1 | 8), dz + r * nB * P.k * p, dO, dA, dS, dV + r * nB * per, dll,
                           dvs + r * nB, dI))
            return -1;
    if (kind == 1) {
        k_beta_dyn<<<nblk(B * (long long)nrow, 128), 128, 0, pb->stream>>>(B, (int)nrow, (int)per, dF, dq, db);
        if (launch_check(pb, "k_beta_dyn")) return -1;
    }
    LegacyEssArgs a{};
    a.f_cur = dF;
    a.ode = dO;
    a.v = dV;
    a.betaN = kind == 1 ? db : dq;
    a.uniforms = du;
    a.lambda = dlam;
    a.v_status = kind == 1 ? dvs : nullptr;
    a.cur = dC;
    a.newTraj = dN;
    a.n_evals = dne;
    a.status = dst;
    a.n_unif = dnu;
    a.p = p;
    a.reps = reps;
    a.like = volz ? 0 : 1;
    a.cap_keeps_last = kind == 1;
    k_eslice_legacy<<<nblk(B, 32), 32, 0, pb->stream>>>(P, B, a);
    if (launch_check(pb, "k_eslice_legacy")) return -1;
    s.back(newTraj, dN, nB * per);
    s.back(n_evals, dne, nB);
    s.back(n_unif, dnu, nB);
    s.back(status, dst, nB);
    return s.finish("lna_eslice_legacy");
}

/* class Foo (generalLNA.cpp:9-372): mode 0 rate_h (out B x 3), 1 dh (B x 6, 3 x 2 column-major), 2 ODE_ones (B x 2),
 * 3 A' dh (B x 4: the product F_vec forms), 4 ODE_rk45 (B x n*3; t = the n grid times, states = initial) */
extern "C" int lna_foo(lna_problem *pb, int64_t B, int mode, int n, const double *A_pre, const double *A_post,
                       const double *param, const double *N, const double *states_or_initial, const double *t, double *out) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (mode < 0 || mode > 6) return fail("lna_foo: mode must be 0..6");
    if (!A_pre || !A_post || !param || !N || !states_or_initial || !out || (mode == 4 && (!t || n < 2)))
        return fail("lna_foo: null input");
    const size_t nB = (size_t)B, per = mode == 0 ? 3 : mode == 1 ? 6 : (mode == 2 || mode == 5) ? 2 : (mode == 3 || mode == 6) ? 4 : (size_t)n * 3;
    Stage s(pb);
    const double *dP = s.in(A_pre, nB * 6), *dQ = s.in(A_post, nB * 6), *dpar = s.in(param, nB * 3), *dN = s.in(N, nB),
                 *dx = s.in(states_or_initial, nB * 2), *dt = mode == 4 ? s.in(t, (size_t)n) : nullptr;
    double *dout = s.tmp<double>(nB * per);
    if (!s.ok) return fail("lna_foo: device staging failed");
    k_foo<<<nblk(B, 32), 32, 0, pb->stream>>>(B, mode, n, dP, dQ, dpar, dN, dx, dt, dout);
    if (launch_check(pb, "k_foo")) return -1;
    s.back(out, dout, nB * per);
    return s.finish("lna_foo");
}

// ------------------------------------------------------------------------------------------------
// data simulator: volz_thin_sir on the device (lna_thin.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int lna_volz_thin_sir(lna_problem *pb, int64_t B, int n, int64_t traj_stride, const double *t, const double *S,
                                 const double *I, const double *betaN, int ns, const double *samp_times,
                                 const int32_t *n_sampled, double t_correct, double lower, uint64_t seed, int64_t item_offset,
                                 uint64_t max_draws, int max_coal, double *coal_times, int32_t *lineages, int32_t *n_coal,
                                 uint64_t *n_draws, int32_t *status) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (!t || !S || !I || !betaN || !samp_times || !n_sampled || !coal_times || !lineages || !n_coal)
        return fail("lna_volz_thin_sir: null argument");
    if (n < 2 || ns < 1 || max_coal < 1 || !(lower > 0.0) || (traj_stride != 0 && traj_stride < n))
        return fail("lna_volz_thin_sir: bad sizes (n >= 2, ns >= 1, max_coal >= 1, lower > 0, stride 0 or >= n)");
    for (int s = 0; s < ns; ++s)
        if (n_sampled[s] < 1) return fail("lna_volz_thin_sir: n_sampled must be positive");
    for (int64_t b = 0; b < (traj_stride ? B : 1); ++b)  // R: `which.min(ts >= 0) - 1` is 0 otherwise and ts[0] stops the loop
        if (!(t[b * traj_stride + n - 1] > t_correct)) return fail("lna_volz_thin_sir: the time grid must extend beyond t_correct");
    const size_t nB = (size_t)B, len = traj_stride ? (size_t)traj_stride * nB : (size_t)n;
    Stage s(pb);
    ThinArgs a{};
    a.t = s.in(t, len);
    a.S = s.in(S, len);
    a.I = s.in(I, len);
    a.betaN = s.in(betaN, len);
    a.stride = traj_stride;
    a.n = n;
    a.samp_times = s.in(samp_times, (size_t)ns);
    a.n_sampled = s.in(n_sampled, (size_t)ns);
    a.ns = ns;
    a.t_correct = t_correct;
    a.lower = lower;
    a.seed = seed;
    a.item_offset = item_offset;
    a.coal_times = s.tmp<double>(nB * max_coal);
    a.lineages = s.tmp<int32_t>(nB * max_coal);
    a.max_coal = max_coal;
    a.n_coal = s.tmp<int32_t>(nB);
    a.status = s.tmp<int32_t>(nB);
    a.n_draws = (unsigned long long *)s.tmp<uint64_t>(nB);
    a.max_draws = max_draws ? max_draws : 40000000000ull;  // (~ a minute of one warp: a hung launch would cost the GPU box)
    if (!s.ok) return fail("lna_volz_thin_sir: device staging failed");
    if (cudaMemsetAsync(a.coal_times, 0, sizeof(double) * nB * max_coal, pb->stream) != cudaSuccess ||
        cudaMemsetAsync(a.lineages, 0, sizeof(int32_t) * nB * max_coal, pb->stream) != cudaSuccess)
        return fail("lna_volz_thin_sir: memset failed");
    k_volz_thin_sir<<<nblk(B, kThinWarps), kThinWarps * 32, 0, pb->stream>>>(B, a);
    if (launch_check(pb, "k_volz_thin_sir")) return -1;
    s.back(coal_times, a.coal_times, nB * max_coal);
    s.back(lineages, a.lineages, nB * max_coal);
    s.back(n_coal, a.n_coal, nB);
    s.back(status, a.status, nB);
    s.back(n_draws, (uint64_t *)a.n_draws, nB);
    return s.finish("lna_volz_thin_sir");
}

/* proposals first .. first + count - 1 of genealogy `item`'s stream (test replay: the CPU restatement consumes them) */
extern "C" int lna_thin_draws(lna_problem *pb, uint64_t seed, uint64_t item, uint64_t first, int64_t count, double *ua,
                              double *ub) {
    if (!pb) return fail("null problem");
    if (count <= 0) return 0;
    Stage s(pb);
    double *da = s.tmp<double>((size_t)count), *db = s.tmp<double>((size_t)count);
    if (!s.ok) return fail("lna_thin_draws: device staging failed");
    k_thin_draws<<<nblk(count, 256), 256, 0, pb->stream>>>(seed, item, first, count, da, db);
    if (launch_check(pb, "k_thin_draws")) return -1;
    s.back(ua, da, (size_t)count);
    s.back(ub, db, (size_t)count);
    return s.finish("lna_thin_draws");
}
