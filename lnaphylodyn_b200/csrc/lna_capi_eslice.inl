// lna_capi_eslice.inl -- C-ABI entry points of the Metropolis / elliptical-slice operators and of the
// resident chain batches (included at the end of lna_capi.cu).

// ------------------------------------------------------------------------------------------------
// candidate scratch (slot-major) + per-chain decision arrays
// ------------------------------------------------------------------------------------------------
struct CandBuf {
    long long NC = 0, B = 0;
    double *cA = nullptr, *cL = nullptr, *cOde = nullptr, *cBeta = nullptr, *cLat = nullptr;
    int *win_slot = nullptr, *n_evals = nullptr, *status = nullptr, *accept = nullptr;
    double *cand_ll = nullptr, *theta = nullptr;
    int *list = nullptr, *count = nullptr;  // compaction list of the two-phase slice evaluation
    double *seg_d = nullptr;                // k_ess_par_seg: parked per-chain state (lna_seg.cuh)
    int *seg_i = nullptr;
};

// slots >= base are used so that callers can keep their own staging below `base`
static bool cand_alloc(lna_problem *pb, Scratch &s, size_t base, long long B, long long NC, CandBuf &cb) {
    const DevProb &P = pb->dp;
    const size_t nA = (size_t)P.p * P.p * P.k, nO = (size_t)P.nrow * (P.p + 1), nB = (size_t)P.nrow;
    cb.B = B;
    cb.NC = NC;
    cb.cA = (double *)s.get(base + 0, sizeof(double) * nA * NC);
    cb.cL = (double *)s.get(base + 1, sizeof(double) * nA * NC);
    cb.cOde = (double *)s.get(base + 2, sizeof(double) * nO * NC);
    cb.cBeta = (double *)s.get(base + 3, sizeof(double) * nB * NC);
    cb.cLat = (double *)s.get(base + 4, sizeof(double) * nO * NC);
    cb.win_slot = (int *)s.get(base + 5, sizeof(int) * B);
    cb.n_evals = (int *)s.get(base + 6, sizeof(int) * B);
    cb.status = (int *)s.get(base + 7, sizeof(int) * B);
    cb.accept = (int *)s.get(base + 8, sizeof(int) * B);
    cb.cand_ll = (double *)s.get(base + 9, sizeof(double) * B);
    cb.theta = (double *)s.get(base + 10, sizeof(double) * B);
    cb.list = (int *)s.get(base + 11, sizeof(int) * B * 2);
    cb.count = (int *)s.get(base + 12, sizeof(int) * 256);  // [0, 24): compaction counts, [31]: k_ess_par_async's work counter, [32, 256): k_ess_par_seg's list counts
    cb.seg_d = (double *)s.get(base + 13, sizeof(double) * kSegD * B);
    cb.seg_i = (int *)s.get(base + 14, sizeof(int) * kSegI * B);
    return cb.list && cb.count && cb.cA && cb.cL && cb.cOde && cb.cBeta && cb.cLat && cb.win_slot && cb.n_evals && cb.status && cb.accept &&
           cb.cand_ll && cb.theta && cb.seg_d && cb.seg_i;
}
static constexpr size_t kCandSlots = 15;

static FullOut cand_out(const CandBuf &cb) {
    return FullOut{OutView{cb.cA, 1, cb.NC}, OutView{cb.cL, 1, cb.NC}, OutView{cb.cOde, 1, cb.NC},
                   OutView{cb.cBeta, 1, cb.NC}, OutView{cb.cLat, 1, cb.NC}};
}

static int auto_width(const lna_problem *pb, long long B) {
    // fill the machine (~512 resident FP64 threads per SM saturate the DFMA pipe) but not beyond
    const long long target = (long long)pb->sm_count * 512;
    int w = 1;
    while (w < 32 && (long long)B * (w * 2) <= target) w *= 2;
    return w;
}

template <int kProp>
static int launch_cand_one(lna_problem *pb, long long nchains, int W, const CandArgs &a) {
    const DevProb &P = pb->dp;
    const long long nthreads = nchains * W;
    const int block = (kProp == PROP_ESS_PAR || kProp == PROP_ESS_CHPT) ? 128 : pick_block(pb, nthreads);
    // Launches that leave the SM sub-partitions at most about one warp each last as long as ONE evaluation takes ONE
    // warp; there every candidate goes to a pipeline of three warps on three sub-partitions (lna_split.cuh).
    const bool split = use_split_eval(pb, nthreads);
#define LC2(WW, SPLIT)                                                                                              \
    do {                                                                                                            \
        if (SPLIT) {                                                                                                \
            const int blk = 32 * kSplitWarps; /* one pipeline of three warps per block, ~53 KB of shared memory */   \
            const size_t sm = full_eval_smem_bytes(pb->tab_bytes, blk) + split_group_bytes();                       \
            auto kern = k_cand_eval<kProp, WW, SPLIT>;                                                              \
            big_smem_attr(kern, pb->device);                                                                        \
            kern<<<nblk(nthreads, 32), blk, sm, pb->stream>>>(P, pb->seg, nchains, a);                              \
        } else {                                                                                                    \
            k_cand_eval<kProp, WW, SPLIT><<<nblk(nthreads, block), block, full_eval_smem_bytes(pb->tab_bytes, block), \
                                            pb->stream>>>(P, pb->seg, nchains, a);                                  \
        }                                                                                                           \
    } while (0)
#define LC(WW)                      \
    do {                            \
        if (split) LC2(WW, true);   \
        else LC2(WW, false);        \
    } while (0)
    // only the (proposal, width) pairs that exist are instantiated: MH kinds are one lane per chain,
    // the MH pair uses 4, the slice kinds any power of two (narrow slice launches are never latency-bound)
    constexpr bool kIsEss = (kProp == PROP_ESS_PAR || kProp == PROP_ESS_CHPT);
    if constexpr (kIsEss) {
        switch (W) {
            case 1: LC2(1, false); break;
            case 2: LC2(2, false); break;
            case 4: LC(4); break;
            case 8: LC(8); break;
            case 16: LC(16); break;
            case 32: LC(32); break;
            default: return fail("speculation width must be 1,2,4,8,16 or 32 (got %d)", W);
        }
    } else if constexpr (kProp == PROP_MH_PAIR) {
        if (W != 4) return fail("the MH pair kernel uses 4 lanes per chain");
        LC(4);
    } else if constexpr (kProp == PROP_MH_TREE) {
        switch (W) {  // 2^tm lanes per chain
            case 4: LC(4); break;
            case 8: LC(8); break;
            case 16: LC(16); break;
            default: return fail("the Metropolis tree kernel uses 4, 8 or 16 lanes per chain");
        }
    } else {
        if (W != 1) return fail("Metropolis kernels use one lane per chain");
        LC(1);
    }
#undef LC2
#undef LC
    if (timeline_on()) {  // label with <proposal kind, lanes per chain, pipelined> and the first candidate of the phase
        static std::mutex mu;
        static std::set<std::string> names;
        char buf[64];
        snprintf(buf, sizeof buf, "k_cand_eval<%d;%d;%d>@%d", kProp, W, split ? 1 : 0, a.first_cand);
        std::lock_guard<std::mutex> lock(mu);
        return launch_check(pb, names.insert(buf).first->c_str());
    }
    return launch_check(pb, "k_cand_eval");
}

// Phase plan of a slice update.  Candidates are evaluated speculatively W lanes per chain; after
// `waves` in-kernel waves the chains that are still unresolved are compacted into a dense list for the
// next phase, so no launch is dominated by idle lanes of chains that already finished.  (The number
// of candidates a chain needs is bell-shaped around ~10 for the vignette problems: tools/eslice_hist.py.)
struct SlicePhase {
    int W, first, waves;  // lanes per chain, candidates already covered, in-kernel waves (0 = until resolved)
};

static int slice_plan(const lna_problem *pb, long long B, int W, SlicePhase *ph) {
    if (W > 0) {  // explicit width: one launch, in-kernel waves
        ph[0] = {W, 0, 0};
        return 1;
    }
    if (const char *e = getenv("LNA_SLICE_PLAN")) {  // experiments (tools/plan_sweep.py): "W:first:waves,W:first:waves,..."
        int n = 0, w, f, v, used = 0;
        while (n < 24 && sscanf(e, "%d:%d:%d%n", &w, &f, &v, &used) == 3) {
            ph[n++] = {w, f, v};
            e += used;
            if (*e == ',') ++e;
        }
        if (n > 0) return n;
    }
    const int Wa = W < 0 ? -W : auto_width(pb, B);  // W < 0 forces the plan of machine-filling width -W (tests)
    switch (Wa) {
        case 32: ph[0] = {32, 0, 0}; return 1;                      // all 21 candidates in one wave
        case 16: ph[0] = {8, 0, 1}; ph[1] = {16, 8, 0}; return 2;   // 1-8, then 9-21 for the unresolved
        case 8: ph[0] = {8, 0, 1}; ph[1] = {8, 8, 1}; ph[2] = {8, 16, 0}; return 3;
        case 4: ph[0] = {4, 0, 2}; ph[1] = {4, 8, 2}; ph[2] = {4, 16, 0}; return 3;
        // With the machine full at <= 2 lanes per chain the early phases are throughput-bound; once most
        // chains are resolved the launches sit on the single-evaluation latency floor, so the tail widens.
        case 2: ph[0] = {2, 0, 2}; ph[1] = {2, 4, 2}; ph[2] = {2, 8, 2}; ph[3] = {4, 12, 1}; ph[4] = {8, 16, 0}; return 5;
        default: {
            int n = 0;
            for (int f = 0; f < 12; f += 2) ph[n++] = {1, f, 2};  // candidates 1..12, compaction every 2
            ph[n++] = {2, 12, 1};
            ph[n++] = {2, 14, 1};
            ph[n++] = {4, 16, 1};
            ph[n++] = {4, 20, 0};
            return n;
        }
    }
}

// scratch slots per chain needed by the plan (every phase keeps its own slot range so that a later
// phase never overwrites the winner of an earlier one)
static int slice_plan_slots(const lna_problem *pb, long long B, int W) {
    SlicePhase ph[24];
    const int n = slice_plan(pb, B, W, ph);
    int tot = 0;
    for (int i = 0; i < n; ++i) tot += ph[i].W;
    return std::max(tot, 4);
}

// ESlice_par_General with sequential candidates, early rejection, resumable evaluation and compaction between launches
// (lna_seg.cuh).  Measured SLOWER than the wave kernels at every batch size (65 536 chains: 10.1 vs 8.06 ms per iteration although
// it executes half the FP64 work: lanes of a warp sit in different intervals of different candidates, so candidate starts, the
// change-ratio loads and the 13 per-interval stores are uncoalesced and serialised -- a tick costs 2.6 x the wave kernel's interval,
// profiles/r2_async_slice.md), so it is OFF unless LNA_ESS_SEG=1 asks for it; kept because it is the sequential reference schedule
// on the device and is held to the same step-for-step tests.
static bool ess_seg_on(const lna_problem *pb, long long planB, const CandArgs &a0, int W_req) {
    const DevProb &P = pb->dp;
    if (!P.async_ok || !P.has_init || a0.ess[6] || W_req != 0) return false;
    (void)planB;
    const char *e = getenv("LNA_ESS_SEG");
    return e && atoi(e) != 0;
}

// ESlice_par_General of resident chains without wave barriers and with provable early rejection (lna_async.cuh):
// W lanes per chain, a free lane takes the chain's next unstarted candidate.  LNA_ESS_ASYNC=0 keeps the wave kernels,
// LNA_ESS_ASYNC_W / LNA_ESS_ASYNC_WPS force the lanes per chain / the resident warps per SM sub-partition.
// Measured with dynamic candidate assignment (tools/async_sweep2.sh, profiles/r2_async_slice.md; ms per iteration, wave kernels
// against the best barrier-free setting): 512 chains 0.376 / 0.466 (the three-warp pipeline wins), 1024: 0.510 / 0.485 (W = 16),
// 2048: 0.839 / 0.664 (16), 3072: 0.915 / 0.833 (16), 4096: 1.085 / 0.941 (8, two warps per sub-partition), 6144: 1.436 / 1.331
// (8, three), 16 384: 2.830 / 2.619 (8, three), 32 768: 4.39 / 4.82 and 65 536: 8.13 / 9.25 (throughput-bound: the wave kernels'
// cheaper ticks win).
constexpr long long kAsyncMaxWarps10 = 118;  // barrier-free kernel up to 11.8 eight-lane warps per SM sub-partition
static int ess_async_width(const lna_problem *pb, long long planB, const CandArgs &a0, int W_req, const CandBuf &cb, long long B,
                           int *wps) {
    const DevProb &P = pb->dp;
    if (!P.async_ok || !P.has_init || a0.ess[6] || W_req != 0) return 0;  // (an explicit speculation width asks for the wave kernel)
    if (const char *e = getenv("LNA_ESS_ASYNC"))
        if (atoi(e) == 0) return 0;
    int w = 8;
    *wps = 2;
    const long long subp = (long long)pb->sm_count * 4, warps8 = planB * 8 / 32, warps16 = planB * 16 / 32;
    if (const char *e = getenv("LNA_ESS_ASYNC_W")) {
        w = atoi(e);
    } else if (warps16 * 100 >= subp * 85 && warps8 * 10 < subp * 15) {
        w = 16;  // 1007 .. 3551 chains on 148 SMs: every candidate the chain is likely to need starts in the first tick
        *wps = 3;
    } else if (warps8 * 10 >= subp * 15 && warps8 * 10 <= subp * 22) {
        w = 8;   // 3552 .. 5209 chains
    } else if (warps8 * 10 > subp * 22 && warps8 * 10 <= subp * kAsyncMaxWarps10) {
        // .. 27 942 chains, ONE stream group (lna_chains_create): all warps resident while three per sub-partition hold them
        // (6144 chains: 1.331 against 1.421 ms), else persistent with two per sub-partition and groups refilling from the
        // batch (8192: 1.495 / 1.558 with three; 12 288: 2.046; 16 384: 2.603; 24 576: 3.547 against 3.645 for the wave kernels
        // on two stream groups)
        w = 8;
        *wps = warps8 <= subp * 3 ? 3 : 2;
    } else {
        return 0;
    }
    if (const char *e = getenv("LNA_ESS_ASYNC_WPS")) *wps = std::max(1, atoi(e));
    if (w != 1 && w != 2 && w != 4 && w != 8 && w != 16 && w != 32) return 0;
    if (cb.NC < B * w) return 0;
    return w;
}

template <int kProp>
static int launch_cand(lna_problem *pb, long long B, int W, const CandArgs &a0, const CandBuf &cb,
                       long long planB = 0) {
    if (planB <= 0) planB = B;
    if constexpr (kProp == PROP_ESS_PAR) {
        if (ess_seg_on(pb, planB, a0, W)) {
            // large batches: one candidate of a chain at a time, early rejection, T ticks per launch, compaction (lna_seg.cuh)
            int T = 16;
            if (const char *e = getenv("LNA_ESS_SEG_T")) T = std::max(1, atoi(e));
            const int nl = std::min(220, ((20 + 1) * pb->dp.k + T - 1) / T + 1);
            const size_t sm = full_eval_smem_bytes(pb->tab_bytes, 128) + sizeof(double) * 128 * (size_t)(kMaxUnif + 9);
            if (cudaMemsetAsync(cb.seg_i, 0, sizeof(int) * kSegI * (size_t)B, pb->stream) != cudaSuccess ||
                cudaMemsetAsync(cb.count + 32, 0, sizeof(int) * 224, pb->stream) != cudaSuccess)
                return fail("memset failed");
            const SegState ss{cb.seg_d, cb.seg_i};
            for (int l = 0; l < nl; ++l) {
                const int *lin = l == 0 ? nullptr : cb.list + (size_t)((l - 1) & 1) * B, *cin = l == 0 ? nullptr : cb.count + 32 + (l - 1);
                k_ess_par_seg<<<nblk(B, 128), 128, sm, pb->stream>>>(pb->dp, pb->seg, B, a0, ss, lin, cin, cb.list + (size_t)(l & 1) * B,
                                                                     cb.count + 32 + l, T);
                if (launch_check(pb, "k_ess_par_seg")) return -1;
            }
            return 0;
        }
        int wps = 2;
        const int wa = ess_async_width(pb, planB, a0, W, cb, B, &wps);
        if (wa) {
            // persistent: at most `wps` warps per SM sub-partition, groups of wa lanes pull chains from a counter
            const long long warps_needed = (B * wa + 31) / 32, warps_cap = (long long)pb->sm_count * 4 * wps;
            const long long warps = std::min(warps_needed, warps_cap);
            const unsigned grid = (unsigned)((warps + 3) / 4);
            // tables + per-thread eps slots + per-group staging of the chain's scalars, uniforms and change ratios
            const size_t sm = full_eval_smem_bytes(pb->tab_bytes, 128) +
                              sizeof(double) * (size_t)(128 / wa) * (size_t)lna::async_group_doubles(pb->dp.nch);
            if (sm > 48 * 1024) {
                if (sm > 200 * 1024) return fail("k_ess_par_async: %d change points do not fit in shared memory", pb->dp.nch);
                big_smem_attr(k_ess_par_async<1>, pb->device);
                big_smem_attr(k_ess_par_async<2>, pb->device);
                big_smem_attr(k_ess_par_async<4>, pb->device);
                big_smem_attr(k_ess_par_async<8>, pb->device);
                big_smem_attr(k_ess_par_async<16>, pb->device);
                big_smem_attr(k_ess_par_async<32>, pb->device);
            }
            CandArgs aa = a0;
            aa.work_counter = cb.count + 31;
            aa.max_waves = getenv("LNA_ASYNC_COUNT_LANES") ? 1 : 0;  // (experiments: exec_count[1] accumulates lane-evaluations)
            if (cudaMemsetAsync(aa.work_counter, 0, sizeof(int), pb->stream) != cudaSuccess) return fail("memset failed");
            switch (wa) {
                case 1: k_ess_par_async<1><<<grid, 128, sm, pb->stream>>>(pb->dp, pb->seg, B, aa); break;
                case 2: k_ess_par_async<2><<<grid, 128, sm, pb->stream>>>(pb->dp, pb->seg, B, aa); break;
                case 4: k_ess_par_async<4><<<grid, 128, sm, pb->stream>>>(pb->dp, pb->seg, B, aa); break;
                case 8: k_ess_par_async<8><<<grid, 128, sm, pb->stream>>>(pb->dp, pb->seg, B, aa); break;
                case 16: k_ess_par_async<16><<<grid, 128, sm, pb->stream>>>(pb->dp, pb->seg, B, aa); break;
                default: k_ess_par_async<32><<<grid, 128, sm, pb->stream>>>(pb->dp, pb->seg, B, aa); break;
            }
            return launch_check(pb, "k_ess_par_async");
        }
    }
    CandArgs a = a0;
    a.first_cand = 0;
    a.slot_base = 0;
    a.max_waves = 0;
    a.chain_list = nullptr;
    a.n_list = nullptr;
    a.next_list = cb.list;
    a.next_count = cb.count;
    constexpr bool kIsEss = (kProp == PROP_ESS_PAR || kProp == PROP_ESS_CHPT);
    if (!kIsEss) return launch_cand_one<kProp>(pb, B, W > 0 ? W : 1, a);
    SlicePhase ph[24];
    int nph = slice_plan(pb, planB, W, ph);
    if (kProp == PROP_ESS_PAR && a0.ess[6]) {
        // self-mixing OriginTraj (ESS_vec[6] = 1): candidate i's increments depend on which earlier candidates failed their
        // factorisation, so the candidates of a chain are evaluated one per wave, in order, in ONE launch
        ph[0] = {1, 0, 0};
        nph = 1;
    }
    if (nph > 1 && cudaMemsetAsync(cb.count, 0, sizeof(int) * 24, pb->stream) != cudaSuccess)
        return fail("memset failed");
    long long slot = 0;
    for (int i = 0; i < nph; ++i) {
        a.first_cand = ph[i].first;
        a.max_waves = ph[i].waves;
        a.slot_base = slot;
        a.chain_list = i == 0 ? nullptr : cb.list + (size_t)((i - 1) & 1) * B;  // ping-pong compaction lists
        a.n_list = i == 0 ? nullptr : cb.count + (i - 1);
        a.next_list = cb.list + (size_t)(i & 1) * B;
        a.next_count = cb.count + i;
        if (launch_cand_one<kProp>(pb, B, ph[i].W, a)) return -1;
        slot += B * ph[i].W;
    }
    return 0;
}

static int launch_commit_fields(lna_problem *pb, long long B, const CandBuf &cb, OutView A, OutView L, OutView ode,
                                OutView betaN, OutView latent) {
    const DevProb &P = pb->dp;
    CommitArgs ca{cb.win_slot, cb.NC, cb.cA, cb.cL, cb.cOde, cb.cBeta, cb.cLat, A, L, ode, betaN, latent,
                  P.p * P.p * P.k, P.nrow * (P.p + 1), P.nrow, 0};
    const int per = 2 * ca.nA + 2 * ca.nOde + ca.nBeta;
    const int ny = (per + lna::kCommitPer - 1) / lna::kCommitPer;
    if (ny > 65535) return fail("k_commit_fields: %d elements per chain exceed the grid.y limit", per);
    k_commit_fields<<<dim3(nblk(B, 128), ny), 128, 0, pb->stream>>>(B, ca);
    return launch_check(pb, "k_commit_fields");
}

static int launch_traj(lna_problem *pb, long long B, const TrajArgs &a) {
    const DevProb &P = pb->dp;
    const size_t sh = pb->tab_bytes + sizeof(double) * kTrajWarps * traj_warp_doubles(P.k, P.nrow);
    if (sh > 48 * 1024) {
        if (sh > 200 * 1024) return fail("k_eslice_traj: k = %d LNA intervals do not fit in shared memory", P.k);
        big_smem_attr(k_eslice_traj, pb->device);  // (once per device: not repeated inside a stream capture)
    }
    k_eslice_traj<<<nblk(B, kTrajWarps), kTrajWarps * 32, sh, pb->stream>>>(P, B, a);
    return launch_check(pb, "k_eslice_traj");
}

static int check_sir_sampler(const lna_problem *pb, const char *what) {
    if (pb->dp.model != LNA_MODEL_SIR || pb->dp.nparam != 2 || pb->dp.iR0 != 0 || pb->dp.iGamma != 1)
        return fail("%s: the reference's samplers are hard-wired to SIR (p = 2, nparam = 2, x_i = (nch, 2, 0, 1))", what);
    if (!pb->dp.has_init) return fail("%s: problem was created without a genealogy", what);
    if (pb->dp.transX != LNA_TRANSX_STANDARD)
        return fail("%s: the Metropolis / slice operators work on the standard scale (transX = \"standard\"), like every "
                    "driver of the reference that calls them", what);
    return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" int lna_param_slice_update(lna_problem *pb, int64_t B, const double *param, const double *theta,
                                      const double *newChs, double *param_new) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B;
    Stage s(pb);
    const double *dp = s.in(param, nB * P.npt), *dt = s.in(theta, nB), *dn = s.in(newChs, nB * P.nch);
    double *dout = s.out(param_new, nB * P.npt);
    if (!s.ok || !dp || !dt || !dn || !dout) return fail("lna_param_slice_update: bad arguments / staging failed");
    k_param_slice_update<<<nblk(B, 128), 128, 0, pb->stream>>>(P, B, item_view(dp, P.npt), dt, item_view(dn, P.nch), dout);
    if (launch_check(pb, "k_param_slice_update")) return -1;
    s.back(param_new, dout, nB * P.npt);
    return s.finish("lna_param_slice_update");
}

extern "C" int lna_param_slice_update_all2(lna_problem *pb, int64_t B, const double *par_old, const double *theta,
                                           const double *newChs, const int32_t *ESS_vec7, const double *prior8,
                                           double *par_new) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (P.model != LNA_MODEL_SIR) return fail("lna_param_slice_update_all2: SIR only (hard-wired in the reference)");
    if (!ESS_vec7 || !prior8) return fail("lna_param_slice_update_all2: null ESS_vec / prior");
    const size_t nB = (size_t)B, npar = 2 + P.npt;
    Stage s(pb);
    const double *dp = s.in(par_old, nB * npar), *dt = s.in(theta, nB), *dn = s.in(newChs, nB * (npar - 1));
    double *dout = s.out(par_new, nB * npar);
    if (!s.ok || !dp || !dt || !dn || !dout) return fail("lna_param_slice_update_all2: bad arguments / staging failed");
    CandArgs a{};
    a.par = item_view(dp, npar);
    a.normals = item_view(dn, npar - 1);
    for (int i = 0; i < 8; ++i) a.pr[i] = prior8[i];
    for (int i = 0; i < 7; ++i) a.ess[i] = ESS_vec7[i];
    k_param_slice_update_all2<<<nblk(B, 128), 128, 0, pb->stream>>>(P, B, a, dt, dout);
    if (launch_check(pb, "k_param_slice_update_all2")) return -1;
    s.back(par_new, dout, nB * npar);
    return s.finish("lna_param_slice_update_all2");
}

// ------------------------------------------------------------------------------------------------
// logOrigin and OriginTraj_new of the self-mixing branch: OriginTraj_new = scale * OriginTraj (:1660-1662, 1673-1678)
__global__ void k_scale_origin(long long B, int n, const double *origin, const double *scale, double *origin_new, double *logOrigin) {
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= B) return;
    const double sc = scale[c];
    double lo = 0.0;
    for (int e = lane; e < n; e += 32) {
        const double v = origin[c * n + e] * sc;
        if (origin_new) origin_new[c * n + e] = v;
        lo -= 0.5 * v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo += __shfl_xor_sync(0xffffffffu, lo, o);
    if (lane == 0 && logOrigin) logOrigin[c] = lo;
}

extern "C" int lna_eslice_par_general_ex(lna_problem *pb, int64_t B, const double *par_old, const double *OriginTraj,
                                         const double *prior8, const int32_t *ESS_vec7, const double *coal_log,
                                         const double *normals, const double *uniforms, int spec_width, double *par_new,
                                         double *Acube, double *Lcube, double *Ode_coarse, double *betaN,
                                         double *LatentTraj, double *CoalLog, double *logOrigin, int32_t *n_evals,
                                         int32_t *status, double *OriginTraj_new) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (check_sir_sampler(pb, "lna_eslice_par_general")) return -1;
    if (!par_old || !OriginTraj || !prior8 || !ESS_vec7 || !coal_log || !normals || !uniforms)
        return fail("lna_eslice_par_general: null input");
    const bool selfmix = ESS_vec7[6] != 0;
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B, npar = 2 + P.npt, kp = (size_t)P.k * P.p;
    const size_t nA = (size_t)P.p * P.p * P.k, nO = (size_t)P.nrow * (P.p + 1), nrow = P.nrow;
    const int W = spec_width;
    Stage s(pb);
    CandArgs a{};
    a.par = item_view(s.in(par_old, nB * npar), npar);
    a.origin = item_view(s.in(OriginTraj, nB * kp), kp);
    a.normals = item_view(s.in(normals, nB * (npar - 1)), npar - 1);
    a.uniforms = item_view(s.in(uniforms, nB * kMaxUnif), kMaxUnif);
    a.coal_log = s.in(coal_log, nB);
    for (int i = 0; i < 8; ++i) a.pr[i] = prior8[i];
    for (int i = 0; i < 7; ++i) a.ess[i] = ESS_vec7[i];
    double *dpar = s.tmp<double>(nB * npar), *dA = s.tmp<double>(nB * nA), *dL = s.tmp<double>(nB * nA);
    double *dO = s.tmp<double>(nB * nO), *dbn = s.tmp<double>(nB * nrow), *dlat = s.tmp<double>(nB * nO);
    double *dlo = s.tmp<double>(nB), *dcl = s.tmp<double>(nB);
    double *dsc = selfmix ? s.tmp<double>(nB) : nullptr, *don = (selfmix && OriginTraj_new) ? s.tmp<double>(nB * kp) : nullptr;
    CandBuf cb;
    if (!s.ok || !cand_alloc(pb, pb->scratch, s.slot, B, (long long)B * slice_plan_slots(pb, B, W), cb)) return fail("lna_eslice_par_general: device staging failed");
    a.eps_scale = dsc;
    a.cand = cand_out(cb);
    a.win_slot = cb.win_slot;
    a.cand_ll = cb.cand_ll;
    a.theta = cb.theta;
    a.n_evals = cb.n_evals;
    a.status = cb.status;
    a.accept = cb.accept;
    if (launch_cand<PROP_ESS_PAR>(pb, B, W, a, cb)) return -1;
    if (launch_commit_fields(pb, B, cb, item_out(dA, nA), item_out(dL, nA), item_out(dO, nO), item_out(dbn, nrow),
                             item_out(dlat, nO)))
        return -1;
    k_commit_par_wide<PROP_ESS_PAR><<<nblk((long long)(nB * npar), 256), 256, 0, pb->stream>>>(P, B, a, item_out(dpar, npar), dcl);
    if (launch_check(pb, "k_commit_par_wide")) return -1;
    if (selfmix) {
        k_scale_origin<<<nblk((long long)B * 32, 128), 128, 0, pb->stream>>>(B, (int)kp, a.origin.base, dsc, don, dlo);
        if (launch_check(pb, "k_scale_origin")) return -1;
    } else {
        k_log_origin<<<nblk((long long)B * 32, 128), 128, 0, pb->stream>>>(B, (int)kp, a.origin, dlo);
        if (launch_check(pb, "k_log_origin")) return -1;
    }
    s.back(par_new, dpar, nB * npar);
    s.back(Acube, dA, nB * nA);
    s.back(Lcube, dL, nB * nA);
    s.back(Ode_coarse, dO, nB * nO);
    s.back(betaN, dbn, nB * nrow);
    s.back(LatentTraj, dlat, nB * nO);
    s.back(CoalLog, cb.cand_ll, nB);
    s.back(logOrigin, dlo, nB);
    s.back(n_evals, cb.n_evals, nB);
    s.back(status, cb.status, nB);
    if (OriginTraj_new) {
        if (selfmix) s.back(OriginTraj_new, don, nB * kp);
        else s.back(OriginTraj_new, a.origin.base, nB * kp);  // OriginTraj_new = OriginTraj (:1580)
    }
    return s.finish("lna_eslice_par_general");
}

extern "C" int lna_eslice_par_general(lna_problem *pb, int64_t B, const double *par_old, const double *OriginTraj,
                                      const double *prior8, const int32_t *ESS_vec7, const double *coal_log,
                                      const double *normals, const double *uniforms, int spec_width, double *par_new,
                                      double *Acube, double *Lcube, double *Ode_coarse, double *betaN,
                                      double *LatentTraj, double *CoalLog, double *logOrigin, int32_t *n_evals,
                                      int32_t *status) {
    return lna_eslice_par_general_ex(pb, B, par_old, OriginTraj, prior8, ESS_vec7, coal_log, normals, uniforms, spec_width,
                                     par_new, Acube, Lcube, Ode_coarse, betaN, LatentTraj, CoalLog, logOrigin, n_evals, status,
                                     nullptr);
}

extern "C" int lna_eslice_change_points(lna_problem *pb, int64_t B, const double *param, const double *initial,
                                        const double *OriginTraj, const double *coal_log, const double *normals,
                                        const double *uniforms, int spec_width, double *param_new, double *Acube,
                                        double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj,
                                        double *CoalLog, double *LogChProb, int32_t *n_evals, int32_t *status) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (check_sir_sampler(pb, "lna_eslice_change_points")) return -1;
    if (!param || !initial || !OriginTraj || !coal_log || !normals || !uniforms)
        return fail("lna_eslice_change_points: null input");
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B, npar = 2 + P.npt, kp = (size_t)P.k * P.p;
    const size_t nA = (size_t)P.p * P.p * P.k, nO = (size_t)P.nrow * (P.p + 1), nrow = P.nrow;
    const int W = spec_width;
    // par = c(initial, param): assemble on the host (small)
    std::vector<double> par(nB * npar);
    for (size_t b = 0; b < nB; ++b) {
        par[b * npar] = initial[b * 2];
        par[b * npar + 1] = initial[b * 2 + 1];
        std::memcpy(&par[b * npar + 2], param + b * P.npt, sizeof(double) * P.npt);
    }
    Stage s(pb);
    CandArgs a{};
    a.par = item_view(s.in(par.data(), nB * npar), npar);
    a.origin = item_view(s.in(OriginTraj, nB * kp), kp);
    a.normals = item_view(s.in(normals, nB * P.nch), P.nch);
    a.uniforms = item_view(s.in(uniforms, nB * kMaxUnif), kMaxUnif);
    a.coal_log = s.in(coal_log, nB);
    double *dpar = s.tmp<double>(nB * npar), *dA = s.tmp<double>(nB * nA), *dL = s.tmp<double>(nB * nA);
    double *dO = s.tmp<double>(nB * nO), *dbn = s.tmp<double>(nB * nrow), *dlat = s.tmp<double>(nB * nO);
    double *dlcp = s.tmp<double>(nB), *dcl = s.tmp<double>(nB);
    CandBuf cb;
    if (!s.ok || !cand_alloc(pb, pb->scratch, s.slot, B, (long long)B * slice_plan_slots(pb, B, W), cb))
        return fail("lna_eslice_change_points: device staging failed");
    // the host vector must outlive the async copy
    if (cudaStreamSynchronize(pb->stream) != cudaSuccess) return fail("lna_eslice_change_points: sync failed");
    a.cand = cand_out(cb);
    a.win_slot = cb.win_slot;
    a.cand_ll = cb.cand_ll;
    a.theta = cb.theta;
    a.n_evals = cb.n_evals;
    a.status = cb.status;
    a.accept = cb.accept;
    if (launch_cand<PROP_ESS_CHPT>(pb, B, W, a, cb)) return -1;
    if (launch_commit_fields(pb, B, cb, item_out(dA, nA), item_out(dL, nA), item_out(dO, nO), item_out(dbn, nrow),
                             item_out(dlat, nO)))
        return -1;
    k_commit_par_wide<PROP_ESS_CHPT><<<nblk((long long)(nB * npar), 256), 256, 0, pb->stream>>>(P, B, a, item_out(dpar, npar), dcl);
    if (launch_check(pb, "k_commit_par_wide")) return -1;
    k_log_chprob<<<nblk((long long)B * 32, 128), 128, 0, pb->stream>>>(B, P.nch, (int)npar, item_view(dpar, npar), dlcp);
    if (launch_check(pb, "k_log_chprob")) return -1;
    std::vector<double> par_out(param_new ? nB * npar : 0);
    s.back(param_new ? par_out.data() : nullptr, dpar, nB * npar);
    s.back(Acube, dA, nB * nA);
    s.back(Lcube, dL, nB * nA);
    s.back(Ode_coarse, dO, nB * nO);
    s.back(betaN, dbn, nB * nrow);
    s.back(LatentTraj, dlat, nB * nO);
    s.back(CoalLog, cb.cand_ll, nB);
    s.back(LogChProb, dlcp, nB);
    s.back(n_evals, cb.n_evals, nB);
    s.back(status, cb.status, nB);
    if (s.finish("lna_eslice_change_points")) return -1;
    if (param_new)
        for (size_t b = 0; b < nB; ++b) std::memcpy(param_new + b * P.npt, &par_out[b * npar + 2], sizeof(double) * P.npt);
    return 0;
}

static int eslice_general_nc_impl(lna_problem *pb, int64_t B, const double *f_cur, const double *Ode_coarse,
                                  const double *Acube, const double *Lcube, const double *betaN,
                                  const double *coal_log, const double *normals, const double *uniforms,
                                  double *LatentTraj, double *OriginTraj_new, double *logOrigin, double *CoalLog,
                                  int32_t *n_evals, int32_t *status, const double *lambda);

extern "C" int lna_eslice_general_nc(lna_problem *pb, int64_t B, const double *f_cur, const double *Ode_coarse,
                                     const double *Acube, const double *Lcube, const double *betaN,
                                     const double *coal_log, const double *normals, const double *uniforms,
                                     double *LatentTraj, double *OriginTraj_new, double *logOrigin, double *CoalLog,
                                     int32_t *n_evals, int32_t *status) {
    return eslice_general_nc_impl(pb, B, f_cur, Ode_coarse, Acube, Lcube, betaN, coal_log, normals, uniforms, LatentTraj,
                                  OriginTraj_new, logOrigin, CoalLog, n_evals, status, nullptr);
}

extern "C" int lna_eslice_general_nc_kingman(lna_problem *pb, int64_t B, const double *f_cur, const double *Ode_coarse,
                                             const double *Acube, const double *Lcube, const double *betaN,
                                             const double *coal_log, const double *lambda, const double *normals,
                                             const double *uniforms, double *LatentTraj, double *OriginTraj_new,
                                             double *logOrigin, double *CoalLog, int32_t *n_evals, int32_t *status) {
    if (!lambda || !coal_log) return fail("lna_eslice_general_nc_kingman: null lambda / coal_log");
    for (int64_t b = 0; b < B; ++b)
        if (coal_log[b] == -99999999.0)
            return fail("lna_eslice_general_nc_kingman: coal_log = -99999999 makes the reference hand the increment matrix to "
                        "coal_loglik3 as a trajectory (LNA_functional.cpp:1720); pass the current likelihood");
    return eslice_general_nc_impl(pb, B, f_cur, Ode_coarse, Acube, Lcube, betaN, coal_log, normals, uniforms, LatentTraj,
                                  OriginTraj_new, logOrigin, CoalLog, n_evals, status, lambda);
}

static int eslice_general_nc_impl(lna_problem *pb, int64_t B, const double *f_cur, const double *Ode_coarse,
                                  const double *Acube, const double *Lcube, const double *betaN,
                                  const double *coal_log, const double *normals, const double *uniforms,
                                  double *LatentTraj, double *OriginTraj_new, double *logOrigin, double *CoalLog,
                                  int32_t *n_evals, int32_t *status, const double *lambda) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (check_sir_sampler(pb, "lna_eslice_general_nc")) return -1;
    if (!f_cur || !Ode_coarse || !Acube || !Lcube || !betaN || !coal_log || !normals || !uniforms)
        return fail("lna_eslice_general_nc: null input");
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B, kp = (size_t)P.k * P.p;
    const size_t nA = (size_t)P.p * P.p * P.k, nO = (size_t)P.nrow * (P.p + 1), nrow = P.nrow;
    Stage s(pb);
    TrajArgs a{};
    a.f_cur = item_view(s.in(f_cur, nB * kp), kp);
    a.normals = item_view(s.in(normals, nB * kp), kp);
    a.uniforms = item_view(s.in(uniforms, nB * kMaxUnif), kMaxUnif);
    a.A = item_view(s.in(Acube, nB * nA), nA);
    a.L = item_view(s.in(Lcube, nB * nA), nA);
    a.ode = item_view(s.in(Ode_coarse, nB * nO), nO);
    a.betaN = item_view(s.in(betaN, nB * nrow), nrow);
    a.coal_log = s.in(coal_log, nB);
    a.kingman = lambda != nullptr;
    a.lambda = s.in(lambda, nB);
    double *dlat = s.tmp<double>(nB * nO), *dor = s.tmp<double>(nB * kp), *dcl = s.tmp<double>(nB), *dlo = s.tmp<double>(nB);
    int32_t *dne = s.tmp<int32_t>(nB), *dst = s.tmp<int32_t>(nB);
    if (!s.ok) return fail("lna_eslice_general_nc: device staging failed");
    a.latent = item_out(dlat, nO);
    a.origin_new = item_out(dor, kp);
    a.CoalLog = dcl;
    a.logOrigin = dlo;
    a.n_evals = dne;
    a.status = dst;
    if (launch_traj(pb, B, a)) return -1;
    s.back(LatentTraj, dlat, nB * nO);
    s.back(OriginTraj_new, dor, nB * kp);
    s.back(logOrigin, dlo, nB);
    s.back(CoalLog, dcl, nB);
    s.back(n_evals, dne, nB);
    s.back(status, dst, nB);
    return s.finish("lna_eslice_general_nc");
}

// ------------------------------------------------------------------------------------------------
// centred parametrisation
// ------------------------------------------------------------------------------------------------
static int centred_launch(lna_problem *pb, int64_t B, int mode, const double *x, const double *O, const double *A,
                          const double *S, double *traj, double *ll, int32_t *st, const double *initial = nullptr) {
    const DevProb &P = pb->dp;
    if (P.p == 2)
        k_centred<2><<<nblk(B, 64), 64, 0, pb->stream>>>(P.k, B, mode, P.t_correct, x, O, A, S, traj, ll, st, initial);
    else
        k_centred<3><<<nblk(B, 64), 64, 0, pb->stream>>>(P.k, B, mode, P.t_correct, x, O, A, S, traj, ll, st, initial);
    return launch_check(pb, "k_centred");
}

extern "C" int lna_log_like_traj_general2(lna_problem *pb, int64_t B, const double *SdeTraj, const double *OdeTraj,
                                          const double *Acube, const double *Scube, double *loglik) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B, per = (size_t)P.nrow * (P.p + 1), cube = (size_t)P.p * P.p * P.k;
    Stage s(pb);
    const double *dX = s.in(SdeTraj, nB * per), *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube),
                 *dS = s.in(Scube, nB * cube);
    double *dll = s.tmp<double>(nB);
    if (!s.ok || !dX || !dO || !dA || !dS) return fail("lna_log_like_traj_general2: bad arguments / staging failed");
    if (centred_launch(pb, B, 0, dX, dO, dA, dS, nullptr, dll, nullptr)) return -1;
    s.back(loglik, dll, nB);
    return s.finish("lna_log_like_traj_general2");
}

/* log_like_trajSIRS (SIRS_period.cpp:82-118): pseudo-inverse density of the rank-2 seasonal SIRS covariance */
extern "C" int lna_log_like_traj_sirs(lna_problem *pb, int64_t B, const double *SdeTraj, const double *OdeTraj,
                                      const double *Acube, const double *Scube, double *loglik) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (P.p != 3) return fail("lna_log_like_traj_sirs: needs a 3-compartment problem");
    const size_t nB = (size_t)B, per = (size_t)P.nrow * (P.p + 1), cube = (size_t)P.p * P.p * P.k;
    Stage s(pb);
    const double *dX = s.in(SdeTraj, nB * per), *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube),
                 *dS = s.in(Scube, nB * cube);
    double *dll = s.tmp<double>(nB);
    if (!s.ok || !dX || !dO || !dA || !dS) return fail("lna_log_like_traj_sirs: bad arguments / staging failed");
    if (centred_launch(pb, B, 2, dX, dO, dA, dS, nullptr, dll, nullptr)) return -1;
    s.back(loglik, dll, nB);
    return s.finish("lna_log_like_traj_sirs");
}

extern "C" int lna_traj_sim_general3(lna_problem *pb, int64_t B, const double *OdeTraj, const double *Acube,
                                     const double *Scube, const double *normals, double *LNA_traj, double *loglike,
                                     int32_t *status) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B, per = (size_t)P.nrow * (P.p + 1), cube = (size_t)P.p * P.p * P.k;
    Stage s(pb);
    const double *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube), *dS = s.in(Scube, nB * cube),
                 *dz = s.in(normals, nB * P.k * P.p);
    double *dT = s.tmp<double>(nB * per), *dll = s.tmp<double>(nB);
    int32_t *dst = s.tmp<int32_t>(nB);
    if (!s.ok || !dO || !dA || !dS || !dz) return fail("lna_traj_sim_general3: bad arguments / staging failed");
    if (centred_launch(pb, B, 1, dz, dO, dA, dS, dT, dll, dst)) return -1;
    s.back(LNA_traj, dT, nB * per);
    s.back(loglike, dll, nB);
    s.back(status, dst, nB);
    return s.finish("lna_traj_sim_general3");
}

/* Traj_sim_SIRS (SIRS_period.cpp:197-240) */
extern "C" int lna_traj_sim_sirs(lna_problem *pb, int64_t B, const double *initial, const double *OdeTraj, const double *Acube,
                                 const double *Scube, const double *normals, double *LNA_traj, double *loglike,
                                 int32_t *status) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    const DevProb &P = pb->dp;
    if (P.p != 3) return fail("lna_traj_sim_sirs: needs a 3-compartment problem");
    const size_t nB = (size_t)B, per = (size_t)P.nrow * 4, cube = 9 * (size_t)P.k;
    Stage s(pb);
    const double *dI = s.in(initial, nB * 3), *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube),
                 *dS = s.in(Scube, nB * cube), *dz = s.in(normals, nB * P.k * 3);
    double *dT = s.tmp<double>(nB * per), *dll = s.tmp<double>(nB);
    int32_t *dst = s.tmp<int32_t>(nB);
    if (!s.ok || !dI || !dO || !dA || !dS || !dz) return fail("lna_traj_sim_sirs: bad arguments / staging failed");
    if (centred_launch(pb, B, 3, dz, dO, dA, dS, dT, dll, dst, dI)) return -1;
    s.back(LNA_traj, dT, nB * per);
    s.back(loglike, dll, nB);
    s.back(status, dst, nB);
    return s.finish("lna_traj_sim_sirs");
}

extern "C" int lna_eslice_general2(lna_problem *pb, int64_t B, const double *f_cur, const double *OdeTraj,
                                   const double *Acube, const double *Scube, const double *betaN,
                                   const double *normals, const double *uniforms, double *newTraj, int32_t *n_evals,
                                   int32_t *status) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (check_sir_sampler(pb, "lna_eslice_general2")) return -1;
    if (!f_cur || !OdeTraj || !Acube || !Scube || !betaN || !normals || !uniforms)
        return fail("lna_eslice_general2: null input");
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)B, per = (size_t)P.nrow * (P.p + 1), cube = (size_t)P.p * P.p * P.k, nrow = P.nrow;
    Stage s(pb);
    const double *dF = s.in(f_cur, nB * per), *dO = s.in(OdeTraj, nB * per), *dA = s.in(Acube, nB * cube),
                 *dS = s.in(Scube, nB * cube), *db = s.in(betaN, nB * nrow), *dz = s.in(normals, nB * P.k * P.p),
                 *du = s.in(uniforms, nB * kMaxUnif);
    double *dV = s.tmp<double>(nB * per), *dN = s.tmp<double>(nB * per), *dll = s.tmp<double>(nB);
    int32_t *dne = s.tmp<int32_t>(nB), *dst = s.tmp<int32_t>(nB), *dst2 = s.tmp<int32_t>(nB);
    if (!s.ok) return fail("lna_eslice_general2: device staging failed");
    // v = Traj_sim_general3(OdeTraj, FTs, t_correct) consumes the k*p normals first (:1801)
    if (centred_launch(pb, B, 1, dz, dO, dA, dS, dV, dll, dst2)) return -1;
    TrajArgs a{};
    a.mode = 1;
    a.f_cur = item_view(dF, per);
    a.normals = item_view(dV, per);
    a.uniforms = item_view(du, kMaxUnif);
    a.ode = item_view(dO, per);
    a.betaN = item_view(db, nrow);
    a.A = item_view(dA, cube);
    a.L = item_view(dS, cube);
    a.coal_log = dll;  // unused in mode 1
    a.latent = item_out(dN, per);
    a.origin_new = OutView{nullptr, 0, 0};
    a.CoalLog = nullptr;
    a.logOrigin = nullptr;
    a.n_evals = dne;
    a.status = dst;
    if (launch_traj(pb, B, a)) return -1;
    s.back(newTraj, dN, nB * per);
    s.back(n_evals, dne, nB);
    s.back(status, dst, nB);
    return s.finish("lna_eslice_general2");
}

// ------------------------------------------------------------------------------------------------
// resident chain batches
// ------------------------------------------------------------------------------------------------
struct ChainPart {  // one stream-group of a resident chain batch
    lna_problem *pb = nullptr;  // this group's problem (own stream + scratch); == the user's problem when G == 1
    long long B = 0;
    long long planB = 0;  // chain count of the WHOLE batch: phase plans and policies are chosen for it
    int W = 1;
    // slot-major state: field[j*B + c]
    double *par_home = nullptr;  // the buffer `par` points to between iterations (see chains_finish)
    double *par = nullptr, *par_alt = nullptr, *origin = nullptr, *A = nullptr, *L = nullptr, *ode = nullptr, *betaN = nullptr,
           *latent = nullptr, *coalLog = nullptr, *logOrigin = nullptr, *summary = nullptr;
    int *counters = nullptr;  // [4][B]: n_full, n_cheap, n_mh_accept, status
    int *tmp_i = nullptr;     // [2][B]
    double *tmp_d = nullptr;  // [B]
    CandBuf cb;
    Scratch cand_scratch;
    std::vector<void *> owned;
    double *stage_n = nullptr, *stage_u = nullptr;  // staging for host-pointer steps
    size_t n_norm = 0, n_unif = 0;
    unsigned long long *exec = nullptr;      // [2] executed evaluations / warp-evaluations (CandArgs::exec_count)
    double *jT = nullptr, *jnew = nullptr;  // joint Metropolis step: [B][16] transforms (padded 4 x 4), [B][4] proposed values
    int jd = 0;
    // CUDA graphs of whole iterations on device-drawn streams (one per distinct option set): the ~14 launches of an
    // iteration are replayed with one cudaGraphLaunch; only the Philox counter (iteration) changes between replays
    struct StepGraph {
        std::vector<unsigned char> key;
        cudaGraph_t graph = nullptr;  // kept alive: draw_node is a handle into it
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t draw_node = nullptr;
        cudaKernelNodeParams draw_params{};
        int n_kernels = 0;
        bool broken = false;
        // argument block of the k_draw_streams node
        long long aB = 0, aoff = 0;
        int ann = 0, anu = 0;
        unsigned long long aseed = 0, aiter = 0;
        double *adn = nullptr, *adu = nullptr;
        void *argv[8] = {};
    };
    std::vector<StepGraph *> graphs;
    std::vector<unsigned char> last_key;
};

__global__ void k_bump(long long B, const int *n_evals, const int *accept, const int *status, const int *win_slot,
                       int *counters, int which_evals, int count_accept) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    counters[(long long)which_evals * B + c] += n_evals[c];
    if (count_accept) counters[2 * B + c] += __popc(accept[c] & 15);
    // NEG/NAN bits of a proposal that was not adopted say nothing about the chain's state
    const int adopted = win_slot ? (win_slot[c] >= 0) : 1;
    counters[3 * B + c] |= adopted ? status[c] : (status[c] & (ST_CHOL_FAIL | ST_CAP_HIT));
}

__global__ void k_summary(long long B, const double *coalLog, const double *logOrigin, const int *counters,
                          double *summary) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    summary[c * 4 + 0] = coalLog[c];
    summary[c * 4 + 1] = logOrigin[c];
    summary[c * 4 + 2] = (double)counters[c];
    summary[c * 4 + 3] = (double)counters[3 * B + c];
}

static void part_destroy(ChainPart *ch);

template <class T>
static T *chains_alloc(ChainPart *ch, size_t n) {
    T *d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    ch->owned.push_back(d);
    return d;
}

static ChainPart *part_create(lna_problem *pb, int64_t B, int64_t planB, double *summary) {
    if (!pb || B <= 0) {
        fail("lna_chains_create: bad arguments");
        return nullptr;
    }
    if (check_sir_sampler(pb, "lna_chains_create")) return nullptr;
    cudaSetDevice(pb->device);
    const DevProb &P = pb->dp;
    ChainPart *ch = new ChainPart();
    ch->pb = pb;
    ch->B = B;
    ch->planB = planB;
    const size_t nB = (size_t)B, npar = 2 + P.npt, kp = (size_t)P.k * P.p;
    const size_t nA = (size_t)P.p * P.p * P.k, nO = (size_t)P.nrow * (P.p + 1), nrow = P.nrow;
    ch->par = chains_alloc<double>(ch, nB * npar);
    ch->par_home = ch->par;
    ch->par_alt = chains_alloc<double>(ch, nB * npar);
    ch->origin = chains_alloc<double>(ch, nB * kp);
    ch->A = chains_alloc<double>(ch, nB * nA);
    ch->L = chains_alloc<double>(ch, nB * nA);
    ch->ode = chains_alloc<double>(ch, nB * nO);
    ch->betaN = chains_alloc<double>(ch, nB * nrow);
    ch->latent = chains_alloc<double>(ch, nB * nO);
    ch->coalLog = chains_alloc<double>(ch, nB);
    ch->logOrigin = chains_alloc<double>(ch, nB);
    ch->summary = summary;  // slice of the batch's contiguous summary block (owned by lna_chains)
    ch->counters = chains_alloc<int>(ch, nB * 4);
    ch->tmp_i = chains_alloc<int>(ch, nB * 2);
    ch->tmp_d = chains_alloc<double>(ch, nB);
    ch->n_norm = (npar - 1) + kp;
    ch->n_unif = LNA_ITER_UNIF;
    const size_t stage_unif = std::max<size_t>(LNA_ITER_UNIF, LNA_ITER2_UNIF);
    ch->stage_n = chains_alloc<double>(ch, nB * ch->n_norm);
    ch->stage_u = chains_alloc<double>(ch, nB * stage_unif);
    ch->exec = chains_alloc<unsigned long long>(ch, 4);  // evaluations, warp-evaluations, warp-ticks (intervals), spare
    if (ch->exec) cudaMemsetAsync(ch->exec, 0, 4 * sizeof(unsigned long long), pb->stream);
    ch->jT = chains_alloc<double>(ch, nB * 16);
    ch->jnew = chains_alloc<double>(ch, nB * 4);
    if (!ch->par || !ch->par_alt || !ch->origin || !ch->A || !ch->L || !ch->ode || !ch->betaN || !ch->latent || !ch->coalLog ||
        !ch->logOrigin || !ch->counters || !ch->tmp_i || !ch->tmp_d || !ch->stage_n || !ch->stage_u || !ch->jT || !ch->jnew) {
        fail("lna_chains_create: device allocation failed");
        part_destroy(ch);
        return nullptr;
    }
    cudaMemsetAsync(ch->counters, 0, sizeof(int) * nB * 4, pb->stream);
    return ch;
}

static void part_destroy(ChainPart *ch) {
    if (!ch) return;
    cudaSetDevice(ch->pb->device);
    cudaStreamSynchronize(ch->pb->stream);
    for (void *p : ch->owned) cudaFree(p);
    ch->cand_scratch.release();
    for (auto *g : ch->graphs) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
    }
    delete ch;
}

static View soa_view(const double *d, long long B) { return View{d, 1, B}; }
static OutView soa_out(double *d, long long B) { return OutView{d, 1, B}; }

static int part_init(ChainPart *ch, const double *par, const double *OriginTraj) {
    if (!ch || !par || !OriginTraj) return fail("lna_chains_init: null argument");
    lna_problem *pb = ch->pb;
    const DevProb &P = pb->dp;
    const long long B = ch->B;
    const size_t nB = (size_t)B, npar = 2 + P.npt, kp = (size_t)P.k * P.p;
    Stage s(pb);
    const double *dpar = s.in(par, nB * npar), *dor = s.in(OriginTraj, nB * kp);
    if (!s.ok) return fail("lna_chains_init: device staging failed");
    k_to_soa<<<nblk((long long)(nB * npar), 256), 256, 0, pb->stream>>>(B, (int)npar, dpar, ch->par);
    if (launch_check(pb, "k_to_soa")) return -1;
    k_to_soa<<<nblk((long long)(nB * kp), 256), 256, 0, pb->stream>>>(B, (int)kp, dor, ch->origin);
    if (launch_check(pb, "k_to_soa")) return -1;
    // New_Param_List + TransformTraj + volz_loglik_nh2 straight into the state (general_change_point.R:507-536)
    FullOut out{soa_out(ch->A, B), soa_out(ch->L, B), soa_out(ch->ode, B), soa_out(ch->betaN, B), soa_out(ch->latent, B)};
    const View vp{ch->par + 2 * B, 1, B}, vi{ch->par, 1, B}, vo{ch->origin, 1, B};
    const int block = pick_block(pb, B);
    k_full_loglik<0, true><<<nblk(B, block), block, full_eval_smem_bytes(pb->tab_bytes, block), pb->stream>>>(P, pb->seg, B, vp, vi, vo, out,
                                                                                 ch->coalLog, ch->counters + 3 * B, 1);
    if (launch_check(pb, "k_full_loglik")) return -1;
    // MCMC_obj$logOrigin = -sum(OriginTraj^2)/2 over all rows (general_change_point.R:530): what `l[i]` records until the
    // first trajectory slice update replaces it
    k_log_origin<<<nblk(B * 32, 128), 128, 0, pb->stream>>>(B, (int)kp, vo, ch->logOrigin);
    if (launch_check(pb, "k_log_origin")) return -1;
    CK(cudaMemsetAsync(ch->counters, 0, sizeof(int) * nB * 3, pb->stream));
    return s.finish("lna_chains_init");
}

// candidate-evaluation arguments common to every stage of the resident chains
static void chains_base_args(ChainPart *ch, View normals, View uniforms, CandArgs &a) {
    const long long B = ch->B;
    a = CandArgs{};
    a.par = soa_view(ch->par, B);
    a.origin = soa_view(ch->origin, B);
    a.normals = normals;
    a.uniforms = uniforms;
    a.coal_log = ch->coalLog;
    a.cand = cand_out(ch->cb);
    a.win_slot = ch->cb.win_slot;
    a.cand_ll = ch->cb.cand_ll;
    a.theta = ch->cb.theta;
    a.n_evals = ch->cb.n_evals;
    a.status = ch->cb.status;
    a.accept = ch->cb.accept;
    a.exec_count = ch->exec;
}

static int chains_ensure_scratch(ChainPart *ch, int W) {
    lna_problem *pb = ch->pb;
    const long long B = ch->B;
    const int Wmax = slice_plan_slots(pb, ch->planB, W);
    if (ch->cb.NC < B * Wmax || ch->cb.B != B) {
        if (!cand_alloc(pb, ch->cand_scratch, 0, B, B * Wmax, ch->cb))
            return fail("lna_chains_step: candidate scratch allocation failed");
    }
    return 0;
}

// evaluate the proposals of one stage, copy the winners into the state, bump the counters
template <int kProp>
static int chains_stage(ChainPart *ch, const CandArgs &a, int W) {
    lna_problem *pb = ch->pb;
    const DevProb &P = pb->dp;
    const long long B = ch->B;
    if (launch_cand<kProp>(pb, B, W, a, ch->cb, ch->planB)) return -1;
    if (launch_commit_fields(pb, B, ch->cb, soa_out(ch->A, B), soa_out(ch->L, B), soa_out(ch->ode, B),
                             soa_out(ch->betaN, B), soa_out(ch->latent, B)))
        return -1;
    const int npar = 2 + P.npt;
    k_commit_par_wide<kProp><<<nblk(B * npar, 256), 256, 0, pb->stream>>>(P, B, a, soa_out(ch->par_alt, B), ch->coalLog);
    if (launch_check(pb, "k_commit_par_wide")) return -1;
    std::swap(ch->par, ch->par_alt);  // update_Par_ESlice_combine does not touch logOrigin (LNA_chpt_slice.R:456-472)
    const bool mh = !(kProp == PROP_ESS_PAR || kProp == PROP_ESS_CHPT);
    k_bump<<<nblk(B, 256), 256, 0, pb->stream>>>(B, ch->cb.n_evals, ch->cb.accept, ch->cb.status, ch->cb.win_slot,
                                                 ch->counters, 0, mh ? 1 : 0);
    return launch_check(pb, "k_bump");
}

// updateTraj_general_NC -> ESlice_general_NC, in place on the state
static int chains_traj(ChainPart *ch, View normals, View uniforms) {
    lna_problem *pb = ch->pb;
    const long long B = ch->B;
    TrajArgs a{};
    a.f_cur = soa_view(ch->origin, B);
    a.normals = normals;
    a.uniforms = uniforms;
    a.A = soa_view(ch->A, B);
    a.L = soa_view(ch->L, B);
    a.ode = soa_view(ch->ode, B);
    a.betaN = soa_view(ch->betaN, B);
    a.coal_log = ch->coalLog;
    a.latent = soa_out(ch->latent, B);
    a.origin_new = soa_out(ch->origin, B);
    a.CoalLog = ch->tmp_d;
    a.logOrigin = ch->logOrigin;
    a.n_evals = ch->tmp_i;
    a.status = ch->tmp_i + B;
    if (launch_traj(pb, B, a)) return -1;
    CK(cudaMemcpyAsync(ch->coalLog, ch->tmp_d, sizeof(double) * (size_t)B, cudaMemcpyDeviceToDevice, pb->stream));
    k_bump<<<nblk(B, 256), 256, 0, pb->stream>>>(B, ch->tmp_i, ch->tmp_i, ch->tmp_i + B, nullptr, ch->counters, 1, 0);
    return launch_check(pb, "k_bump");
}

static int chains_finish(ChainPart *ch) {
    lna_problem *pb = ch->pb;
    // every stage flips the parameter double buffer; an iteration with an odd number of stages would leave the state in the
    // other buffer, and a replayed CUDA graph of the iteration (captured pointers) would then read the wrong one
    if (ch->par != ch->par_home) {
        const size_t bytes = sizeof(double) * (size_t)ch->B * (size_t)(2 + pb->dp.npt);
        CK(cudaMemcpyAsync(ch->par_home, ch->par, bytes, cudaMemcpyDeviceToDevice, pb->stream));
        std::swap(ch->par, ch->par_alt);
    }
    k_summary<<<nblk(ch->B, 256), 256, 0, pb->stream>>>(ch->B, ch->coalLog, ch->logOrigin, ch->counters, ch->summary);
    return launch_check(pb, "k_summary");
}

// Both Metropolis entries of the vignettes speculatively in one latency period (3 evaluations per chain) or one after the
// other (2 evaluations, two latency periods): LNA_MH_PAIR=0/1 forces the choice (experiments).
static bool mh_pair_policy(long long planB, int sm_count) {
    if (const char *e = getenv("LNA_MH_PAIR")) return atoi(e) != 0;
    return planB * 4 <= (long long)sm_count * 512;
}

// General_MCMC_with_ESlice iteration body (R/LNA_chpt_slice.R:775-847)
static int part_step_dev(ChainPart *ch, const lna_mcmc_opts *o, const double *normals,
                                   const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step: null argument");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const DevProb &P = pb->dp;
    const int W = o->spec_width;  // 0 = auto (two-phase when the machine-filling width is 16)
    if (chains_ensure_scratch(ch, W)) return -1;
    const bool general = o->n_mh > 0;
    const long long sn = (long long)ch->n_norm, su = general ? (long long)LNA_ITER_UNIF_GENERAL : (long long)ch->n_unif;
    const int ess_u = general ? 8 : 3;  // first uniform of the parameter slice update
    const int npar = 2 + P.npt;
    CandArgs a;
    const char *tree_env = getenv("LNA_MH_TREE");  // 0 forces the sequential entries (tests, experiments)
    int m_max = 0;  // entries per speculative tree: 2^m lanes per chain must leave the launch latency-bound
    while (m_max < 4 && ch->planB * (2LL << m_max) <= (long long)pb->sm_count * 128) ++m_max;
    const bool tree = general && !(tree_env && atoi(tree_env) == 0) && o->n_mh >= 2 && o->n_mh <= 4 && m_max >= 2;
    if (tree) {
        // few chains: the entry list speculatively, m_max entries per latency period (PROP_MH_TREE, lna_eslice.cuh)
        for (int e = 0; e < o->n_mh; ++e) {
            if (o->mh_kind[e] == 0 && (o->mh_index[e] < 0 || o->mh_index[e] > 3))
                return fail("lna_chains_step: mh_index must be 0 (S0), 1 (I0), 2 (R0) or 3 (gamma)");
            if (o->mh_kind[e] < 0 || o->mh_kind[e] > 2) return fail("lna_chains_step: mh_kind must be 0, 1 or 2");
        }
        for (int e0 = 0; e0 < o->n_mh; e0 += m_max) {
            const int m = std::min(m_max, o->n_mh - e0), tW = 1 << std::max(m, 2);
            if (ch->cb.NC < ch->B * tW && !cand_alloc(pb, ch->cand_scratch, 0, ch->B, ch->B * tW, ch->cb))
                return fail("lna_chains_step: candidate scratch allocation failed");
            chains_base_args(ch, View{normals, sn, 1}, View{uniforms, su, 1}, a);
            a.tm = m;
            for (int e = 0; e < m; ++e) {
                const int g = e0 + e;
                a.tkind[e] = o->mh_kind[g] == 0 ? 0 : (o->mh_kind[g] == 1 ? 2 : 3);
                a.tidx[e] = o->mh_index[g];
                a.tlog[e] = o->mh_is_log[g];
                a.tpk[e] = 0;
                a.tpr[e][0] = o->mh_prior[g][0];
                a.tpr[e][1] = o->mh_prior[g][1];
                a.tprop[e] = o->mh_prop[g];
                a.tu[e] = 2 * g;
            }
            a.jnew = ch->jnew;
            if (chains_stage<PROP_MH_TREE>(ch, a, tW)) return -1;
        }
    } else if (general) {
        // Update_Param_each_Norm with an arbitrary entry list: one Update_Param per entry, in order
        if (o->n_mh > 4) return fail("lna_chains_step: n_mh must be 0..4");
        for (int e = 0; e < o->n_mh; ++e) {
            chains_base_args(ch, View{normals, sn, 1}, View{uniforms + 2 * e, su, 1}, a);
            a.mh_pr[0] = o->mh_prior[e][0];
            a.mh_pr[1] = o->mh_prior[e][1];
            a.mh_prop = o->mh_prop[e];
            if (o->mh_kind[e] == 0) {
                if (o->mh_index[e] < 0 || o->mh_index[e] > 3)
                    return fail("lna_chains_step: mh_index must be 0 (S0), 1 (I0), 2 (R0) or 3 (gamma)");
                a.mh_index = o->mh_index[e];
                a.mh_is_log = o->mh_is_log[e];
                a.mh_prior_kind = 0;
                if (chains_stage<PROP_MH_SCALAR>(ch, a, 1)) return -1;
            } else if (o->mh_kind[e] == 1) {
                chains_base_args(ch, View{normals, sn, 1}, View{uniforms + 2 * e + 1, su, 1}, a);  // (accept draw only)
                if (chains_stage<PROP_HYPER_NOOP>(ch, a, 1)) return -1;
            } else if (o->mh_kind[e] == 2) {
                a.mh_prior_kind = 2;
                if (chains_stage<PROP_MH_HYPER>(ch, a, 1)) return -1;
            } else {
                return fail("lna_chains_step: mh_kind must be 0, 1 or 2");
            }
        }
    }
    // 1. Update_Param_each_Norm: gamma random walk (entry c) and the no-op hyper entry (d)  [A.17]
    chains_base_args(ch, View{normals, sn, 1}, View{uniforms, su, 1}, a);
    a.gamma_prior[0] = o->gamma_prior[0];
    a.gamma_prior[1] = o->gamma_prior[1];
    a.gamma_prop = o->gamma_prop;
    if (general) {
        // (done above)
    } else if (o->hyper_noop_mh && mh_pair_policy(ch->planB, pb->sm_count)) {
        // few chains: both MH entries speculatively in one latency period (3 evaluations per chain)
        if (chains_stage<PROP_MH_PAIR>(ch, a, 4)) return -1;
    } else {
        // many chains: the machine is full anyway, so evaluate only what is consumed (2 per chain)
        if (chains_stage<PROP_GAMMA_RW>(ch, a, 1)) return -1;
        if (o->hyper_noop_mh) {
            chains_base_args(ch, View{normals, sn, 1}, View{uniforms + 2, su, 1}, a);
            if (chains_stage<PROP_HYPER_NOOP>(ch, a, 1)) return -1;
        }
    }
    // 2. update_Par_ESlice_combine -> ESlice_par_General
    chains_base_args(ch, View{normals, sn, 1}, View{uniforms + ess_u, su, 1}, a);
    for (int i = 0; i < 8; ++i) a.pr[i] = o->prior8[i];
    for (int i = 0; i < 7; ++i) a.ess[i] = o->ESS_vec[i];
    if (chains_stage<PROP_ESS_PAR>(ch, a, W)) return -1;
    if (o->update_traj && o->use_ess2) {
        // 3'. ESS_vec2: a second parameter slice update instead of the trajectory update (:826-833)
        chains_base_args(ch, View{normals + (npar - 1), sn, 1}, View{uniforms + ess_u + LNA_ESLICE_MAX_UNIF, su, 1}, a);
        for (int i = 0; i < 8; ++i) a.pr[i] = o->prior8[i];
        for (int i = 0; i < 7; ++i) a.ess[i] = o->ESS_vec2[i];
        if (chains_stage<PROP_ESS_PAR>(ch, a, W)) return -1;
        return chains_finish(ch);
    }
    // 3. updateTraj_general_NC -> ESlice_general_NC
    if (o->update_traj &&
        chains_traj(ch, View{normals + (npar - 1), sn, 1}, View{uniforms + ess_u + LNA_ESLICE_MAX_UNIF, su, 1}))
        return -1;
    return chains_finish(ch);
}

// General_MCMC2 iteration body (R/LNA_chpt_slice.R:576-691, stages i < burn1 and burn1 <= i < warmup2)
static int part_step_mcmc2_dev(ChainPart *ch, const lna_mcmc2_opts *o, const double *normals,
                                         const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step_mcmc2: null argument");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const DevProb &P = pb->dp;
    if (o->n_mh < 0 || o->n_mh > 4) return fail("lna_chains_step_mcmc2: n_mh must be 0..4");
    const int W = o->spec_width;
    if (chains_ensure_scratch(ch, W)) return -1;
    const long long sn = (long long)(P.nch + P.k * P.p) + (o->joint ? 4 : 0), su = (long long)LNA_ITER2_UNIF;
    CandArgs a;
    if (o->joint) {  // i >= warmup2: update_Param_joint(method = "admcmc") replaces the scalar entries and the hyper step
        if (o->joint_d < 1 || o->joint_d > 4) return fail("lna_chains_step_mcmc2: joint_d must be 1..4");
        if (!ch->jT || ch->jd != o->joint_d)
            return fail("lna_chains_step_mcmc2: joint step without a matching lna_chains_set_joint_transform (d = %d)", o->joint_d);
        chains_base_args(ch, View{normals, sn, 1}, View{uniforms, su, 1}, a);
        a.jd = o->joint_d;
        for (int i = 0; i < a.jd; ++i) {
            if (o->joint_index[i] < 1 || o->joint_index[i] > 4 || (o->joint_index[i] == 4 && i != a.jd - 1))
                return fail("lna_chains_step_mcmc2: joint_index must be 1 (I0), 2 (R0), 3 (gamma) or, last, 4 (hyper)");
            a.jidx[i] = o->joint_index[i];
            a.jlog[i] = o->joint_is_log[i];
            a.jkind[i] = o->joint_prior_kind[i];
            a.jpr[i][0] = o->joint_prior[i][0];
            a.jpr[i][1] = o->joint_prior[i][1];
        }
        a.jnorm_off = P.nch + P.k * P.p;
        a.jT = ch->jT;
        a.jnew = ch->jnew;
        if (chains_stage<PROP_MH_JOINT>(ch, a, 1)) return -1;
    }
    // The Metropolis entries (Update_Param_each + update_hyper) speculatively, as trees of m entries per latency period:
    // m = the largest tree whose 2^m lanes per chain still leave the launch latency-bound (4 up to 1184 chains, 3 up to
    // 2368, 2 up to 4736); with more chains the machine is full and the entries run one by one.
    const int tm = o->joint ? 0 : o->n_mh + (o->update_hyper ? 1 : 0);
    const char *tree_env = getenv("LNA_MH_TREE");  // 0 forces the sequential entries (tests, experiments)
    const bool tree_off = tree_env && atoi(tree_env) == 0;
    int m_max = 0;
    while (m_max < 4 && ch->planB * (2LL << m_max) <= (long long)pb->sm_count * 128) ++m_max;
    const bool tree = !tree_off && tm >= 2 && m_max >= 2;
    if (tree) {
        struct Ent {
            int kind, idx, lg, pk, tu;
            double pr0, pr1, prop;
        } ent[4];
        for (int e = 0; e < o->n_mh; ++e) {
            if (o->mh_index[e] < 1 || o->mh_index[e] > 3)
                return fail("lna_chains_step_mcmc2: mh_index must be 1 (I0), 2 (R0) or 3 (gamma)");
            ent[e] = {0, o->mh_index[e], o->mh_is_log[e], o->mh_prior_kind[e], 2 * e, o->mh_prior[e][0], o->mh_prior[e][1],
                      o->mh_prop[e]};
        }
        if (o->update_hyper) ent[o->n_mh] = {1, -1, 0, 2, 8, o->hyper_prior[0], o->hyper_prior[1], o->hyper_prop};
        for (int e0 = 0; e0 < tm; e0 += m_max) {
            const int m = std::min(m_max, tm - e0), tW = 1 << std::max(m, 2);
            if (ch->cb.NC < ch->B * tW && !cand_alloc(pb, ch->cand_scratch, 0, ch->B, ch->B * tW, ch->cb))
                return fail("lna_chains_step_mcmc2: candidate scratch allocation failed");
            chains_base_args(ch, View{normals, sn, 1}, View{uniforms, su, 1}, a);
            a.tm = m;
            for (int e = 0; e < m; ++e) {
                const Ent &t = ent[e0 + e];
                a.tkind[e] = t.kind;
                a.tidx[e] = t.idx;
                a.tlog[e] = t.lg;
                a.tpk[e] = t.pk;
                a.tpr[e][0] = t.pr0;
                a.tpr[e][1] = t.pr1;
                a.tprop[e] = t.prop;
                a.tu[e] = t.tu;
            }
            a.jnew = ch->jnew;
            if (chains_stage<PROP_MH_TREE>(ch, a, tW)) return -1;
        }
    }
    // Update_Param_each: one Metropolis entry per listed parameter, sequentially
    for (int e = 0; !tree && !o->joint && e < o->n_mh; ++e) {
        if (o->mh_index[e] < 1 || o->mh_index[e] > 3)
            return fail("lna_chains_step_mcmc2: mh_index must be 1 (I0), 2 (R0) or 3 (gamma)");
        chains_base_args(ch, View{normals, sn, 1}, View{uniforms + 2 * e, su, 1}, a);
        a.mh_index = o->mh_index[e];
        a.mh_is_log = o->mh_is_log[e];
        a.mh_prior_kind = o->mh_prior_kind[e];
        a.mh_pr[0] = o->mh_prior[e][0];
        a.mh_pr[1] = o->mh_prior[e][1];
        a.mh_prop = o->mh_prop[e];
        if (chains_stage<PROP_MH_SCALAR>(ch, a, 1)) return -1;
    }
    if (o->update_hyper && !o->joint && !tree) {  // update_hyper (general_change_point.R:281-313)
        chains_base_args(ch, View{normals, sn, 1}, View{uniforms + 8, su, 1}, a);
        a.mh_pr[0] = o->hyper_prior[0];
        a.mh_pr[1] = o->hyper_prior[1];
        a.mh_prop = o->hyper_prop;
        if (chains_stage<PROP_MH_HYPER>(ch, a, 1)) return -1;
    }
    if (o->update_chpt) {  // update_ChangePoint_ESlice -> ESlice_change_points
        chains_base_args(ch, View{normals, sn, 1}, View{uniforms + 10, su, 1}, a);
        if (chains_stage<PROP_ESS_CHPT>(ch, a, W)) return -1;
    }
    if (o->update_traj &&
        chains_traj(ch, View{normals + P.nch, sn, 1}, View{uniforms + 10 + LNA_ESLICE_MAX_UNIF, su, 1}))
        return -1;
    return chains_finish(ch);
}

static int part_step_mcmc2(ChainPart *ch, const lna_mcmc2_opts *o, const double *normals,
                                     const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step_mcmc2: null argument");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const DevProb &P = pb->dp;
    const size_t nB = (size_t)ch->B, nn = (size_t)(P.nch + P.k * P.p) + (o->joint ? 4 : 0);
    if (nn > ch->n_norm || (size_t)LNA_ITER2_UNIF > ch->n_unif + 8) return fail("lna_chains_step_mcmc2: staging too small");
    CK(cudaMemcpyAsync(ch->stage_n, normals, sizeof(double) * nB * nn, cudaMemcpyHostToDevice, pb->stream));
    CK(cudaMemcpyAsync(ch->stage_u, uniforms, sizeof(double) * nB * LNA_ITER2_UNIF, cudaMemcpyHostToDevice, pb->stream));
    return part_step_mcmc2_dev(ch, o, ch->stage_n, ch->stage_u);
}

static int part_step(ChainPart *ch, const lna_mcmc_opts *o, const double *normals,
                               const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step: null argument");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const size_t nB = (size_t)ch->B;
    CK(cudaMemcpyAsync(ch->stage_n, normals, sizeof(double) * nB * ch->n_norm, cudaMemcpyHostToDevice, pb->stream));
    const size_t nu = o->n_mh > 0 ? (size_t)LNA_ITER_UNIF_GENERAL : ch->n_unif;
    CK(cudaMemcpyAsync(ch->stage_u, uniforms, sizeof(double) * nB * nu, cudaMemcpyHostToDevice, pb->stream));
    return part_step_dev(ch, o, ch->stage_n, ch->stage_u);
}

// ------------------------------------------------------------------------------------------------
// on-device replayable streams
// ------------------------------------------------------------------------------------------------
static int draw_streams_dev(lna_problem *pb, long long B, long long chain_offset, int n_norm, int n_unif,
                            uint64_t seed, uint64_t iteration, double *dn, double *du) {
    const long long per = (n_norm + 1) / 2 + (n_unif + 1) / 2, n = B * per;
    if (per >= 65536 || (iteration >> 48) != 0)
        return fail("lna_draw_streams: the Philox counter holds 2^16 output pairs per chain-iteration and iterations < 2^48");
    k_draw_streams<<<nblk(n, 256), 256, 0, pb->stream>>>(B, chain_offset, n_norm, n_unif, seed, iteration, dn, du);
    return launch_check(pb, "k_draw_streams");
}

extern "C" int lna_draw_streams(lna_problem *pb, int64_t B, int64_t chain_offset, int n_norm, int n_unif,
                                uint64_t seed, uint64_t iteration, double *normals, double *uniforms) {
    if (!pb) return fail("null problem");
    if (B <= 0) return 0;
    if (n_norm < 0 || n_unif < 0 || !normals || !uniforms) return fail("lna_draw_streams: bad arguments");
    const size_t nB = (size_t)B;
    Stage s(pb);
    double *dn = s.tmp<double>(nB * n_norm), *du = s.tmp<double>(nB * n_unif);
    if (!s.ok) return fail("lna_draw_streams: device staging failed");
    if (draw_streams_dev(pb, B, chain_offset, n_norm, n_unif, seed, iteration, dn, du)) return -1;
    s.back(normals, dn, nB * n_norm);
    s.back(uniforms, du, nB * n_unif);
    return s.finish("lna_draw_streams");
}

static int part_step_rng(ChainPart *ch, const lna_mcmc_opts *o, uint64_t seed, uint64_t iteration,
                                   int64_t chain_offset) {
    if (!ch || !o) return fail("lna_chains_step_rng: null argument");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const int nu = o->n_mh > 0 ? (int)LNA_ITER_UNIF_GENERAL : (int)ch->n_unif;
    if (draw_streams_dev(pb, ch->B, chain_offset, (int)ch->n_norm, nu, seed, iteration, ch->stage_n, ch->stage_u))
        return -1;
    return part_step_dev(ch, o, ch->stage_n, ch->stage_u);
}

static int part_step_mcmc2_rng(ChainPart *ch, const lna_mcmc2_opts *o, uint64_t seed, uint64_t iteration,
                               int64_t chain_offset) {
    if (!ch || !o) return fail("lna_chains_step_mcmc2_rng: null argument");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const DevProb &P = pb->dp;
    const int nn = P.nch + P.k * P.p + (o->joint ? 4 : 0);
    if ((size_t)nn > ch->n_norm) return fail("lna_chains_step_mcmc2_rng: staging too small");
    if (draw_streams_dev(pb, ch->B, chain_offset, nn, (int)LNA_ITER2_UNIF, seed, iteration, ch->stage_n, ch->stage_u)) return -1;
    return part_step_mcmc2_dev(ch, o, ch->stage_n, ch->stage_u);
}

// One iteration on device-drawn streams through a CUDA graph.  The first call with a given option set runs eagerly (it
// sizes the scratch buffers and sets the kernels' shared-memory attributes), the second one is captured on the group's stream,
// later ones replay the instantiated graph with the k_draw_streams node's `iteration` argument updated.  Same kernels, same
// arguments, same order as the eager path: same bits.  LNA_GRAPH=0 keeps the eager launches.
template <class Opts, class Eager>
static int part_step_graph(ChainPart *ch, int kind, const Opts *o, uint64_t seed, uint64_t iteration, int64_t chain_offset,
                           Eager eager) {
    static const bool enabled = [] {
        const char *e = getenv("LNA_GRAPH");
        return !(e && atoi(e) == 0);
    }();
    if (!enabled) return eager();
    lna_problem *pb = ch->pb;
    std::vector<unsigned char> key(sizeof(int) + sizeof(Opts) + sizeof(uint64_t) + sizeof(int64_t));
    std::memcpy(key.data(), &kind, sizeof(int));
    std::memcpy(key.data() + sizeof(int), o, sizeof(Opts));
    std::memcpy(key.data() + sizeof(int) + sizeof(Opts), &seed, sizeof(uint64_t));
    std::memcpy(key.data() + sizeof(int) + sizeof(Opts) + sizeof(uint64_t), &chain_offset, sizeof(int64_t));
    auto replay = [&](ChainPart::StepGraph *g) -> int {
        if (g->broken) return eager();
        g->aiter = iteration;
        cudaError_t e = cudaGraphExecKernelNodeSetParams(g->exec, g->draw_node, &g->draw_params);
        if (e == cudaSuccess) e = cudaGraphLaunch(g->exec, pb->stream);
        if (e != cudaSuccess) {  // (never observed after the first successful replay; the eager path is always valid)
            if (getenv("LNA_GRAPH_DEBUG")) fprintf(stderr, "lna: graph replay failed (%s), eager launches from here on\n", cudaGetErrorString(e));
            cudaGetLastError();
            g->broken = true;
            return eager();
        }
        pb->launches += g->n_kernels;
        if (pb->parent) pb->parent->launches += g->n_kernels;
        return 0;
    };
    for (auto *g : ch->graphs)
        if (g->key == key) return replay(g);
    if (ch->last_key != key || ch->graphs.size() >= 8) {
        ch->last_key = key;
        return eager();
    }
    // second iteration with these options: capture
    cudaSetDevice(pb->device);
    const int64_t l0 = pb->launches, lp0 = pb->parent ? pb->parent->launches : 0;
    if (cudaStreamBeginCapture(pb->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return eager();
    }
    const int rc = eager();
    cudaGraph_t graph = nullptr;
    const cudaError_t ec = cudaStreamEndCapture(pb->stream, &graph);
    const int nk = (int)(pb->launches - l0);
    pb->launches = l0;  // (nothing ran yet)
    if (pb->parent) pb->parent->launches = lp0;
    if (rc != 0 || ec != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc != 0) return rc;
        ch->last_key.clear();
        return eager();
    }
    auto *g = new ChainPart::StepGraph();
    g->key = key;
    g->n_kernels = nk;
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    std::vector<cudaGraphNode_t> nodes(nn);
    cudaGraphGetNodes(graph, nodes.data(), &nn);
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp{};
        if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) continue;
        if (kp.func == (void *)k_draw_streams) {
            g->draw_node = nd;
            g->draw_params = kp;
            void **a = kp.kernelParams;
            g->aB = *(long long *)a[0];
            g->aoff = *(long long *)a[1];
            g->ann = *(int *)a[2];
            g->anu = *(int *)a[3];
            g->aseed = *(unsigned long long *)a[4];
            g->aiter = *(unsigned long long *)a[5];
            g->adn = *(double **)a[6];
            g->adu = *(double **)a[7];
        }
    }
    void *argv[8] = {&g->aB, &g->aoff, &g->ann, &g->anu, &g->aseed, &g->aiter, &g->adn, &g->adu};
    std::memcpy(g->argv, argv, sizeof argv);
    g->draw_params.kernelParams = g->argv;
    g->draw_params.extra = nullptr;
    cudaError_t ei = g->draw_node ? cudaGraphInstantiate(&g->exec, graph, 0) : cudaErrorUnknown;
    if (ei != cudaSuccess) {
        cudaGraphDestroy(graph);
        cudaGetLastError();
        delete g;
        ch->last_key.clear();
        return eager();
    }
    g->graph = graph;
    ch->graphs.push_back(g);
    return replay(g);
}

static int part_get(ChainPart *ch, double *par, double *OriginTraj, double *LatentTraj, double *coalLog,
                              double *logOrigin, int32_t *counters) {
    if (!ch) return fail("lna_chains_get: null chains");
    lna_problem *pb = ch->pb;
    const DevProb &P = pb->dp;
    const long long B = ch->B;
    const size_t nB = (size_t)B, npar = 2 + P.npt, kp = (size_t)P.k * P.p, nO = (size_t)P.nrow * (P.p + 1);
    Stage s(pb);
    auto fetch = [&](double *h, const double *soa, size_t n) -> int {
        if (!h) return 0;
        double *d = s.tmp<double>(nB * n);
        if (!d) return -1;
        k_from_soa<<<nblk((long long)(nB * n), 256), 256, 0, pb->stream>>>(B, (int)n, soa, d);
        if (launch_check(pb, "k_from_soa")) return -1;
        s.back(h, d, nB * n);
        return 0;
    };
    if (fetch(par, ch->par, npar) || fetch(OriginTraj, ch->origin, kp) || fetch(LatentTraj, ch->latent, nO)) return -1;
    if (coalLog) s.back(coalLog, ch->coalLog, nB);
    if (logOrigin) s.back(logOrigin, ch->logOrigin, nB);
    if (counters) {
        // [4][B] -> B x 4
        std::vector<int> tmp(nB * 4);
        if (cudaStreamSynchronize(pb->stream) != cudaSuccess ||
            cudaMemcpy(tmp.data(), ch->counters, sizeof(int) * nB * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail("lna_chains_get: counter copy failed");
        for (size_t c = 0; c < nB; ++c)
            for (int j = 0; j < 4; ++j) counters[c * 4 + j] = tmp[(size_t)j * nB + c];
    }
    return s.finish("lna_chains_get");
}



// ------------------------------------------------------------------------------------------------
// public handle: a resident chain batch = G stream-groups of chains.  With many chains the late slice
// phases and the Metropolis steps of one group sit on the single-evaluation latency floor; running two
// groups on two streams lets one group's latency-bound launches overlap the other's throughput-bound ones
// (measured +29 % at 65 536 chains, +35 % at 16 384, +5 % at 4096; tools/eslice_groups.py).
// ------------------------------------------------------------------------------------------------
struct lna_chains {
    lna_problem *pb = nullptr;
    long long B = 0;
    int G = 1;
    std::vector<lna_problem *> gpb;  // per-group problems (clones; gpb[0] == pb when G == 1)
    std::vector<ChainPart *> part;
    std::vector<long long> off;      // first chain of each group, off[G] = B
    double *summary = nullptr;       // B x 4, contiguous over the groups
    cudaEvent_t ev_in = nullptr;
    std::vector<cudaEvent_t> ev_out;
    size_t n_norm = 0, n_unif = 0, n_norm2 = 0;
    // per-iteration history kept on the device (params[i,], l[i], l1[i] of the R drivers): one buffer per group,
    // [niter][npar + 2][Bg] in the groups' own slot-major layout, so recording is three device-to-device copies
    std::vector<double *> hist;
    long long hist_niter = 0;
    // free-running stream groups: the groups' streams are NOT joined after every step, so one group's latency-bound stages
    // overlap another group's throughput-bound ones ACROSS iterations (chains of different groups never interact); the
    // caller observes results through lna_chains_get / lna_chains_join
    int free_running = 0;
};

// history slot of one iteration of one group: par (npar x Bg), then logOrigin (Bg), then coalLog (Bg)
__global__ void k_hist_gather(long long Bg, int npar, long long niter, const double *hist, double *par_out, double *l_out,
                              double *l1_out, long long Btot, long long c0) {
    // one thread per (chain, iteration, entry): par_out is Btot x niter x npar (chain-major, like the R arrays per chain)
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_it = (long long)(npar + 2);
    if (idx >= Bg * niter * per_it) return;
    const int j = (int)(idx % per_it);
    const long long i = (idx / per_it) % niter, c = idx / (per_it * niter);
    const double v = hist[(i * per_it + j) * Bg + c];
    if (j < npar) {
        if (par_out) par_out[((c0 + c) * niter + i) * npar + j] = v;
    } else if (j == npar) {
        if (l_out) l_out[(c0 + c) * niter + i] = v;
    } else {
        if (l1_out) l1_out[(c0 + c) * niter + i] = v;
    }
    (void)Btot;
}

// order the groups' streams after everything already enqueued on the user's stream ...
static int chains_fork(lna_chains *ch) {
    if (ch->G == 1) return 0;
    CK(cudaEventRecord(ch->ev_in, ch->pb->stream));
    for (int g = 0; g < ch->G; ++g) CK(cudaStreamWaitEvent(ch->gpb[g]->stream, ch->ev_in, 0));
    return 0;
}
// ... and the user's stream after the groups
static int chains_join_now(lna_chains *ch) {
    if (ch->G == 1) return 0;
    for (int g = 0; g < ch->G; ++g) {
        CK(cudaEventRecord(ch->ev_out[g], ch->gpb[g]->stream));
        CK(cudaStreamWaitEvent(ch->pb->stream, ch->ev_out[g], 0));
    }
    return 0;
}
static int chains_join(lna_chains *ch) { return ch->free_running ? 0 : chains_join_now(ch); }

// busy-wait of `ns` nanoseconds on one thread: offsets the stream groups against each other
__global__ void k_delay(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}

extern "C" int lna_chains_set_free_running(lna_chains *ch, int on) {
    if (!ch) return fail("lna_chains_set_free_running: null chains");
    cudaSetDevice(ch->pb->device);
    if (ch->free_running && !on && chains_join_now(ch)) return -1;
    if (on && !ch->free_running && ch->G > 1) {
        // Identical groups started together stay in phase (all in their Metropolis stage, then all in the first slice phase,
        // ...), which is the lock-step this mode is meant to break: group g starts g * stagger later, so that at any moment
        // the groups are spread over the stages of an iteration.  LNA_STAGGER_US overrides the offset (0 = none).
        double us = 700.0 / ch->G;
        if (const char *e = getenv("LNA_STAGGER_US")) us = atof(e);
        if (chains_fork(ch)) return -1;
        for (int g = 1; g < ch->G && us > 0.0; ++g) {
            k_delay<<<1, 1, 0, ch->gpb[g]->stream>>>((unsigned long long)(us * 1e3 * g));
            if (launch_check(ch->gpb[g], "k_delay")) return -1;
        }
    }
    ch->free_running = on ? 1 : 0;
    return 0;
}
extern "C" int lna_chains_join(lna_chains *ch) {
    if (!ch) return fail("lna_chains_join: null chains");
    cudaSetDevice(ch->pb->device);
    return chains_join_now(ch);
}
extern "C" int lna_chains_sync(lna_chains *ch) {
    if (!ch) return fail("lna_chains_sync: null chains");
    cudaSetDevice(ch->pb->device);
    for (int g = 0; g < ch->G; ++g) CK(cudaStreamSynchronize(ch->gpb[g]->stream));
    CK(cudaStreamSynchronize(ch->pb->stream));
    return 0;
}
extern "C" int lna_chains_groups(const lna_chains *ch) { return ch ? ch->G : 0; }

extern "C" lna_chains *lna_chains_create(lna_problem *pb, int64_t B) {
    if (!pb || B <= 0) {
        fail("lna_chains_create: bad arguments");
        return nullptr;
    }
    if (check_sir_sampler(pb, "lna_chains_create")) return nullptr;
    cudaSetDevice(pb->device);
    lna_chains *ch = new lna_chains();
    ch->pb = pb;
    ch->B = B;
    // two stream groups from 8192 chains on (measured with the iterations replayed from CUDA graphs: one group is faster up
    // to 4096 chains, 1.08 vs 1.16 ms per iteration there; two win from 8192, profiles/r2_eslice_timeline.md)
    ch->G = B >= 8192 ? 2 : 1;
    // the barrier-free slice kernel keeps the machine busy from one stream (persistent groups refill from the batch)
    if (pb->dp.async_ok && pb->dp.has_init && (long long)B * 8 / 32 * 10 <= (long long)pb->sm_count * 4 * kAsyncMaxWarps10) ch->G = 1;
    if (const char *e = getenv("LNA_CHAIN_GROUPS")) ch->G = std::max(1, std::min(8, atoi(e)));
    if (ch->G > B) ch->G = 1;
    const DevProb &P = pb->dp;
    ch->n_norm = (size_t)(2 + P.npt - 1) + (size_t)P.k * P.p;
    ch->n_unif = LNA_ITER_UNIF;
    ch->n_norm2 = (size_t)P.nch + (size_t)P.k * P.p;
    if (cudaMalloc(&ch->summary, sizeof(double) * 4 * (size_t)B) != cudaSuccess) {
        fail("lna_chains_create: device allocation failed");
        delete ch;
        return nullptr;
    }
    cudaEventCreateWithFlags(&ch->ev_in, cudaEventDisableTiming);
    ch->off.assign(ch->G + 1, 0);
    for (int g = 0; g < ch->G; ++g) {
        int64_t b0, b1;
        lna_shard(B, g, ch->G, &b0, &b1);
        ch->off[g] = b0;
        ch->off[g + 1] = b1;
        lna_problem *gp = ch->G == 1 ? pb : problem_clone(pb);
        ChainPart *pt = gp ? part_create(gp, b1 - b0, B, ch->summary + 4 * b0) : nullptr;
        if (!pt) {
            if (gp && gp != pb) lna_problem_destroy(gp);
            lna_chains_destroy(ch);
            return nullptr;
        }
        ch->gpb.push_back(gp);
        ch->part.push_back(pt);
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ch->ev_out.push_back(e);
    }
    return ch;
}

extern "C" void lna_chains_destroy(lna_chains *ch) {
    if (!ch) return;
    cudaSetDevice(ch->pb->device);
    for (size_t g = 0; g < ch->part.size(); ++g) {
        part_destroy(ch->part[g]);
        if (ch->gpb[g] != ch->pb) lna_problem_destroy(ch->gpb[g]);
    }
    for (cudaEvent_t e : ch->ev_out) cudaEventDestroy(e);
    if (ch->ev_in) cudaEventDestroy(ch->ev_in);
    if (ch->summary) cudaFree(ch->summary);
    for (double *h : ch->hist)
        if (h) cudaFree(h);
    delete ch;
}

// ---- on-device history of the drivers' per-iteration records -------------------------------------------------------------
extern "C" int lna_chains_history_begin(lna_chains *ch, int64_t niter) {
    if (!ch || niter <= 0) return fail("lna_chains_history_begin: bad arguments");
    cudaSetDevice(ch->pb->device);
    for (double *h : ch->hist)
        if (h) cudaFree(h);
    ch->hist.assign(ch->G, nullptr);
    ch->hist_niter = 0;
    const size_t per_it = (size_t)(2 + ch->pb->dp.npt) + 2;
    for (int g = 0; g < ch->G; ++g)
        if (cudaMalloc(&ch->hist[g], sizeof(double) * (size_t)niter * per_it * (size_t)ch->part[g]->B) != cudaSuccess) {
            cudaGetLastError();
            return fail("lna_chains_history_begin: device allocation failed (%lld iterations x %lld chains)", (long long)niter,
                        ch->B);
        }
    ch->hist_niter = niter;
    return 0;
}

extern "C" int lna_chains_history_record(lna_chains *ch, int64_t i) {
    if (!ch || ch->hist_niter == 0) return fail("lna_chains_history_record: no history (lna_chains_history_begin)");
    if (i < 0 || i >= ch->hist_niter) return fail("lna_chains_history_record: iteration %lld outside 0..%lld", (long long)i,
                                                  ch->hist_niter - 1);
    const size_t npar = (size_t)(2 + ch->pb->dp.npt), per_it = npar + 2;
    if (chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g) {
        ChainPart *pt = ch->part[g];
        const size_t Bg = (size_t)pt->B;
        double *dst = ch->hist[g] + (size_t)i * per_it * Bg;
        cudaStream_t st = pt->pb->stream;
        CK(cudaMemcpyAsync(dst, pt->par, sizeof(double) * npar * Bg, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst + npar * Bg, pt->logOrigin, sizeof(double) * Bg, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst + (npar + 1) * Bg, pt->coalLog, sizeof(double) * Bg, cudaMemcpyDeviceToDevice, st));
    }
    return chains_join(ch);
}

extern "C" int lna_chains_history_get(lna_chains *ch, double *par, double *l, double *l1) {
    if (!ch || ch->hist_niter == 0) return fail("lna_chains_history_get: no history (lna_chains_history_begin)");
    lna_problem *pb = ch->pb;
    cudaSetDevice(pb->device);
    const long long niter = ch->hist_niter;
    const int npar = 2 + pb->dp.npt;
    const size_t nB = (size_t)ch->B;
    double *dpar = nullptr, *dl = nullptr, *dl1 = nullptr;
    cudaError_t e = cudaSuccess;
    if (par) e = cudaMalloc(&dpar, sizeof(double) * nB * niter * npar);
    if (e == cudaSuccess && l) e = cudaMalloc(&dl, sizeof(double) * nB * niter);
    if (e == cudaSuccess && l1) e = cudaMalloc(&dl1, sizeof(double) * nB * niter);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    for (int g = 0; e == cudaSuccess && g < ch->G; ++g) {
        const long long Bg = ch->part[g]->B, n = Bg * niter * (npar + 2);
        k_hist_gather<<<nblk(n, 256), 256, 0, pb->stream>>>(Bg, npar, niter, ch->hist[g], dpar, dl, dl1, ch->B, ch->off[g]);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && par) e = cudaMemcpyAsync(par, dpar, sizeof(double) * nB * niter * npar, cudaMemcpyDeviceToHost, pb->stream);
    if (e == cudaSuccess && l) e = cudaMemcpyAsync(l, dl, sizeof(double) * nB * niter, cudaMemcpyDeviceToHost, pb->stream);
    if (e == cudaSuccess && l1) e = cudaMemcpyAsync(l1, dl1, sizeof(double) * nB * niter, cudaMemcpyDeviceToHost, pb->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(pb->stream);
    if (dpar) cudaFree(dpar);
    if (dl) cudaFree(dl);
    if (dl1) cudaFree(dl1);
    return e == cudaSuccess ? 0 : fail("lna_chains_history_get: %s", cudaGetErrorString(e));
}

extern "C" int lna_chains_init(lna_chains *ch, const double *par, const double *OriginTraj) {
    if (!ch || !par || !OriginTraj) return fail("lna_chains_init: null argument");
    const DevProb &P = ch->pb->dp;
    const size_t npar = 2 + P.npt, kp = (size_t)P.k * P.p;
    for (int g = 0; g < ch->G; ++g)
        if (part_init(ch->part[g], par + ch->off[g] * npar, OriginTraj + ch->off[g] * kp)) return -1;
    return 0;
}

extern "C" int lna_chains_set_joint_transform(lna_chains *ch, int d, const double *T) {
    if (!ch || !T) return fail("lna_chains_set_joint_transform: null argument");
    if (d < 1 || d > 4) return fail("lna_chains_set_joint_transform: d must be 1..4");
    std::vector<double> pad;
    for (int g = 0; g < ch->G; ++g) {
        ChainPart *pt = ch->part[g];
        cudaSetDevice(pt->pb->device);
        const size_t nB = (size_t)pt->B;
        pad.assign(nB * 16, 0.0);
        const double *src = T + (size_t)ch->off[g] * d * d;
        for (size_t c = 0; c < nB; ++c)
            for (int i = 0; i < d; ++i)
                for (int q = 0; q < d; ++q) pad[c * 16 + i * 4 + q] = src[c * d * d + i * d + q];
        CK(cudaMemcpyAsync(pt->jT, pad.data(), sizeof(double) * nB * 16, cudaMemcpyHostToDevice, pt->pb->stream));
        CK(cudaStreamSynchronize(pt->pb->stream));  // (pad is a host temporary)
        pt->jd = d;
    }
    return 0;
}

extern "C" int lna_chains_step_dev(lna_chains *ch, const lna_mcmc_opts *o, const double *normals,
                                   const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step: null argument");
    if (chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g)
        if (part_step_dev(ch->part[g], o, normals + ch->off[g] * ch->n_norm,
                          uniforms + ch->off[g] * (o->n_mh > 0 ? (size_t)LNA_ITER_UNIF_GENERAL : ch->n_unif)))
            return -1;
    return chains_join(ch);
}

extern "C" int lna_chains_step(lna_chains *ch, const lna_mcmc_opts *o, const double *normals, const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step: null argument");
    if (chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g)
        if (part_step(ch->part[g], o, normals + ch->off[g] * ch->n_norm,
                      uniforms + ch->off[g] * (o->n_mh > 0 ? (size_t)LNA_ITER_UNIF_GENERAL : ch->n_unif)))
            return -1;
    return chains_join(ch);
}

extern "C" int lna_chains_step_rng(lna_chains *ch, const lna_mcmc_opts *o, uint64_t seed, uint64_t iteration,
                                   int64_t chain_offset) {
    if (!ch || !o) return fail("lna_chains_step_rng: null argument");
    if (!ch->free_running && chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g) {
        ChainPart *pt = ch->part[g];
        const int64_t off = chain_offset + ch->off[g];
        if (part_step_graph(pt, 0, o, seed, iteration, off, [&] { return part_step_rng(pt, o, seed, iteration, off); })) return -1;
    }
    return chains_join(ch);
}

extern "C" int lna_chains_step_mcmc2_dev(lna_chains *ch, const lna_mcmc2_opts *o, const double *normals,
                                         const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step_mcmc2: null argument");
    if (chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g)
        if (part_step_mcmc2_dev(ch->part[g], o, normals + ch->off[g] * (ch->n_norm2 + (o->joint ? 4 : 0)),
                                uniforms + ch->off[g] * (size_t)LNA_ITER2_UNIF))
            return -1;
    return chains_join(ch);
}

extern "C" int lna_chains_step_mcmc2(lna_chains *ch, const lna_mcmc2_opts *o, const double *normals,
                                     const double *uniforms) {
    if (!ch || !o || !normals || !uniforms) return fail("lna_chains_step_mcmc2: null argument");
    if (chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g)
        if (part_step_mcmc2(ch->part[g], o, normals + ch->off[g] * (ch->n_norm2 + (o->joint ? 4 : 0)),
                            uniforms + ch->off[g] * (size_t)LNA_ITER2_UNIF))
            return -1;
    return chains_join(ch);
}

extern "C" int lna_chains_step_mcmc2_rng(lna_chains *ch, const lna_mcmc2_opts *o, uint64_t seed, uint64_t iteration,
                                         int64_t chain_offset) {
    if (!ch || !o) return fail("lna_chains_step_mcmc2_rng: null argument");
    if (!ch->free_running && chains_fork(ch)) return -1;
    for (int g = 0; g < ch->G; ++g) {
        ChainPart *pt = ch->part[g];
        const int64_t off = chain_offset + ch->off[g];
        if (part_step_graph(pt, 1, o, seed, iteration, off, [&] { return part_step_mcmc2_rng(pt, o, seed, iteration, off); }))
            return -1;
    }
    return chains_join(ch);
}

extern "C" int lna_chains_get(lna_chains *ch, double *par, double *OriginTraj, double *LatentTraj, double *coalLog,
                              double *logOrigin, int32_t *counters) {
    if (!ch) return fail("lna_chains_get: null chains");
    const DevProb &P = ch->pb->dp;
    const size_t npar = 2 + P.npt, kp = (size_t)P.k * P.p, nO = (size_t)P.nrow * (P.p + 1);
    for (int g = 0; g < ch->G; ++g) {
        const size_t o = (size_t)ch->off[g];
        if (part_get(ch->part[g], par ? par + o * npar : nullptr, OriginTraj ? OriginTraj + o * kp : nullptr,
                     LatentTraj ? LatentTraj + o * nO : nullptr, coalLog ? coalLog + o : nullptr,
                     logOrigin ? logOrigin + o : nullptr, counters ? counters + o * 4 : nullptr))
            return -1;
    }
    return 0;
}

extern "C" const double *lna_chains_summary_dev(lna_chains *ch) { return ch ? ch->summary : nullptr; }

extern "C" int lna_chains_exec_counts(lna_chains *ch, uint64_t *evals, uint64_t *warp_evals) {
    if (!ch) return fail("lna_chains_exec_counts: null chains");
    uint64_t tot[2] = {0, 0};
    for (int g = 0; g < ch->G; ++g) {
        ChainPart *pt = ch->part[g];
        cudaSetDevice(pt->pb->device);
        unsigned long long h[4] = {0, 0, 0, 0};
        CK(cudaStreamSynchronize(pt->pb->stream));
        CK(cudaMemcpy(h, pt->exec, sizeof h, cudaMemcpyDeviceToHost));
        tot[0] += h[0];
        tot[1] += h[1] + (h[2] + (unsigned long long)pt->pb->dp.k / 2) / (unsigned long long)pt->pb->dp.k;  // warp-ticks -> evaluations
    }
    if (evals) *evals = tot[0];
    if (warp_evals) *warp_evals = tot[1];
    return 0;
}
