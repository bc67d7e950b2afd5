// lna_eslice.cuh -- batched Metropolis / elliptical-slice state machines (sm_100a).
//
// Reference (file:line into MingweiWilliamTang/LNAphyloDyn):
//   Update_Param                 src/LNA_functional.cpp:1212-1243
//   Param_Slice_update(_all2)    src/LNA_functional.cpp:1246-1253, 1268-1314
//   ESlice_change_points         src/LNA_functional.cpp:1317-1409
//   ESlice_par_General           src/LNA_functional.cpp:1532-1688
//   ESlice_general_NC            src/LNA_functional.cpp:1690-1778
//   Update_Param_each_Norm       R/LNA_chpt_slice.R:169-265 (gamma random walk, no-op hyper entry)
//
// Key fact used here: the sequence of slice angles theta_1, theta_2, ... depends only on the
// pre-drawn uniforms (the bracket shrinks towards 0 by the SIGN of theta), never on likelihood
// values.  So W candidates of one chain are evaluated speculatively by W adjacent lanes and a
// warp ballot picks the first one the sequential loop would have accepted -- exactly the same
// result, at 1/W of the sequential latency.
#pragma once
#include "lna_kernels.cuh"
#include "lna_split.cuh"

namespace lna {

enum : int {
    PROP_PLAIN = 0, PROP_GAMMA_RW = 1, PROP_HYPER_NOOP = 2, PROP_ESS_PAR = 3, PROP_ESS_CHPT = 4, PROP_MH_PAIR = 5,
    PROP_MH_SCALAR = 6,  // one entry of Update_Param_each (R/LNA_chpt_slice.R:266-365)
    PROP_MH_HYPER = 7,   // update_hyper (R/general_change_point.R:281-313)
    PROP_MH_JOINT = 8,   // update_Param_joint, method "admcmc" (R/LNA_chpt_slice.R:9-88): General_MCMC2's stage i >= warmup2
    PROP_MH_TREE = 9     // up to 4 consecutive Metropolis entries (Update_Param_each + update_hyper) in ONE latency period
};

constexpr int kMaxUnif = 22;
constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;  // R's M_LN_SQRT_2PI

// R::dnorm(x, mu, sigma, log=TRUE)  (nmath/dnorm.c)
__device__ __forceinline__ double dnorm_log(double x, double mu, double sigma) {
    const double z = (x - mu) / sigma;
    return -(kLnSqrt2Pi + 0.5 * z * z + log(sigma));
}

// R::dunif(x, a, b, log=TRUE)  (nmath/dunif.c)
__device__ __forceinline__ double dunif_log(double x, double a, double b) {
    return (a <= x && x <= b) ? -log(b - a) : -INFINITY;
}
// dgamma(x, shape, rate, log=TRUE) in closed form (R evaluates it through dpois_raw; R is not available
// here, so this path is pinned CPU-restatement <-> GPU only)
__device__ __forceinline__ double dgamma_log(double x, double shape, double rate) {
    return shape * log(rate) - lgamma(shape) + (shape - 1.0) * log(x) - rate * x;
}

// Contraction-proof rotation a*cos + b*sin: evaluated identically wherever it is inlined, so the
// parameters a candidate was evaluated at and the ones committed to the chain state are the same bits.
__device__ __forceinline__ double rot(double a, double ct, double b, double sn) {
    return __dadd_rn(__dmul_rn(a, ct), __dmul_rn(b, sn));
}

// Everything a proposal needs besides per-change-point data.
struct PropScalars {
    double S0, I0, R0, gam;       // evaluated values
    double ct, sn;                // cos/sin(theta)
    double hyper_old, hyper_new;  // change-ratio scale before / after
    int rot_ch;                   // ESS_vec[5]; 0 also marks the "revert to the current parameters" candidate
    int noop;                     // MH pair: this lane applies ch <- exp(log(ch) * hyper / hyper)
};

// par layout (SIR): [S0, I0, R0, gamma, ch_1..ch_nch, hyper]  
template <int kProp>
struct ChProp {
    View par;
    View nrm;
    long long c;
    int off;  // index of ch_1 inside par
    PropScalars x;
    struct Raw {
        double c, nu;
    };
    __device__ __forceinline__ Raw raw(int j) const {
        Raw r;
        r.c = par.at(c, off + j);
        r.nu = 0.0;
        if (kProp == PROP_ESS_PAR) r.nu = nrm.at(c, 3 + j);
        if (kProp == PROP_ESS_CHPT) r.nu = nrm.at(c, j);
        return r;
    }
    __device__ __forceinline__ double eval(const Raw &r) const {
        const double cj = r.c;
        if (kProp == PROP_HYPER_NOOP || kProp == PROP_MH_HYPER) {
            // exp(log(ch) * hyper / hyper'), R/LNA_chpt_slice.R:238-241 with hyper' = hyper (ulp-level
            // perturbation, A.17) and update_hyper (general_change_point.R:292-293) with hyper' = hyper e^u
            return exp(__ddiv_rn(__dmul_rn(log(cj), x.hyper_old), x.hyper_new));
        } else if (kProp == PROP_MH_TREE) {
            // (the hyper entry of the tree: exactly PROP_MH_HYPER's expression)
            const double v = exp(__ddiv_rn(__dmul_rn(log(cj), x.hyper_old), x.hyper_new));
            return x.rot_ch ? v : cj;
        } else if (kProp == PROP_MH_JOINT) {
            // only when the hyper-parameter is among the jointly proposed ones (:65-68): exp(log(ch) * hyper / hyper')
            const double v = exp(__ddiv_rn(__dmul_rn(log(cj), x.hyper_old), x.hyper_new));
            return x.rot_ch ? v : cj;
        } else if (kProp == PROP_MH_PAIR) {
            // same instruction stream for all lanes of the pair kernel; the flag selects per lane
            const double pert = exp(__ddiv_rn(__dmul_rn(log(cj), x.hyper_old), x.hyper_old));
            return x.noop ? pert : cj;
        } else if (kProp == PROP_ESS_PAR) {
            if (!x.rot_ch) return cj;
            const double z = __dmul_rn(log(cj), x.hyper_old);  // :1296
            const double zp = rot(z, x.ct, r.nu, x.sn);        // :1299
            return exp(__ddiv_rn(zp, x.hyper_new));            // :1310
        } else if (kProp == PROP_ESS_CHPT) {
            const double nu = __ddiv_rn(r.nu, x.hyper_old);    // :1344
            const double v = exp(rot(log(cj), x.ct, nu, x.sn));  // :1249-1251
            return x.rot_ch ? v : cj;
        } else {
            return cj;
        }
    }
    __device__ __forceinline__ double ch(int j) const { return eval(raw(j)); }
};

// Slice-angle generator: replays the reference's bracket logic from the uniforms.
struct ThetaSeq {
    double th, lo, hi;
    int j;  // number of angles generated so far
    __device__ __forceinline__ void init() {
        th = kTwoPi;
        lo = 0.0;
        hi = kTwoPi;
        j = 0;
    }
    // advance to angle number `i` (1-based); unif(q) = q-th uniform of the call (0 = u, 1 = theta_1 ...)
    template <class U>
    __device__ __forceinline__ void advance(int i, const U &unif) {
        while (j < i) {
            ++j;
            if (j == 1) {
                th = 0.0 + (kTwoPi - 0.0) * unif(1);  // R::runif(0, 2*pi)
                lo = th - kTwoPi;
                hi = th;
            } else {
                if (th < 0.0) lo = th;
                else hi = th;
                th = lo + (hi - lo) * unif(j);  // R::runif(theta_min, theta_max)
            }
        }
    }
};

struct CandArgs {
    View par;        // chain state, npar per chain
    View origin;     // k*p per chain
    View normals;    // per-call normals
    View uniforms;   // per-call uniforms
    const double *coal_log;  // [B]
    double pr[8];    // priorList (4 x (mean, sd))
    int ess[7];
    double gamma_prior[2], gamma_prop;
    // PROP_MH_SCALAR: random walk on par[mh_index] (log scale if mh_is_log), prior dnorm (0) / dunif (1)
    // PROP_MH_HYPER : hyper' = hyper * exp(u), prior dgamma(shape = mh_pr[0], rate = mh_pr[1])
    int mh_index, mh_is_log, mh_prior_kind;
    double mh_pr[2], mh_prop;
    // PROP_MH_JOINT: raw' = raw + T z over the jd listed parameters (1 = I0, 2 = R0, 3 = gamma, 4 = hyper, the last one if
    // present), T = this chain's 4 x 4 row-major transform (eigenvectors * sqrt(eigenvalues) of its proposal covariance, as
    // MASS::mvrnorm builds it), z = jd normals at column jnorm_off of the normals stream
    int jd, jidx[4], jlog[4], jkind[4], jnorm_off;  // jkind: 0 dnorm, 1 dunif, 2 dgamma
    double jpr[4][2];
    const double *jT;  // [B][16]
    double *jnew;      // [B][4] proposed values on the natural scale (read by the commit kernel)
    // PROP_MH_TREE: tm consecutive Metropolis entries of General_MCMC2 (R/LNA_chpt_slice.R:266-365 + update_hyper). Entry e:
    // tkind 0 = scalar random walk on par[tidx] (like PROP_MH_SCALAR), 1 = hyper step (like PROP_MH_HYPER, last entry),
    // 2 = the no-op hyper entry (like PROP_HYPER_NOOP), 3 = hyper log-normal walk (PROP_MH_HYPER with prior kind 2);
    // uniforms: walk at column tu[e], accept at tu[e] + 1 of the stream
    int tm, tkind[4], tidx[4], tlog[4], tpk[4], tu[4];
    double tpr[4][2], tprop[4];
    FullOut cand;    // candidate scratch: slot-major (sb = 1, sj = NC)
    int *win_slot;   // [B] winning scratch slot, -1 = keep current state
    double *cand_ll; // [B] log-likelihood of the winner (or of the rejected MH proposal)
    double *theta;   // [B] accepted angle (ESS) / proposed log-gamma (MH gamma)
    int *n_evals;    // [B] evaluations the sequential reference loop would have made
    int *status;     // [B]
    int *accept;     // [B] MH accept flag (MH kinds)
    // two-phase slice evaluation: a launch covers candidates first_cand+1 .. ; after `max_waves` waves the
    // still-unresolved chains are appended to next_list (compaction) for a wider follow-up launch
    int first_cand;
    long long slot_base;    // scratch slots of this launch start here (phase B must not overwrite phase-A winners)
    int max_waves;          // 0 = until resolved
    const int *chain_list;  // null = chains 0..B-1
    const int *n_list;      // device count of chain_list entries
    int *next_list, *next_count;
    // measurement: [0] += evaluations executed (lanes that ran a FULL evaluation, speculative ones included),
    // [1] += warp-evaluations executed (what the FP64 pipe is charged: a warp costs the same whatever its lane mask)
    unsigned long long *exec_count;
    // ESS_vec[6] = 1 (self-mixing OriginTraj, :1576-1578, 1660-1662): the accepted candidate's scale of OriginTraj per chain
    // (output, may be null); such calls run one candidate per chain and wave (the scale depends on which earlier
    // candidates failed their Cholesky factorisation)
    double *eps_scale;
    // k_ess_par_async (lna_async.cuh): next chain of the batch to hand to a group of lanes (zeroed before the launch)
    int *work_counter;
};

// Whitened scalars of Param_Slice_update_all2 (:1287-1296) and their rotated images (:1299-1311).
struct EssParPre {
    double z0, z1, z2, zh;  // (log I0 - m)/s, (log R0 - m)/s, (log gamma - m)/s, (log hyper - m)/s
};

__device__ __forceinline__ void ess_par_scalars(const CandArgs &a, long long c, int npar, const EssParPre &z,
                                                double theta, PropScalars &x) {
    double sn, ct;
    sincos(theta, &sn, &ct);
    x.ct = ct;
    x.sn = sn;
    const double I0 = a.par.at(c, 1), R0 = a.par.at(c, 2), gam = a.par.at(c, 3), hy = a.par.at(c, npar - 1);
    x.S0 = a.par.at(c, 0);
    x.hyper_old = hy;
    x.I0 = a.ess[0] ? exp(__dadd_rn(__dmul_rn(rot(z.z0, ct, a.normals.at(c, 0), sn), a.pr[1]), a.pr[0])) : I0;
    x.R0 = a.ess[1] ? exp(__dadd_rn(__dmul_rn(rot(z.z1, ct, a.normals.at(c, 1), sn), a.pr[3]), a.pr[2])) : R0;
    x.gam = a.ess[2] ? exp(__dadd_rn(__dmul_rn(rot(z.z2, ct, a.normals.at(c, 2), sn), a.pr[5]), a.pr[4])) : gam;
    x.hyper_new =
        a.ess[4] ? exp(__dadd_rn(__dmul_rn(rot(z.zh, ct, a.normals.at(c, npar - 2), sn), a.pr[7]), a.pr[6])) : hy;
    x.rot_ch = a.ess[5];
}

__device__ __forceinline__ EssParPre ess_par_pre(const CandArgs &a, long long c, int npar) {
    EssParPre z;
    z.z0 = __ddiv_rn(__dsub_rn(log(a.par.at(c, 1)), a.pr[0]), a.pr[1]);
    z.z1 = __ddiv_rn(__dsub_rn(log(a.par.at(c, 2)), a.pr[2]), a.pr[3]);
    z.z2 = __ddiv_rn(__dsub_rn(log(a.par.at(c, 3)), a.pr[4]), a.pr[5]);
    z.zh = __ddiv_rn(__dsub_rn(log(a.par.at(c, npar - 1)), a.pr[6]), a.pr[7]);
    return z;
}

// ================================================================================================
// Candidate evaluation kernel.  W lanes per chain (W = 1 for the MH kinds); slot = global thread id.
//   PROP_GAMMA_RW / PROP_HYPER_NOOP : one FULL evaluation + MH test (Update_Param)
//   PROP_ESS_PAR  : ESlice_par_General  (candidates 1..20, then revert)
//   PROP_ESS_CHPT : ESlice_change_points (candidates 1..21, then revert; a Cholesky failure aborts)
// SIR (p = 2) only, like the reference's samplers (hard-wired subvec(0,1) / subvec(2,..)).
// ================================================================================================
// kSplit: every candidate is evaluated by a pipeline of three warps (lna_split.cuh); the block then holds blockDim/96 groups,
// warps 3j, 3j+1, 3j+2 = roles P, M, E of candidates [32 j, 32 j + 32).  The three warps run the same control flow below on
// the same (handed-back) likelihoods; only E writes results.
template <int kProp, int W, bool kSplit>
__global__ void __launch_bounds__(128, 3)
k_cand_eval(const __grid_constant__ DevProb P, const __grid_constant__ SegDt SD, long long B, CandArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp_in_block = threadIdx.x >> 5;
    const int role = kSplit ? (warp_in_block % kSplitWarps) : 2;  // 0 = P (mean ODE), 1 = M (moments), 2 = E / the only warp
    const int group = kSplit ? (warp_in_block / kSplitWarps) : 0;
    const bool writer = role == 2;
    SplitCtx sx{};
    if (kSplit) {
        unsigned char *group_smem = smem + full_eval_smem_bytes(tables_bytes(P.nch, P.k, P.n0), blockDim.x) +
                                    (size_t)group * split_group_bytes();
        sx = split_ctx(group_smem, threadIdx.x & 31);
        if (role == 0 && (threadIdx.x & 31) == 0) split_reset(sx);
    }
    const Tables T = stage_tables(P, smem, true);  // (block-wide barrier inside: the counters above are visible after it)
    const long long t0 = kSplit ? ((long long)blockIdx.x * (blockDim.x / (32 * kSplitWarps)) + group) * 32 + (threadIdx.x & 31)
                                : (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // one FULL evaluation of this lane's candidate; `on` = the lane really has one (the pipelined version runs all lanes,
    // because its hand-over is warp-wide, and only suppresses the stores of the idle ones)
    auto evaluate = [&](auto &chs, const auto &eps, const PropScalars &x, long long slot, bool on, double &ll, int &st) {
        if (a.exec_count && writer) {
            const unsigned am = __activemask(), m = __ballot_sync(am, on);
            if (m && (threadIdx.x & 31) == __ffs(am) - 1) {
                atomicAdd(a.exec_count, (unsigned long long)__popc(m));
                atomicAdd(a.exec_count + 1, 1ull);
            }
        }
        if (kSplit) {
            if (role == 0) sir_split_P(sx, P, T, x.R0, x.gam, x.S0, x.I0, ll, st);
            else if (role == 1) sir_split_M(sx, P, SD, T, x.R0, x.gam, ll, st);
            else sir_split_E<true>(sx, P, T, chs, eps, x.R0, x.gam, x.S0, x.I0, a.cand, slot, on, ll, st);
            if (!on) {
                ll = kSentinel;
                st = 0;
            }
        } else if (on) {
            sir_full_eval<true>(P, SD, T, chs, eps, x.R0, x.gam, x.S0, x.I0, a.cand, slot, ll, st);
        }
    };
    const long long ci = t0 / W;  // index into the (optional) compacted chain list
    const int w = (int)(t0 % W);
    const long long t = a.slot_base + t0;  // scratch slot of this thread's candidate
    bool valid = ci < B;
    long long c = ci;
    if (a.chain_list) {
        valid = ci < (long long)__ldg(a.n_list);
        c = valid ? (long long)__ldg(a.chain_list + ci) : 0;
    }
    const int lane = threadIdx.x & 31;
    const unsigned gmask = (W >= 32) ? 0xffffffffu : (((1u << W) - 1u) << (lane & ~(W - 1)));
    const int npar = 2 + P.npt;
    const long long cc = valid ? c : 0;  // keep loads in range for idle lanes
    constexpr bool kIsEss = (kProp == PROP_ESS_PAR || kProp == PROP_ESS_CHPT);
    constexpr int kMaxCand = (kProp == PROP_ESS_PAR) ? 20 : (kProp == PROP_ESS_CHPT ? 21 : 1);

    auto unif = [&](int q) { return a.uniforms.at(cc, q); };
    EpsFromView eps{a.origin, cc, P.k};

    if (kProp == PROP_MH_TREE) {
        // ---- up to four consecutive Metropolis entries in ONE latency period (W = 2^tm lanes per chain) --------------
        // Entry e's proposal touches its own parameter only and is drawn from its own uniform, so what the sequential code
        // evaluates at entry e is determined by WHICH earlier entries were accepted: lane l = 2^e - 1 + A evaluates entry e
        // on top of the accepted subset A of entries 0..e-1 (1 + 2 + 4 + 8 = 15 candidates), and the accept tests are then
        // chained exactly as the sequential code chains them.  Same evaluations, same decisions, 1/tm of the latency.
        const int m = a.tm;
        const int e_l = 31 - __clz(w + 1), A_l = (w + 1) - (1 << e_l);  // this lane's (entry, accepted subset)
        const bool used = w < (1 << m) - 1;
        PropScalars x;
        x.S0 = a.par.at(cc, 0);
        x.I0 = a.par.at(cc, 1);
        x.R0 = a.par.at(cc, 2);
        x.gam = a.par.at(cc, 3);
        x.hyper_old = x.hyper_new = a.par.at(cc, npar - 1);
        x.ct = 1.0;
        x.sn = 0.0;
        x.rot_ch = 0;
        x.noop = 0;
        double newv[4], off[4];
        for (int e = 0; e < m; ++e) {  // every lane computes all proposals (a few scalar operations each)
            if (a.tkind[e] == 0) {
                const int j = a.tidx[e];
                const double cur = a.par.at(cc, j);
                const double raw = a.tlog[e] ? log(cur) : cur, po = a.tprop[e];
                const double rnew = raw + (-po + (po - (-po)) * unif(a.tu[e]));
                off[e] = a.tpk[e] == 0 ? dnorm_log(rnew, a.tpr[e][0], a.tpr[e][1]) - dnorm_log(raw, a.tpr[e][0], a.tpr[e][1])
                                       : dunif_log(rnew, a.tpr[e][0], a.tpr[e][1]) - dunif_log(raw, a.tpr[e][0], a.tpr[e][1]);
                newv[e] = a.tlog[e] ? exp(rnew) : rnew;
            } else if (a.tkind[e] == 2) {
                newv[e] = x.hyper_old;  // no proposal, offset 0: only ch <- exp(log(ch) * hyper / hyper)
                off[e] = 0.0;
            } else {
                const double hp = a.tprop[e], uu = -hp + (hp - (-hp)) * unif(a.tu[e]);
                newv[e] = x.hyper_old * exp(uu);
                if (a.tkind[e] == 3) {
                    const double ln = log(newv[e]), lo = log(x.hyper_old);
                    off[e] = (dnorm_log(ln, a.tpr[e][0], a.tpr[e][1]) - ln) - uu - (dnorm_log(lo, a.tpr[e][0], a.tpr[e][1]) - lo);
                } else {
                    off[e] = dgamma_log(newv[e], a.tpr[e][0], a.tpr[e][1]) - dgamma_log(x.hyper_old, a.tpr[e][0], a.tpr[e][1]) + uu;
                }
            }
        }
        for (int e = 0; e < m; ++e) {
            const bool on_e = used && (e == e_l || (e < e_l && ((A_l >> e) & 1)));
            if (!on_e) continue;
            if (a.tkind[e] == 0) {
                if (a.tidx[e] == 0) x.S0 = newv[e];
                if (a.tidx[e] == 1) x.I0 = newv[e];
                if (a.tidx[e] == 2) x.R0 = newv[e];
                if (a.tidx[e] == 3) x.gam = newv[e];
            } else {
                x.hyper_new = newv[e];
                x.rot_ch = 1;
            }
        }
        double ll = kSentinel;
        int st = 0;
        {
            ChProp<PROP_MH_TREE> chs{a.par, a.normals, cc, 4, x};
            evaluate(chs, eps, x, t, valid && used, ll, st);
        }
        // chained accept tests (every lane of the group computes them: the shuffles are warp-wide)
        const int base = lane & ~(W - 1);
        double cur = valid ? a.coal_log[cc] : 0.0;
        int A = 0, last = -1, stc = 0;
        for (int e = 0; e < m; ++e) {
            const int src = base + (1 << e) - 1 + A;
            const double ll_e = __shfl_sync(0xffffffffu, ll, src);
            const int st_e = __shfl_sync(0xffffffffu, st, src);
            const bool acc = !(st_e & ST_CHOL_FAIL) && (log(unif(a.tu[e] + 1)) < ll_e - cur + off[e]);
            stc |= acc ? st_e : (st_e & (ST_CHOL_FAIL | ST_CAP_HIT));
            if (acc) {
                cur = ll_e;
                last = (1 << e) - 1 + A;
                A |= 1 << e;
            }
        }
        if (valid && w == 0 && writer) {
            a.win_slot[c] = last >= 0 ? (int)t + last : -1;
            a.cand_ll[c] = cur;
            a.theta[c] = 0.0;
            a.n_evals[c] = m;
            a.status[c] = stc;
            a.accept[c] = A;
            for (int e = 0; e < m; ++e) a.jnew[c * 4 + e] = newv[e];
        }
        return;
    } else if (kProp == PROP_MH_PAIR) {
        // ---- both entries of Update_Param_each_Norm in ONE latency period (W = 4 lanes per chain) -----
        //   lane 0: gamma' proposal, current ch            (entry c)
        //   lane 1: gamma' accepted, ch <- exp(log ch*h/h) (entry d after an accepted c)
        //   lane 2: gamma  kept,     ch <- exp(log ch*h/h) (entry d after a rejected c)
        // The MH tests are then chained exactly as the sequential code would (uniforms 1 and 2).
        PropScalars x;
        x.S0 = a.par.at(cc, 0);
        x.I0 = a.par.at(cc, 1);
        x.R0 = a.par.at(cc, 2);
        x.gam = a.par.at(cc, 3);
        x.hyper_old = x.hyper_new = a.par.at(cc, npar - 1);
        x.ct = 1.0;
        x.sn = 0.0;
        x.rot_ch = 0;
        const double raw = log(x.gam), po = a.gamma_prop;
        const double rnew = raw + (-po + (po - (-po)) * unif(0));
        const double offset =
            dnorm_log(rnew, a.gamma_prior[0], a.gamma_prior[1]) - dnorm_log(raw, a.gamma_prior[0], a.gamma_prior[1]);
        if (w <= 1) x.gam = exp(rnew);
        x.noop = (w >= 1);
        double ll = kSentinel;
        int st = 0;
        {  // one instruction stream for the three lanes (no divergence)
            ChProp<PROP_MH_PAIR> chs{a.par, a.normals, cc, 4, x};
            evaluate(chs, eps, x, t, valid && w <= 2, ll, st);
        }
        const int base = lane & ~3;
        const double ll0 = __shfl_sync(0xffffffffu, ll, base), ll1 = __shfl_sync(0xffffffffu, ll, base + 1),
                     ll2 = __shfl_sync(0xffffffffu, ll, base + 2);
        const int st0 = __shfl_sync(0xffffffffu, st, base), st1 = __shfl_sync(0xffffffffu, st, base + 1),
                  st2 = __shfl_sync(0xffffffffu, st, base + 2);
        if (valid && w == 0 && writer) {
            const double cl = a.coal_log[c];
            const bool acc1 = !(st0 & ST_CHOL_FAIL) && (log(unif(1)) < ll0 - cl + offset);
            const double cur = acc1 ? ll0 : cl;
            const double lln = acc1 ? ll1 : ll2;
            const int stn = acc1 ? st1 : st2;
            const bool acc2 = !(stn & ST_CHOL_FAIL) && (log(unif(2)) < lln - cur + 0.0);
            const int slot = acc2 ? (int)t + (acc1 ? 1 : 2) : (acc1 ? (int)t : -1);
            a.win_slot[c] = slot;
            a.cand_ll[c] = acc2 ? lln : cur;
            a.theta[c] = rnew;
            a.n_evals[c] = 2;
            a.status[c] = acc2 ? stn : (acc1 ? st0 : ((st0 | stn) & ST_CHOL_FAIL));
            a.accept[c] = (acc1 ? 1 : 0) | (acc2 ? 2 : 0);
        }
        return;
    } else if (!kIsEss) {
        // ---- Metropolis step: Update_Param_each_Norm entry + Update_Param ------------------------
        if (!valid && !kSplit) return;
        c = cc;  // (idle lanes of a pair warp evaluate chain 0 and write nothing)
        PropScalars x;
        x.S0 = a.par.at(c, 0);
        x.I0 = a.par.at(c, 1);
        x.R0 = a.par.at(c, 2);
        x.gam = a.par.at(c, 3);
        x.hyper_old = x.hyper_new = a.par.at(c, npar - 1);
        x.ct = 1.0;
        x.sn = 0.0;
        x.rot_ch = 0;
        x.noop = 0;
        double offset = 0.0, prop = 0.0;
        int iu = 0;
        if (kProp == PROP_GAMMA_RW) {
            // log gamma' = log gamma + runif(-po, po); offset = dnorm(new) - dnorm(old)  (:217-218)
            const double raw = log(x.gam), po = a.gamma_prop;
            const double rnew = raw + (-po + (po - (-po)) * unif(iu++));
            offset = dnorm_log(rnew, a.gamma_prior[0], a.gamma_prior[1]) - dnorm_log(raw, a.gamma_prior[0], a.gamma_prior[1]);
            x.gam = exp(rnew);
            prop = rnew;
        }
        if (kProp == PROP_MH_SCALAR) {
            const int j = a.mh_index;
            const double cur = a.par.at(c, j);
            const double raw = a.mh_is_log ? log(cur) : cur, po = a.mh_prop;
            const double rnew = raw + (-po + (po - (-po)) * unif(iu++));
            offset = a.mh_prior_kind == 0
                         ? dnorm_log(rnew, a.mh_pr[0], a.mh_pr[1]) - dnorm_log(raw, a.mh_pr[0], a.mh_pr[1])
                         : dunif_log(rnew, a.mh_pr[0], a.mh_pr[1]) - dunif_log(raw, a.mh_pr[0], a.mh_pr[1]);
            const double val = a.mh_is_log ? exp(rnew) : rnew;
            if (j == 0) x.S0 = val;
            if (j == 1) x.I0 = val;
            if (j == 2) x.R0 = val;
            if (j == 3) x.gam = val;
            prop = rnew;
        }
        if (kProp == PROP_MH_JOINT) {
            // RawTransParam_new = mvrnorm(1, RawTransParam, PCOV) (:37-39), the prior terms in list order (:41-61), the
            // change ratios rescaled when the hyper-parameter moves (:65-68); one uniform: the accept test of Update_Param
            double off = 0.0;
            const double *Tm = a.jT + c * 16;
            for (int i = 0; i < a.jd; ++i) {
                const int j = a.jidx[i] == 4 ? npar - 1 : a.jidx[i];
                const double cur = a.par.at(c, j);
                const double raw = a.jlog[i] ? log(cur) : cur;
                double rnew = raw;
                for (int q = 0; q < a.jd; ++q) rnew = __dadd_rn(rnew, __dmul_rn(Tm[i * 4 + q], a.normals.at(c, a.jnorm_off + q)));
                const double p0 = a.jpr[i][0], p1 = a.jpr[i][1];
                off += a.jkind[i] == 0   ? dnorm_log(rnew, p0, p1) - dnorm_log(raw, p0, p1)
                       : a.jkind[i] == 1 ? dunif_log(rnew, p0, p1) - dunif_log(raw, p0, p1)
                                         : dgamma_log(rnew, p0, p1) - dgamma_log(raw, p0, p1);
                const double val = a.jlog[i] ? exp(rnew) : rnew;
                if (a.jidx[i] == 1) x.I0 = val;
                if (a.jidx[i] == 2) x.R0 = val;
                if (a.jidx[i] == 3) x.gam = val;
                if (a.jidx[i] == 4) {
                    x.hyper_new = val;
                    x.rot_ch = 1;
                }
                if (valid && writer) a.jnew[c * 4 + i] = val;
            }
            offset = off;
            iu = 1;  // uniform 0 of the slot is the (unused) walk draw of the scalar entries
            prop = offset;
        }
        if (kProp == PROP_MH_HYPER) {
            const double hp = a.mh_prop, uu = -hp + (hp - (-hp)) * unif(iu++);
            x.hyper_new = x.hyper_old * exp(uu);
            if (a.mh_prior_kind == 2) {
                // Update_Param_each_Norm with hyper = T (:219-227): dlnorm(new) - u - dlnorm(old)
                const double ln = log(x.hyper_new), lo = log(x.hyper_old);
                offset = (dnorm_log(ln, a.mh_pr[0], a.mh_pr[1]) - ln) - uu - (dnorm_log(lo, a.mh_pr[0], a.mh_pr[1]) - lo);
            } else {  // update_hyper of General_MCMC2 (general_change_point.R:296-299)
                offset = dgamma_log(x.hyper_new, a.mh_pr[0], a.mh_pr[1]) - dgamma_log(x.hyper_old, a.mh_pr[0], a.mh_pr[1]) + uu;
            }
            prop = x.hyper_new;
        }
        ChProp<kProp> chs{a.par, a.normals, c, 4, x};
        double ll = kSentinel;
        int st = 0;
        evaluate(chs, eps, x, t, valid, ll, st);
        if (!valid || !writer) return;
        const double u = unif(iu);
        const double acc_a = ll - a.coal_log[c] + offset;
        // a Cholesky failure is an exception in the reference: tryCatch keeps the old state
        const bool acc = !(st & ST_CHOL_FAIL) && (log(u) < acc_a);
        a.win_slot[c] = acc ? (int)t : -1;
        a.cand_ll[c] = ll;
        a.theta[c] = prop;
        a.n_evals[c] = 1;
        a.status[c] = st;
        a.accept[c] = acc ? 1 : 0;
        return;
    } else {
        // ---- elliptical slice: speculative waves of W candidates ---------------------------------
        EssParPre zpre{};
        double hyper_old = 0.0, logy = 0.0;
        if (valid) {
            if (kProp == PROP_ESS_PAR) zpre = ess_par_pre(a, c, npar);
            hyper_old = a.par.at(c, npar - 1);
            logy = a.coal_log[c] + log(unif(0));
        }
        ThetaSeq ts;
        ts.init();
        bool done = !valid;
        // ESS_vec[6] = 1: OriginTraj_new = OriginTraj cos(theta) + OriginTraj_new sin(theta) accumulates over the candidates
        // whose New_Param_List succeeded, i.e. candidate i is evaluated on c_i OriginTraj, c_i = cos(theta_i) + c_{i-1} sin(theta_i)
        constexpr bool kSelfMix = (kProp == PROP_ESS_PAR && W == 1 && !kSplit);
        double eps_c = 1.0;
        for (int wave = 0;; ++wave) {
            const int i = a.first_cand + wave * W + w + 1;  // candidate number in the sequential loop
            // candidate kMaxCand+1 is the reference's cap: revert to the current parameters and recompute
            // (:1609-1622 / :1363-1374).  It is evaluated speculatively in the same wave as the last shrinks.
            const bool is_revert = (i == kMaxCand + 1);
            const bool active = !done && i <= kMaxCand + 1;
            double ll = kSentinel;
            int st = 0;
            const bool any_active = kSplit ? __any_sync(0xffffffffu, active) != 0 : false;  // (all lanes vote)
            if (active || any_active) {
                PropScalars x;
                if (active && !is_revert) ts.advance(i, unif);
                const double th = (is_revert || !active) ? 0.0 : ts.th;
                if (kProp == PROP_ESS_PAR) {
                    ess_par_scalars(a, cc, npar, zpre, th, x);
                } else {
                    double sn, ct;
                    sincos(th, &sn, &ct);
                    x.ct = ct;
                    x.sn = sn;
                    x.hyper_old = x.hyper_new = hyper_old;
                    x.rot_ch = 1;
                }
                x.S0 = a.par.at(cc, 0);
                if (kProp != PROP_ESS_PAR || is_revert || !active) {  // current parameters, bit for bit
                    x.I0 = a.par.at(cc, 1);
                    x.R0 = a.par.at(cc, 2);
                    x.gam = a.par.at(cc, 3);
                }
                if (is_revert || !active) x.rot_ch = 0;
                x.noop = 0;
                // one instruction stream for shrink and revert candidates (no divergence inside the warp)
                ChProp<kProp> chs{a.par, a.normals, cc, 4, x};
                bool mixed = false;
                if constexpr (kSelfMix) {
                    if (a.ess[6]) {
                        const double c_try = (active && !is_revert) ? fma(eps_c, x.sn, x.ct) : 1.0;
                        EpsScaled epsS{eps, c_try};
                        evaluate(chs, epsS, x, t, active, ll, st);
                        if (active && !is_revert && !(st & ST_CHOL_FAIL)) eps_c = c_try;  // (a failed candidate leaves it, :1641-1651)
                        if (is_revert) eps_c = 1.0;
                        mixed = true;
                    }
                }
                if (!mixed) evaluate(chs, eps, x, t, active, ll, st);
            }
            const bool failed = active && (st & ST_CHOL_FAIL);
            // ESlice_par_General catches the exception and shrinks on (:1641-1651); ESlice_change_points
            // does not catch it, so the first failing candidate aborts the update.  The revert candidate
            // always ends the loop.
            const bool stop_here =
                active && (is_revert || (ll > logy && !failed) || (kProp == PROP_ESS_CHPT && failed));
            const unsigned m = __ballot_sync(0xffffffffu, stop_here) & gmask;
            if (!done && m) {
                const int winner = __ffs(m) - 1;
                if (lane == winner && writer) {
                    a.win_slot[c] = failed ? -1 : (int)t;
                    a.cand_ll[c] = ll;
                    a.theta[c] = is_revert ? 0.0 : ts.th;
                    a.n_evals[c] = i;
                    a.status[c] = st | (is_revert ? ST_CAP_HIT : 0);
                    if (kSelfMix && a.eps_scale) a.eps_scale[c] = eps_c;
                }
                done = true;
            }
            if (!done && a.max_waves && wave + 1 == a.max_waves) {
                if (w == 0 && writer) a.next_list[atomicAdd(a.next_count, 1)] = (int)c;  // hand over to the next phase
                done = true;
            }
            if (!__any_sync(0xffffffffu, !done)) break;
        }
    }
}

// ================================================================================================
// Commit: copy the winning candidate's FT / coarse ODE / betaN / LatentTraj from scratch into the
// destination (chain state or the caller's output arrays) and materialise the new parameter vector.
// One thread per (chain, element).
// ================================================================================================
struct CommitArgs {
    const int *win_slot;
    long long NC;                       // scratch slots
    const double *cA, *cL, *cOde, *cBeta, *cLat;  // candidate scratch (slot-major)
    OutView A, L, ode, betaN, latent;  // destinations
    int nA, nOde, nBeta;               // elements per item: p*p*k, nrow*(p+1), nrow
    int copy_when_rejected;            // standalone entry points also want the evaluated candidate back
};

constexpr int kCommitPer = 8;  // elements per thread
__global__ void k_commit_fields(long long B, CommitArgs a) {
    // grid.y = block of kCommitPer consecutive elements, grid.x covers the chains (no 64-bit div/mod per thread).  A thread
    // first issues its kCommitPer gather loads -- independent, all in flight together -- then stores: one element per thread
    // left the copy launch- and latency-bound (29 us for 4096 chains, 203 us for 65 536: a third of the HBM rate).
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    const int slot = a.win_slot[c];
    if (slot < 0) return;
    const int per = 2 * a.nA + 2 * a.nOde + a.nBeta, e0 = blockIdx.y * kCommitPer;
    // field of flat element e: A | L | ode | latent | betaN
    auto src = [&](int e) -> const double * {
        if (e < a.nA) return a.A.ok() ? a.cA + (long long)e * a.NC + slot : nullptr;
        e -= a.nA;
        if (e < a.nA) return a.L.ok() ? a.cL + (long long)e * a.NC + slot : nullptr;
        e -= a.nA;
        if (e < a.nOde) return a.ode.ok() ? a.cOde + (long long)e * a.NC + slot : nullptr;
        e -= a.nOde;
        if (e < a.nOde) return a.latent.ok() ? a.cLat + (long long)e * a.NC + slot : nullptr;
        e -= a.nOde;
        return a.betaN.ok() ? a.cBeta + (long long)e * a.NC + slot : nullptr;
    };
    double v[kCommitPer];
    bool have[kCommitPer];
#pragma unroll
    for (int i = 0; i < kCommitPer; ++i) {
        const double *q = e0 + i < per ? src(e0 + i) : nullptr;
        have[i] = q != nullptr;
        v[i] = have[i] ? __ldcs(q) : 0.0;  // (scratch is read once: streaming load)
    }
#pragma unroll
    for (int i = 0; i < kCommitPer; ++i) {
        if (!have[i]) continue;
        int e = e0 + i;
        if (e < a.nA) {
            a.A.put(c, e, v[i]);
            continue;
        }
        e -= a.nA;
        if (e < a.nA) {
            a.L.put(c, e, v[i]);
            continue;
        }
        e -= a.nA;
        if (e < a.nOde) {
            a.ode.put(c, e, v[i]);
            continue;
        }
        e -= a.nOde;
        if (e < a.nOde) {
            a.latent.put(c, e, v[i]);
            continue;
        }
        e -= a.nOde;
        a.betaN.put(c, e, v[i]);
    }
}

// logOrigin of ESlice_par_General: -1/2 sum over ALL rows of OriginTraj (:1673-1678); one warp per chain.
__global__ void k_log_origin(long long B, int n, View origin, double *logOrigin) {
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= B) return;
    double lo = 0.0;
    for (int e = lane; e < n; e += 32) {
        const double v = origin.at(c, e);
        lo -= 0.5 * v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo += __shfl_xor_sync(0xffffffffu, lo, o);
    if (lane == 0) logOrigin[c] = lo;
}

// LogChProb of ESlice_change_points: -1/2 sum (log(ch) * hyper)^2 (:1395-1399); one warp per chain.
__global__ void k_log_chprob(long long B, int nch, int npar, View par, double *out) {
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= B) return;
    const double hy = par.at(c, npar - 1);
    double acc = 0.0;
    for (int j = lane; j < nch; j += 32) {
        const double d = log(par.at(c, 4 + j)) * hy;
        acc += -0.5 * d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[c] = acc;
}

// ================================================================================================
// ESlice_general_NC (volz likelihood): CHEAP evaluations only (rotation of the latent increments,
// TransformTraj, per-cell coalescent sum).  ONE WARP PER CHAIN.
//
// TransformTraj is linear in the increments: with f' = f cos(th) + v sin(th),
//     X(th) = cos(th) * X_f + sin(th) * X_v,
// where X_f / X_v are the latent deviations driven by f / v alone.  The warp computes the two
// recurrences once (lanes 0 and 1), after which every slice candidate is a lane-parallel map over
// the coarse rows plus a warp-shuffle reduction of the per-cell coalescent terms.
// ================================================================================================
struct TrajArgs {
    View f_cur, normals, uniforms;      // k*p, k*p (column-major k x p), 22
    View A, L, ode, betaN;              // FT, coarse ODE, betaN of the current parameters
    const double *coal_log;             // [B]
    OutView latent, origin_new;         // outputs (may alias the chain state: per-lane read-before-write)
    double *CoalLog, *logOrigin;
    int *n_evals, *status;
    // mode 0: ESlice_general_NC (non-centred increments f, v; recurrences through FT)
    // mode 1: ESlice_general2 (:1783-1857): f_cur and `normals` are (k+1) x (p+1) TRAJECTORIES (current and the
    //         centred proposal drawn by Traj_sim_general3); the deviations from the ODE are rotated directly
    int mode;
    // volz = FALSE (:1717-1721, 1733-1760): the Kingman likelihood coal_loglik3(LogTraj(newTraj), lambda) -- log(log I), NaN
    // as soon as I < 1, and a NaN ends the loop -- with scale lambda[c]; the 20-cap still recomputes with volz_loglik_nh2
    int kingman;
    const double *lambda;
};

constexpr int kTrajWarps = 4;
__host__ __device__ inline size_t traj_warp_doubles(int k, int nrow) { return (size_t)7 * nrow + (size_t)8 * k; }

__global__ void __launch_bounds__(kTrajWarps * 32) k_eslice_traj(DevProb P, long long B, TrajArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables T = stage_tables(P, smem);
    const int k = P.k, nrow = P.nrow;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * kTrajWarps + w;
    if (c >= B) return;  // whole warps exit together
    double *ws = reinterpret_cast<double *>(smem + tables_bytes(P.nch, P.k, P.n0)) + (size_t)w * traj_warp_doubles(k, nrow);
    double *oS = ws, *oI = oS + nrow, *bet = oI + nrow, *XfS = bet + nrow, *XfI = XfS + nrow, *XvS = XfI + nrow,
           *XvI = XvS + nrow;
    double *Am = XvI + nrow, *cf = Am + 4 * k, *cv = cf + 2 * k;
    const unsigned full = 0xffffffffu;

    for (int r = lane; r < nrow; r += 32) {
        oS[r] = a.ode.at(c, nrow + r);
        oI[r] = a.ode.at(c, 2 * nrow + r);
        bet[r] = a.betaN.at(c, r);
    }
    if (a.mode == 1) {
        for (int r = lane; r < nrow; r += 32) {
            XfS[r] = a.f_cur.at(c, nrow + r) - oS[r];
            XfI[r] = a.f_cur.at(c, 2 * nrow + r) - oI[r];
            XvS[r] = a.normals.at(c, nrow + r) - oS[r];
            XvI[r] = a.normals.at(c, 2 * nrow + r) - oI[r];
        }
    }
    for (int i = lane; i < k && a.mode == 0; i += 32) {
        const double l00 = a.L.at(c, 4 * i), l10 = a.L.at(c, 4 * i + 1), l11 = a.L.at(c, 4 * i + 3);
        const double f0 = a.f_cur.at(c, i), f1 = a.f_cur.at(c, k + i);
        const double v0 = a.normals.at(c, i), v1 = a.normals.at(c, k + i);
        cf[2 * i] = l00 * f0;
        cf[2 * i + 1] = fma(l10, f0, l11 * f1);
        cv[2 * i] = l00 * v0;
        cv[2 * i + 1] = fma(l10, v0, l11 * v1);
#pragma unroll
        for (int e = 0; e < 4; ++e) Am[4 * i + e] = a.A.at(c, 4 * i + e);
    }
    __syncwarp();
    if (lane < 2 && a.mode == 0) {  // lane 0: recurrence driven by f, lane 1: by v
        const double *cc = lane == 0 ? cf : cv;
        double *XS = lane == 0 ? XfS : XvS, *XI = lane == 0 ? XfI : XvI;
        double x0 = 0.0, x1 = 0.0;
        XS[0] = 0.0;
        XI[0] = 0.0;
        for (int i = 0; i < k; ++i) {
            const double n0 = fma(Am[4 * i], x0, Am[4 * i + 2] * x1) + cc[2 * i];
            const double n1 = fma(Am[4 * i + 1], x0, Am[4 * i + 3] * x1) + cc[2 * i + 1];
            x0 = n0;
            x1 = n1;
            XS[i + 1] = x0;
            XI[i + 1] = x1;
        }
    }
    __syncwarp();

    auto unif = [&](int q) { return a.uniforms.at(c, q); };
    // one CHEAP evaluation at angle (ct, sn): returns the reference's volz_loglik_nh2 value and whether
    // ANY latent state (all rows, both columns) is negative (the extra loop condition, :1738)
    auto cheap = [&](double ct, double sn, bool &negall, int &st) -> double {
        double acc = 0.0;
        int neg_any = 0, neg_chk = 0, nan = 0;
        for (int r = lane; r < nrow; r += 32) {
            const double S = oS[r] + fma(ct, XfS[r], sn * XvS[r]);
            const double I = oI[r] + fma(ct, XfI[r], sn * XvI[r]);
            const int neg = (S < 0.0) | (I < 0.0);
            neg_any |= neg;
            if (r <= P.n0) neg_chk |= neg;
            const int i = P.Lrow - r;
            if (i >= 0 && i < P.n0 && T.cellCnt[i] > 0) {
                const double ratio = S / I;
                const double term = fma(-2.0 * T.cellCD[i] * bet[r], ratio, T.cellY[i] * log(bet[r] * ratio));
                if (!(S > 0.0) || !(I > 0.0) || term != term) nan = 1;
                acc += term;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
        negall = __any_sync(full, neg_any);
        st = 0;
        if (__any_sync(full, neg_chk)) {
            st = ST_NEG_STATE;
            return kSentinel;
        }
        if (__any_sync(full, nan)) {
            st = ST_NAN;
            return kSentinel;
        }
        return acc;
    };

    // coal_loglik3 (:1016-1054) on LogTraj(newTraj) (SIR_LNA.cpp:40-46): cell c uses row n0k - c, f = log(log I)
    const double lam = a.kingman ? a.lambda[c] : 1.0, llam = log(lam);
    auto kingman = [&](double ct, double sn, bool &negall) -> double {
        double acc = 0.0;
        int neg_any = 0;
        for (int r = lane; r < nrow; r += 32) {
            const double S = oS[r] + fma(ct, XfS[r], sn * XvS[r]);
            const double I = oI[r] + fma(ct, XfI[r], sn * XvI[r]);
            neg_any |= (S < 0.0) | (I < 0.0);
            const int i = P.Lrow - r;
            if (i >= 0 && i < P.n0 && r >= 1 && T.cellCnt[i] > 0) {
                const double f = log(log(I));
                acc += fma(-lam * T.cellCD[i], exp(-f), T.cellY[i] * (llam - f));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
        negall = __any_sync(full, neg_any);
        return acc;
    };
    int evals = 0, st = 0;
    bool negall;
    double logy;
    const double lu = log(unif(0));
    if (a.mode == 0 && a.coal_log[c] != -99999999.0) {
        logy = a.coal_log[c] + lu;
    } else {  // recompute the current likelihood like the reference (:1717-1719; always in ESlice_general2)
        logy = cheap(1.0, 0.0, negall, st) + lu;
        evals++;
    }
    ThetaSeq ts;
    ts.init();
    ts.advance(1, unif);
    double sn, ct;
    sincos(ts.th, &sn, &ct);
    double ll = a.kingman ? kingman(ct, sn, negall) : cheap(ct, sn, negall, st);
    evals++;
    int i = 0;
    bool reverted = false;
    while (negall || ll <= logy) {  // (a NaN Kingman value compares false and ends the loop, like the reference)
        i += 1;
        if (i > 20) {  // cap: back to the current increments, recompute (:1741-1747)
            reverted = true;
            ct = 1.0;
            sn = 0.0;
            if (a.mode == 0) {  // ESlice_general2 just returns f_cur, without another evaluation (:1835-1839)
                ll = cheap(ct, sn, negall, st);
                evals++;
            }
            st |= ST_CAP_HIT;
            break;
        }
        ts.advance(i + 1, unif);
        sincos(ts.th, &sn, &ct);
        ll = a.kingman ? kingman(ct, sn, negall) : cheap(ct, sn, negall, st);
        evals++;
    }
    if (a.kingman && ll != ll) st |= ST_NAN;
    // outputs: latent rows, rotated increments, logOrigin without the LAST row (:1768)
    for (int r = lane; r < nrow; r += 32) {
        const bool keep = a.mode == 1 && reverted;  // newTraj = f_cur, bit for bit
        a.latent.put(c, r, a.mode == 1 ? a.f_cur.at(c, r) : a.ode.at(c, r));
        a.latent.put(c, nrow + r, keep ? a.f_cur.at(c, nrow + r) : oS[r] + fma(ct, XfS[r], sn * XvS[r]));
        a.latent.put(c, 2 * nrow + r, keep ? a.f_cur.at(c, 2 * nrow + r) : oI[r] + fma(ct, XfI[r], sn * XvI[r]));
    }
    double lo = 0.0;
    for (int e = lane; e < 2 * k && a.mode == 0; e += 32) {
        const double f = a.f_cur.at(c, e);
        const double v = reverted ? f : rot(f, ct, a.normals.at(c, e), sn);
        a.origin_new.put(c, e, v);
        if ((e % k) < k - 1) lo -= 0.5 * v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo += __shfl_xor_sync(full, lo, o);
    if (lane == 0) {
        if (a.CoalLog) a.CoalLog[c] = ll;
        if (a.logOrigin) a.logOrigin[c] = lo;
        a.n_evals[c] = evals;
        a.status[c] = st;
    }
}

// New parameter vector of the accepted proposal, one thread per (chain, entry).  par_out must NOT
// alias the chain's current parameters (the resident chains double-buffer `par`).
template <int kProp>
__global__ void k_commit_par_wide(DevProb P, long long B, CandArgs a, OutView par_out, double *coalLog) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int npar = 2 + P.npt;
    if (idx >= B * npar) return;
    const long long c = idx % B;
    const int j = (int)(idx / B);
    const double old = a.par.at(c, j);
    const bool win = a.win_slot[c] >= 0;
    if (win && j == 0 && coalLog) coalLog[c] = a.cand_ll[c];
    const bool reverted = a.status[c] & ST_CAP_HIT;
    if (!win || reverted) {
        par_out.put(c, j, old);
        return;
    }
    const bool is_ch = j >= 4 && j < npar - 1;
    double v = old;
    if (kProp == PROP_MH_PAIR) {
        const int acc = a.accept[c];
        if (j == 3 && (acc & 1)) v = exp(a.theta[c]);
        if (is_ch && (acc & 2)) {
            PropScalars x{};
            x.hyper_old = x.hyper_new = a.par.at(c, npar - 1);
            ChProp<PROP_HYPER_NOOP> chs{a.par, a.normals, c, 4, x};
            v = chs.ch(j - 4);
        }
    } else if (kProp == PROP_GAMMA_RW) {
        if (j == 3) v = exp(a.theta[c]);
    } else if (kProp == PROP_MH_SCALAR) {
        if (j == a.mh_index) v = a.mh_is_log ? exp(a.theta[c]) : a.theta[c];
    } else if (kProp == PROP_MH_HYPER) {
        if (j == npar - 1) v = a.theta[c];
        if (is_ch) {
            PropScalars x{};
            x.hyper_old = a.par.at(c, npar - 1);
            x.hyper_new = a.theta[c];
            ChProp<PROP_MH_HYPER> chs{a.par, a.normals, c, 4, x};
            v = chs.ch(j - 4);
        }
    } else if (kProp == PROP_MH_TREE) {
        const int acc = a.accept[c];
        for (int e = 0; e < a.tm; ++e) {
            if (!((acc >> e) & 1)) continue;
            if (a.tkind[e] == 0) {
                if (j == a.tidx[e]) v = a.jnew[c * 4 + e];
            } else {
                if (j == npar - 1) v = a.jnew[c * 4 + e];
                if (is_ch) {
                    PropScalars x{};
                    x.hyper_old = a.par.at(c, npar - 1);
                    x.hyper_new = a.jnew[c * 4 + e];
                    x.rot_ch = 1;
                    ChProp<PROP_MH_TREE> chs{a.par, a.normals, c, 4, x};
                    v = chs.ch(j - 4);
                }
            }
        }
    } else if (kProp == PROP_MH_JOINT) {
        bool has_hyper = false;
        double hyper_new = a.par.at(c, npar - 1);
        for (int i = 0; i < a.jd; ++i) {
            const int jj = a.jidx[i] == 4 ? npar - 1 : a.jidx[i];
            if (jj == j) v = a.jnew[c * 4 + i];
            if (a.jidx[i] == 4) {
                has_hyper = true;
                hyper_new = a.jnew[c * 4 + i];
            }
        }
        if (is_ch && has_hyper) {
            PropScalars x{};
            x.hyper_old = a.par.at(c, npar - 1);
            x.hyper_new = hyper_new;
            x.rot_ch = 1;
            ChProp<PROP_MH_JOINT> chs{a.par, a.normals, c, 4, x};
            v = chs.ch(j - 4);
        }
    } else if (kProp == PROP_HYPER_NOOP) {
        if (is_ch) {
            PropScalars x{};
            x.hyper_old = x.hyper_new = a.par.at(c, npar - 1);
            ChProp<PROP_HYPER_NOOP> chs{a.par, a.normals, c, 4, x};
            v = chs.ch(j - 4);
        }
    } else if (kProp == PROP_ESS_PAR) {
        if (j >= 1) {
            const EssParPre z = ess_par_pre(a, c, npar);
            PropScalars x;
            ess_par_scalars(a, c, npar, z, a.theta[c], x);
            if (is_ch) {
                ChProp<PROP_ESS_PAR> chs{a.par, a.normals, c, 4, x};
                v = chs.ch(j - 4);
            } else {
                v = j == 1 ? x.I0 : j == 2 ? x.R0 : j == 3 ? x.gam : x.hyper_new;
            }
        }
    } else if (kProp == PROP_ESS_CHPT) {
        if (is_ch) {
            PropScalars x{};
            double sn, ct;
            sincos(a.theta[c], &sn, &ct);
            x.ct = ct;
            x.sn = sn;
            x.hyper_old = x.hyper_new = a.par.at(c, npar - 1);
            x.rot_ch = 1;
            ChProp<PROP_ESS_CHPT> chs{a.par, a.normals, c, 4, x};
            v = chs.ch(j - 4);
        }
    }
    par_out.put(c, j, v);
}

// Param_Slice_update / Param_Slice_update_all2 as stand-alone operators (one thread per item).
__global__ void k_param_slice_update(DevProb P, long long B, View param, const double *theta, View newChs,
                                     double *param_new) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double sn, ct;
    sincos(theta[b], &sn, &ct);
    for (int j = 0; j < P.npt; ++j) param_new[b * P.npt + j] = param.at(b, j);
    for (int j = 0; j < P.nch; ++j)
        param_new[b * P.npt + P.nparam + j] = exp(rot(log(param.at(b, P.nparam + j)), ct, newChs.at(b, j), sn));
}

__global__ void k_param_slice_update_all2(DevProb P, long long B, CandArgs a, const double *theta, double *par_new) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    const int npar = 2 + P.npt;
    const EssParPre z = ess_par_pre(a, c, npar);
    PropScalars x;
    ess_par_scalars(a, c, npar, z, theta[c], x);
    ChProp<PROP_ESS_PAR> chs{a.par, a.normals, c, 4, x};
    double *o = par_new + c * npar;
    o[0] = x.S0;
    o[1] = x.I0;
    o[2] = x.R0;
    o[3] = x.gam;
    for (int j = 0; j < P.nch; ++j) o[4 + j] = chs.ch(j);
    o[npar - 1] = x.hyper_new;
}

// ================================================================================================
// Counter-based random streams (Philox4x32-10): chain c, iteration it, slot q -> one 128-bit block.
// Replaces the reference's calls into R's global RNG (arma::randn / R::runif, unpinned third-party) for
// resident chains; the streams are replayable on the host (lna_draw_streams) so that every chain can be
// re-run through the CPU restatement draw for draw.
// ================================================================================================
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in the OPEN interval (0,1), like unif_rand(): 52 random bits b -> (b + 1/2) 2^-52, every value exactly
// representable (min 2^-53, max 1 - 2^-53)
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    const unsigned long long b = (((unsigned long long)hi << 32) | lo) >> 12;  // 52 bits
    return ((double)b + 0.5) * (1.0 / 4503599627370496.0);
}
// one thread per (chain, pair of outputs): normals via Box-Muller, uniforms directly.
// Philox key = the 64-bit seed alone; counter = (iteration[31:0], iteration[47:32] << 16 | q, chain[31:0], chain[63:32]):
// seed, chain, iteration and slot occupy disjoint bits, so no two (seed, chain, iteration, q) share a block
// (the host checks q < 2^16 and iteration < 2^48).
__global__ void k_draw_streams(long long B, long long chain_offset, int n_norm, int n_unif, unsigned long long seed,
                               unsigned long long iteration, double *normals, double *uniforms) {
    const int npair_n = (n_norm + 1) / 2, npair_u = (n_unif + 1) / 2, per = npair_n + npair_u;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * per) return;
    const long long c = idx / per;
    const int q = (int)(idx % per);
    const unsigned long long chain = (unsigned long long)(chain_offset + c);
    uint32_t r[4];
    philox4x32_10((uint32_t)iteration, ((uint32_t)(iteration >> 32) << 16) | (uint32_t)q, (uint32_t)chain,
                  (uint32_t)(chain >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double a = u01(r[0], r[1]), b = u01(r[2], r[3]);
    if (q < npair_n) {
        const double rad = sqrt(-2.0 * log(a));
        double sn, cs;
        sincospi(2.0 * b, &sn, &cs);
        normals[c * n_norm + 2 * q] = rad * cs;
        if (2 * q + 1 < n_norm) normals[c * n_norm + 2 * q + 1] = rad * sn;
    } else {
        const int j = q - npair_n;
        uniforms[c * n_unif + 2 * j] = a;
        if (2 * j + 1 < n_unif) uniforms[c * n_unif + 2 * j + 1] = b;
    }
}

// item-major <-> slot-major transposes for the resident chain state
__global__ void k_to_soa(long long B, int n, const double *src_item_major, double *dst_soa) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n) return;
    const long long c = idx % B;
    const int j = (int)(idx / B);
    dst_soa[idx] = src_item_major[c * n + j];
}
__global__ void k_from_soa(long long B, int n, const double *src_soa, double *dst_item_major) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n) return;
    const long long c = idx / n;
    const int j = (int)(idx % n);
    dst_item_major[idx] = src_soa[(long long)j * B + c];
}

}  // namespace lna
