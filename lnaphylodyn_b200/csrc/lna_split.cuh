// lna_split.cuh -- one FULL evaluation spread over THREE warps on three SM sub-partitions (sm_100a).
//
// Why (profiles/r1d_latency_variants_negative.md): the Metropolis and slice stages of the resident chains launch at most
// about one warp per SM sub-partition, so each launch lasts as long as ONE evaluation takes ONE warp: 2000 fine steps x 47
// FP64 instructions + 40 interval epilogues + 39 change-ratio exp/div/log chains.  A sub-partition's FP64 pipe has 16 lanes
// -- every warp-wide FP64 instruction occupies it for 2 cycles whatever the dependencies -- so inside one warp that floor is
// the warp's own pipe time, while the other FP64 pipes of the SM idle.  Here the evaluation is cut along the reference's own
// seams into a pipeline of three warps that hold the same 32 trajectories (lane = trajectory):
//
//   P  ODE_rk45 (src/LNA_functional.cpp:455-491): the mean trajectory, 22 FP64 instructions per fine step on the 9-deep RK
//      chain; writes (S, I) at every left endpoint into a shared-memory ring, and the state at each coarse time into a second one;
//   M  SigmaF / KF_param (:536-634): the Euler recurrences of F0 and Sigma along that trajectory (27 FP64 instructions per
//      step, shallow chains); hands (F0, Sigma, S, I) of each finished LNA interval on through a second ring;
//   E  first the change ratios of the proposal (param_transform's exp/div/log chains, :63-86, written to shared memory for
//      all three warps), then per interval KF_param_chol's Cholesky (:655-665), TransformTraj (:890-898), the
//      volz_loglik_nh2 term (:1119-1176) and the candidate's outputs; finally hands the likelihood back to P and M so that
//      the three warps take the same decisions afterwards.
//
// Counters in shared memory order the hand-overs (single producer / single consumer each; a warp polls with nanosleep).
// Every expression is the one of sir_full_eval (lna_device.cuh), in the same order: same bits.
#pragma once
#include "lna_device.cuh"

namespace lna {

constexpr int kSplitPiece = 25;  // fine steps per piece of the P -> M ring
constexpr int kSplitBufs = 3;    // pieces in the P -> M ring (slack for M's per-interval hand-over)
constexpr int kSplitIvals = 4;   // intervals in the M -> E ring and in the P -> M ring of coarse-time states
constexpr int kSplitIvalVals = 9;  // F0 (4), Sigma (3), S, I at the coarse time
constexpr int kSplitMaxCh = 48;  // change points held in shared memory (more: the host keeps the one-warp kernel)
constexpr int kSplitWarps = 3;

__host__ __device__ inline size_t split_group_bytes() {
    return sizeof(double) * (size_t)(kSplitBufs * kSplitPiece * 64 + kSplitIvals * kSplitIvalVals * 32 + kSplitIvals * 64 +
                                     kSplitMaxCh * 32 + 32) +
           sizeof(int) * (32 + 8);
}

struct SplitCtx {
    double *ring_pm;   // [kSplitBufs][kSplitPiece][2][32]
    double *ring_me;   // [kSplitIvals][kSplitIvalVals][32]
    double *coarse;    // [kSplitIvals][2][32]  mean state at the end of each LNA interval (P -> M)
    double *ratios;    // [kSplitMaxCh][32]
    double *res_ll;    // [32]
    int *res_st;       // [32]
    volatile int *pm_produced, *pm_consumed, *me_produced, *me_consumed, *ratios_ready, *res_seq, *coarse_produced;
    int piece, ival, evals;  // this warp's running counts (the warps of a group count the same events)
    int lane;
};

__device__ __forceinline__ SplitCtx split_ctx(unsigned char *group_smem, int lane) {
    SplitCtx c;
    c.ring_pm = reinterpret_cast<double *>(group_smem);
    c.ring_me = c.ring_pm + kSplitBufs * kSplitPiece * 64;
    c.coarse = c.ring_me + kSplitIvals * kSplitIvalVals * 32;
    c.ratios = c.coarse + kSplitIvals * 64;
    c.res_ll = c.ratios + kSplitMaxCh * 32;
    c.res_st = reinterpret_cast<int *>(c.res_ll + 32);
    int *f = c.res_st + 32;
    c.pm_produced = f;
    c.pm_consumed = f + 1;
    c.me_produced = f + 2;
    c.me_consumed = f + 3;
    c.ratios_ready = f + 4;
    c.res_seq = f + 5;
    c.coarse_produced = f + 6;
    c.piece = 0;
    c.ival = 0;
    c.evals = 0;
    c.lane = lane;
    return c;
}
__device__ __forceinline__ void split_reset(const SplitCtx &c) {
    *c.pm_produced = 0;
    *c.pm_consumed = 0;
    *c.me_produced = 0;
    *c.me_consumed = 0;
    *c.ratios_ready = 0;
    *c.res_seq = 0;
    *c.coarse_produced = 0;
}

#ifndef LNA_SPLIT_SLEEP_NS
#define LNA_SPLIT_SLEEP_NS 0  // 0 = spin on the shared-memory word (one dependent LDS per ~30 cycles)
#endif
__device__ __forceinline__ void split_wait(volatile int *ctr, int need) {
    while (*ctr < need) {
        if (LNA_SPLIT_SLEEP_NS > 0) __nanosleep(LNA_SPLIT_SLEEP_NS);
    }
    __threadfence_block();
}
__device__ __forceinline__ void split_post(volatile int *ctr, int value, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) *ctr = value;
}

// Change-point cursor over the ratios warp E publishes (P and M wait for a ratio only if they get there first).
struct SplitChCursor {
    const SplitCtx &c;
    const int *chidx;
    int nch, nxt, next_s, base;
    __device__ __forceinline__ SplitChCursor(const SplitCtx &cx, const int *ci, int n)
        : c(cx), chidx(ci), nch(n), nxt(0), base(cx.evals * n) {
        next_s = nch > 0 ? chidx[0] : 0x7fffffff;
    }
    __device__ __forceinline__ bool advance(int s, double &R0cum) {
        if (next_s > s) return false;
        do {
            split_wait(c.ratios_ready, base + nxt + 1);
            R0cum *= c.ratios[nxt * 32 + c.lane];
            ++nxt;
            next_s = nxt < nch ? chidx[nxt] : 0x7fffffff;
        } while (next_s <= s);
        return true;
    }
};

// hand-back of the result to the warps that did not compute it
__device__ __forceinline__ void split_take_result(SplitCtx &c, double &loglik, int &status) {
    split_wait(c.res_seq, c.evals + 1);
    loglik = c.res_ll[c.lane];
    status |= c.res_st[c.lane];
    c.evals += 1;
}

// ---- P: the mean ODE -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sir_split_P(SplitCtx &c, const DevProb &P, const Tables &T, double R0, double gam, double S,
                                            double I, double &loglik, int &status) {
    const double h2 = P.h2, dt6 = P.dt6;
    const double hg = h2 * gam, dg6 = dt6 * gam;
    const double invN = P.invN;
    const int g = P.g, k = P.k;
    double R0cum = R0;
    SplitChCursor cur(c, T.chidx, P.nch);
    cur.advance(0, R0cum);
    double beta = (R0cum * gam) * invN, hb = h2 * beta;
    for (int seg = 0; seg < k; ++seg) {
        int s = seg * g;
        const int s_end = s + g;
        while (s < s_end) {
            if (cur.advance(s, R0cum)) {
                beta = (R0cum * gam) * invN;
                hb = h2 * beta;
            }
            const int piece_end = min(min(s_end, cur.next_s), s + kSplitPiece);
            const int n = c.piece;
            split_wait(c.pm_consumed, n - (kSplitBufs - 1));  // the piece that last used this buffer has been read
            double *buf = c.ring_pm + (n % kSplitBufs) * (kSplitPiece * 64) + c.lane;
            LNA_DO_PRAGMA(unroll LNA_UNROLL)
            for (; s < piece_end; ++s) {
                buf[0] = S;
                buf[32] = I;
                buf += 64;
                const double bS = beta * S;
                const double a1 = bS * I;
                const double I2 = fma(h2, a1, fma(-hg, I, I));
                const double a2 = fma(-hb, a1, bS) * I2;
                const double I3 = fma(h2, a2, fma(-hg, I2, I));
                const double a3 = fma(-hb, a2, bS) * I3;
                const double I4 = fma(h2, a3, fma(-hg, I3, I));
                const double a4 = fma(-hb, a3, bS) * I4;
                const double QA = fma(2.0, a2 + a3, a1) + a4;
                const double PI = fma(2.0, I2 + I3, I) + I4;
                S = fma(-dt6, QA, S);
                I = fma(-dg6, PI, fma(dt6, QA, I));
            }
            split_post(c.pm_produced, n + 1, c.lane);
            c.piece = n + 1;
        }
        // the mean state at the coarse time, posted at once (M closes the interval with it and must not wait for the
        // next piece; measured: that wait made M the bottleneck of the pipeline)
        {
            const int m = c.ival;
            split_wait(c.me_produced, m - (kSplitIvals - 1));  // M has long finished with the interval that used this slot
            double *q = c.coarse + (m % kSplitIvals) * 64 + c.lane;
            q[0] = S;
            q[32] = I;
            split_post(c.coarse_produced, m + 1, c.lane);
            c.ival = m + 1;
        }
    }
    split_take_result(c, loglik, status);
}

// ---- M: LNA moments along the mean trajectory ---------------------------------------------------------------------------
__device__ __forceinline__ void sir_split_M(SplitCtx &c, const DevProb &P, const SegDt &SD, const Tables &T, double R0,
                                            double gam, double &loglik, int &status) {
    const double invN = P.invN;
    const int g = P.g, k = P.k;
    const bool seg_const = k <= kSegConst;
    double R0cum = R0;
    SplitChCursor cur(c, T.chidx, P.nch);
    cur.advance(0, R0cum);
    double beta = (R0cum * gam) * invN;
    for (int seg = 0; seg < k; ++seg) {
        double p00 = 1.0, p01 = 0.0, p10 = 0.0, p11 = 1.0;  // F0
        double s00 = 0.0, s01 = 0.0, s11 = 0.0;              // Sigma (symmetric)
        int s = seg * g;
        const int s_end = s + g;
        const double dts = seg_const ? SD.v[seg] : T.dt_seg[seg];
        const double c1 = fma(-dts, gam, 1.0);        // 1 - gam dts
        const double c2 = fma(-2.0 * dts, gam, 1.0);  // 1 - 2 gam dts
        while (s < s_end) {
            if (cur.advance(s, R0cum)) beta = (R0cum * gam) * invN;
            const int piece_end = min(min(s_end, cur.next_s), s + kSplitPiece);
            const int n = c.piece;
            split_wait(c.pm_produced, n + 1);
            const double *buf = c.ring_pm + (n % kSplitBufs) * (kSplitPiece * 64) + c.lane;
            LNA_DO_PRAGMA(unroll LNA_UNROLL)
            for (; s < piece_end; ++s) {
                const double Sx = buf[0], Ix = buf[32];
                buf += 64;
                const double bI = beta * Ix, bS = beta * Sx;
                const double a1 = bS * Ix;   // h0 = beta*S*I
                const double gI = gam * Ix;  // h1
                const double t0 = bS * p10, t1 = bS * p11;
                const double u0 = fma(bI, p00, t0);
                const double u1 = fma(bI, p01, t1);
                p00 = fma(-dts, u0, p00);
                p01 = fma(-dts, u1, p01);
                p10 = fma(dts, u0, c1 * p10);
                p11 = fma(dts, u1, c1 * p11);
                const double q0 = bS * s01, q1 = bS * s11;
                const double w0 = fma(bI, s00, q0);
                const double w1 = fma(bI, s01, q1);
                s00 = fma(dts, fma(-2.0, w0, a1), s00);
                s01 = fma(dts, (w0 - w1) - a1, c1 * s01);
                s11 = fma(dts, fma(2.0, w1, a1 + gI), c2 * s11);
            }
            split_post(c.pm_consumed, n + 1, c.lane);
            c.piece = n + 1;
        }
        // the mean state at the coarse time (P posts it when it finishes the interval)
        double Sc, Ic;
        {
            const int m = c.ival;
            split_wait(c.coarse_produced, m + 1);
            const double *q = c.coarse + (m % kSplitIvals) * 64 + c.lane;
            Sc = q[0];
            Ic = q[32];
        }
        // hand the interval on to E
        {
            const int m = c.ival;
            split_wait(c.me_consumed, m - (kSplitIvals - 1));
            double *q = c.ring_me + (m % kSplitIvals) * (kSplitIvalVals * 32) + c.lane;
            q[0] = p00;
            q[32] = p01;
            q[64] = p10;
            q[96] = p11;
            q[128] = s00;
            q[160] = s01;
            q[192] = s11;
            q[224] = Sc;
            q[256] = Ic;
            split_post(c.me_produced, m + 1, c.lane);
            c.ival = m + 1;
        }
    }
    split_take_result(c, loglik, status);
}

// ---- E: change ratios, interval epilogues, coalescent likelihood, outputs ----------------------------------------------------
template <bool kStore, class ChSrc, class EpsSrc>
__device__ __forceinline__ void sir_split_E(SplitCtx &c, const DevProb &P, const Tables &T, const ChSrc &chs,
                                            const EpsSrc &eps, double R0, double gam, double S, double I, const FullOut &out,
                                            long long b, bool do_store, double &loglik, int &status) {
    const double invN = P.invN;
    const int g = P.g, k = P.k, nrow = P.nrow, nch = P.nch;
    // the proposal's change ratios, for all three warps
    // (their global loads run three ratios ahead: one exposed load latency per ratio made this phase 44 k cycles, long enough
    // for the M -> E ring to fill and stall M and then P at the start of every evaluation)
    {
        const int base = c.evals * nch;
        typedef typename ChSrc::Raw Raw;
        Raw r0{}, r1{}, r2{};
        if (nch > 0) r0 = chs.raw(0);
        if (nch > 1) r1 = chs.raw(1);
        if (nch > 2) r2 = chs.raw(2);
        for (int j = 0; j < nch; ++j) {
            const Raw cur_raw = r0;
            r0 = r1;
            r1 = r2;
            if (j + 3 < nch) r2 = chs.raw(j + 3);
            c.ratios[j * 32 + c.lane] = chs.eval(cur_raw);
            split_post(c.ratios_ready, base + j + 1, c.lane);
        }
    }
    double R0cum = R0;
    SplitChCursor cur(c, T.chidx, nch);
    cur.advance(0, R0cum);
    double beta = (R0cum * gam) * invN;
    double X0 = 0.0, X1 = 0.0;
    VolzAcc va;
    int st = 0;
    if (kStore && do_store) {
        if (out.ode.ok()) {
            out.ode.put(b, 0, T.ctimes[0]);
            out.ode.put(b, nrow, S);
            out.ode.put(b, 2 * nrow, I);
        }
        if (out.latent.ok()) {
            out.latent.put(b, 0, T.ctimes[0]);
            out.latent.put(b, nrow, S);
            out.latent.put(b, 2 * nrow, I);
        }
        if (out.betaN.ok()) out.betaN.put(b, 0, beta);
    }
    if (P.has_init) va.row(P, T, 0, S, I, beta, (S < 0.0) | (I < 0.0));

    for (int seg = 0; seg < k; ++seg) {
        const double e0 = eps.at(seg, 0), e1 = eps.at(seg, 1);  // (in flight while this warp waits for the interval)
        const int m = c.ival;
        split_wait(c.me_produced, m + 1);
        const double *q = c.ring_me + (m % kSplitIvals) * (kSplitIvalVals * 32) + c.lane;
        const double p00 = q[0], p01 = q[32], p10 = q[64], p11 = q[96];
        const double s00 = q[128], s01 = q[160], s11 = q[192];
        S = q[224];
        I = q[256];
        split_post(c.me_consumed, m + 1, c.lane);
        c.ival = m + 1;
        // Cholesky of Sigma + 1e-8 I  (lower)
        const double c00 = s00 + kJitter;
        double l00 = sqrt(c00);
        double l10 = s01 / l00;
        const double c11 = (s11 + kJitter) - l10 * l10;
        double l11 = sqrt(c11);
        if (!(c00 > 0.0) || !(c11 > 0.0)) st |= ST_CHOL_FAIL;
        // X <- A X + L eps
        const double nX0 = fma(p00, X0, p01 * X1) + l00 * e0;
        const double nX1 = fma(p10, X0, p11 * X1) + fma(l10, e0, l11 * e1);
        X0 = nX0;
        X1 = nX1;
        const int r = seg + 1;
        if (cur.advance(r * g, R0cum)) beta = (R0cum * gam) * invN;  // beta at the coarse time (betaTs on coarse times)
        const double lS = S + X0, lI = I + X1;
        if (kStore && do_store) {
            if (out.A.ok()) {
                out.A.put(b, seg * 4 + 0, p00);
                out.A.put(b, seg * 4 + 1, p10);
                out.A.put(b, seg * 4 + 2, p01);
                out.A.put(b, seg * 4 + 3, p11);
            }
            if (out.L.ok()) {
                out.L.put(b, seg * 4 + 0, l00);
                out.L.put(b, seg * 4 + 1, l10);
                out.L.put(b, seg * 4 + 2, 0.0);
                out.L.put(b, seg * 4 + 3, l11);
            }
            if (out.ode.ok()) {
                out.ode.put(b, r, T.ctimes[r]);
                out.ode.put(b, nrow + r, S);
                out.ode.put(b, 2 * nrow + r, I);
            }
            if (out.latent.ok()) {
                out.latent.put(b, r, T.ctimes[r]);
                out.latent.put(b, nrow + r, lS);
                out.latent.put(b, 2 * nrow + r, lI);
            }
            if (out.betaN.ok()) out.betaN.put(b, r, beta);
        }
        if (P.has_init) va.row(P, T, r, lS, lI, beta, (lS < 0.0) | (lI < 0.0));
    }
    if (st & ST_CHOL_FAIL) {
        status |= ST_CHOL_FAIL;
        loglik = kSentinel;
    } else {
        loglik = va.result(status);
    }
    c.res_ll[c.lane] = loglik;
    c.res_st[c.lane] = status;
    split_post(c.res_seq, c.evals + 1, c.lane);
    c.evals += 1;
}

// ================================================================================================
// FULL evaluations of a SMALL batch (one Update_Param / New_Param_List call of a single chain, a handful of proposals):
// the same pipeline, one group of three warps per 32 items.  Item-major inputs like k_full_loglik.
// ================================================================================================
template <bool kStore>
__global__ void __launch_bounds__(96, 4)
k_full_loglik_split(const __grid_constant__ DevProb P, const __grid_constant__ SegDt SD, long long B, View param, View initial,
                    View origin, FullOut out, double *coalLog, int *status, int group) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 0 = P, 1 = M, 2 = E
    SplitCtx sx = split_ctx(smem + full_eval_smem_bytes(tables_bytes(P.nch, P.k, P.n0), blockDim.x), lane);
    if (role == 0 && lane == 0) split_reset(sx);
    const Tables T = stage_tables(P, smem, true);
    const long long b0 = (long long)blockIdx.x * 32 + lane;
    const bool on = b0 < B;
    const long long b = on ? b0 : 0;  // idle lanes evaluate item 0 and write nothing (the hand-overs are warp-wide)
    const long long bc = group > 1 ? b / group : b;  // chain-shared initial state / latent increments (k_full_loglik)
    const double R0 = param.at(b, P.iR0), gam = param.at(b, 1);
    const double S0 = initial.at(bc, 0), I0 = initial.at(bc, 1);
    double ll = kSentinel;
    int st = 0;
    if (role == 0) {
        sir_split_P(sx, P, T, R0, gam, S0, I0, ll, st);
    } else if (role == 1) {
        sir_split_M(sx, P, SD, T, R0, gam, ll, st);
    } else {
        ChFromParam chs{param, b, P.nparam};
        if (origin.ok()) {
            EpsFromView eps{origin, bc, P.k};
            sir_split_E<kStore>(sx, P, T, chs, eps, R0, gam, S0, I0, out, b, on, ll, st);
        } else {
            EpsZero eps;
            sir_split_E<kStore>(sx, P, T, chs, eps, R0, gam, S0, I0, out, b, on, ll, st);
        }
        if (on) {
            if (coalLog) coalLog[b] = ll;
            if (status) status[b] = st;
        }
    }
}

}  // namespace lna
