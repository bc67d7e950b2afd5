// lna_device.cuh -- device-side building blocks of the LNA + coalescent likelihood path (sm_100a).
//
// Design: one THREAD per trajectory, everything that evolves in time lives in registers
// (mean state, fundamental matrix F0, symmetric covariance Sigma, latent deviation X, likelihood
// accumulator).  The per-genealogy tables (change-point step indices, per-segment dt, per-cell
// coalescent sums) are staged in shared memory once per block and read with warp-uniform indices
// (broadcast).  FP64 throughout; nothing here is a dense contraction, so no tensor cores.
//
// Reference semantics (file:line into MingweiWilliamTang/LNAphyloDyn):
//   rhs / Jacobian / hazards   src/LNA_functional.cpp:107-127,156-198,248-344
//   "RK4" (k4 at dt/2)         src/LNA_functional.cpp:481-489
//   Euler F0 / Sigma           src/LNA_functional.cpp:573-589
//   chol(Sigma+1e-8 I)^T       src/LNA_functional.cpp:655-665
//   TransformTraj              src/LNA_functional.cpp:890-898
//   volz_loglik_nh2            src/LNA_functional.cpp:1119-1176
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifndef LNA_UNROLL
#define LNA_UNROLL 2  // fine steps unrolled in the time loop
#endif
#define LNA_DO_PRAGMA_(x) _Pragma(#x)
#define LNA_DO_PRAGMA(x) LNA_DO_PRAGMA_(x)

namespace lna {

constexpr double kSentinel = -10000000.0;
constexpr double kJitter = 0.00000001;
constexpr double kTwoPi = 2.0 * 3.14159265358979323846264338327950288;

enum : int { ST_CHOL_FAIL = 1, ST_NEG_STATE = 2, ST_NAN = 4, ST_CAP_HIT = 8 };

// Problem description passed by value to every kernel (pointers are device pointers).
struct DevProb {
    int model, p, nch, nparam, iR0, iGamma;
    int g, k, nt, nrow, npt;
    int n0;        // Init$ng - 1 : number of coalescent cells
    int Lrow;      // first coarse row with time >= t_correct (capped at nrow-1)
    int idxS, idxI;// state columns used by the Volz likelihood (x_i[2], x_i[3])
    int has_init;
    int transX;    // 0 standard, 1 log (per-export kernels only; the fused evaluation is standard-scale)
    double N, dt_ode, t_correct;
    // dt-derived constants computed on the host: as kernel parameters they are constant-bank operands,
    // which keeps the DFMAs that use them at full issue rate (a DFMA with three distinct REGISTER
    // sources issues at 2/3 rate on B200: tools/fp64_microbench.cu)
    double h2, dt6, dt3, invN;
    const int *chidx;      // [nch] first fine index s with times[s] >= cht_j (monotone)
    const double *dt_seg;  // [k]   times[i*g+1]-times[i*g]   (SigmaF's per-block dt, A.3)
    const double *times;   // [nt]
    const double *cht;     // [nch]
    const double *cellCD;  // [n0]  sum_{e in cell} C_e*D_e
    const double *cellY;   // [n0]  sum_{e in cell} y_e
    const int *cellCnt;    // [n0]  gridrep
    // rowUB[r] = upper bound of the coalescent terms of all trajectory rows > r, whatever the trajectory: cell i contributes
    // at most y_i (log(y_i / (2 a_i)) - 1) (the maximum of -2 a lambda + y log lambda over lambda > 0); lna_async.cuh
    const double *rowUB;   // [nrow]
    int async_ok;          // change points on LNA-interval boundaries (what k_ess_par_async needs)
    // raw events (for the per-event mirror of volz_loglik_nh2 / coal_loglik)
    int E;
    const double *evC, *evD, *evY;
    const int *evStart;    // [n0+1] prefix sums of gridrep
};

// Shared-memory copy of the small tables.
constexpr int kSegConst = 64;  // per-segment dt lives in the parameter block up to this many LNA intervals
struct SegDt {
    double v[kSegConst];
};

struct Tables {
    const int *chidx;
    const double *dt_seg;
    const double *cellCD;
    const double *cellY;
    const int *cellCnt;
    const double *ctimes;  // [nrow] times[r*g]: column 0 of every coarse trajectory
    double *eps_slot;      // this thread's staging slot for the latent increments of one interval (kEpsSlot doubles)
};
constexpr int kEpsSlot = 4;  // doubles per thread (p <= 3, padded)

__host__ __device__ inline size_t tables_bytes(int nch, int k, int n0) {
    size_t b = 0;
    b += sizeof(double) * (size_t)k;
    b += sizeof(double) * (size_t)(k + 2);  // coarse times (nrow <= k + 2)
    b += sizeof(double) * (size_t)n0 * 2;
    b += sizeof(int) * (size_t)nch;
    b += sizeof(int) * (size_t)n0;
    return (b + 15) & ~(size_t)15;
}
// dynamic shared memory of a kernel that runs FULL evaluations: the tables + one staging slot per thread
__host__ __device__ inline size_t full_eval_smem_bytes(size_t tab_bytes, int block) {
    return tab_bytes + sizeof(double) * (size_t)kEpsSlot * (size_t)block;
}

// `with_eps_slots`: the kernel was launched with full_eval_smem_bytes() and wants per-thread staging slots
__device__ inline Tables stage_tables(const DevProb &P, unsigned char *smem, bool with_eps_slots = false) {
    double *dts = reinterpret_cast<double *>(smem);
    double *ct = dts + P.k;
    double *cd = ct + (P.k + 2);
    double *yy = cd + P.n0;
    int *ci = reinterpret_cast<int *>(yy + P.n0);
    int *cc = ci + P.nch;
    double *slots = reinterpret_cast<double *>(smem + tables_bytes(P.nch, P.k, P.n0));
    for (int i = threadIdx.x; i < P.k; i += blockDim.x) dts[i] = P.dt_seg[i];
    for (int i = threadIdx.x; i < P.nrow && i < P.k + 2; i += blockDim.x) ct[i] = i * P.g < P.nt ? P.times[i * P.g] : 0.0;
    for (int i = threadIdx.x; i < P.nch; i += blockDim.x) ci[i] = P.chidx[i];
    if (P.has_init)
        for (int i = threadIdx.x; i < P.n0; i += blockDim.x) {
            cd[i] = P.cellCD[i];
            yy[i] = P.cellY[i];
            cc[i] = P.cellCnt[i];
        }
    __syncthreads();
    return Tables{ci, dts, cd, yy, cc, ct, with_eps_slots ? slots + (size_t)threadIdx.x * kEpsSlot : nullptr};
}

// 8-byte asynchronous global -> shared copy (LDGSTS): no destination register, so no scoreboard is held across the
// time loop; completion is awaited where the value is consumed.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Strided views: element j of item b at base[b*sb + j*sj].  (sb = n, sj = 1) is the C-ABI's
// item-major layout; (sb = 1, sj = B) is the coalesced struct-of-arrays layout of resident chains.
struct View {
    const double *base;
    long long sb, sj;
    __device__ __forceinline__ double at(long long b, int j) const { return __ldg(base + b * sb + (long long)j * sj); }
    __device__ __forceinline__ bool ok() const { return base != nullptr; }
};
struct OutView {
    double *base;
    long long sb, sj;
    __device__ __forceinline__ void put(long long b, int j, double v) const { base[b * sb + (long long)j * sj] = v; }
    __device__ __forceinline__ bool ok() const { return base != nullptr; }
};

// Where the optional per-item outputs of a FULL evaluation go (any base may be null).
struct FullOut {
    OutView A, L, ode, betaN, latent;  // p*p*k, p*p*k, nrow*(p+1), nrow, nrow*(p+1)
};

// ------------------------------------------------------------------------------------------------
// Change-ratio sources.  The FULL evaluation pulls ch_j when the time loop crosses change point j,
// so a proposal never has to be materialised as a 40-odd element vector per thread.
// ------------------------------------------------------------------------------------------------
// A source exposes   Raw raw(j)   (the global loads, issued one change point ahead so their latency
// hides behind the previous run of fine steps) and   double eval(Raw)   (the arithmetic).
struct ChFromParam {  // ch_j = param[nparam + j]
    View param;
    long long b;
    int nparam;
    struct Raw {
        double c;
    };
    __device__ __forceinline__ Raw raw(int j) const { return Raw{param.at(b, nparam + j)}; }
    __device__ __forceinline__ double eval(const Raw &r) const { return r.c; }
    __device__ __forceinline__ double ch(int j) const { return eval(raw(j)); }
};

// Cursor over the change points of one trajectory: `next_s` is the fine-step index at which the next
// change ratio applies (INT_MAX when none is left), so the time loop only compares two registers.
template <class ChSrc>
struct ChCursor {
    const ChSrc &src;
    const int *chidx;
    int nch, nxt, next_s;
    typename ChSrc::Raw pre;
    __device__ __forceinline__ ChCursor(const ChSrc &s, const int *ci, int n) : src(s), chidx(ci), nch(n), nxt(0) {
        next_s = nch > 0 ? chidx[0] : 0x7fffffff;
        pre = src.raw(0);  // nch == 0 reads the (in-range) slot after the model parameters; never used
    }
    // apply every change with index <= s to the running R0; returns true if anything changed
    __device__ __forceinline__ bool advance(int s, double &R0cum) {
        if (next_s > s) return false;
        do {
            R0cum *= src.eval(pre);
            ++nxt;
            next_s = nxt < nch ? chidx[nxt] : 0x7fffffff;
            pre = src.raw(nxt < nch ? nxt : 0);
        } while (next_s <= s);
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// Coalescent accumulator (volz_loglik_nh2 with the per-event sums hoisted per cell).
// Cell i (0..n0-1) uses coarse row Lrow-i; rows 0..n0 are checked for negative states.
// ------------------------------------------------------------------------------------------------
struct VolzAcc {
    double acc = 0.0;
    int flags = 0;
    __device__ __forceinline__ void row(const DevProb &P, const Tables &T, int r, double S, double I, double beta,
                                        bool any_neg) {
        if (r <= P.n0 && any_neg) flags |= ST_NEG_STATE;
        const int i = P.Lrow - r;
        if (i >= 0 && i < P.n0 && T.cellCnt[i] > 0) {
            // reference per event: -2*C*D*beta*exp(-log I)*exp(log S) + y*(log beta - log I + log S)
            const double ratio = S / I;
            const double term = fma(-2.0 * T.cellCD[i] * beta, ratio, T.cellY[i] * log(beta * ratio));
            if (!(S > 0.0) || !(I > 0.0) || term != term) flags |= ST_NAN;
            acc += term;
        }
    }
    __device__ __forceinline__ double result(int &status) const {
        if (flags & ST_NEG_STATE) {
            status |= ST_NEG_STATE;
            return kSentinel;
        }
        if (flags & ST_NAN) {
            status |= ST_NAN;
            return kSentinel;
        }
        return acc;
    }
};

// ------------------------------------------------------------------------------------------------
// Partial evaluations: a FULL evaluation cut at LNA-interval boundaries into pieces that different warps run one after the
// other (k_full_loglik_bal: the pieces of one warp-load go to two SM sub-partitions so that all four finish together).
// Everything that crosses a cut is the trajectory's running state -- 6 doubles and 3 ints per trajectory; beta and the
// change-point cursor are rebuilt from (R0cum, nxt) with the expressions the uncut evaluation uses, so a cut changes no bit.
// ------------------------------------------------------------------------------------------------
struct PartState {
    double S, I, X0, X1, R0cum, acc;
    int nxt, flags, st, pad;
};
struct NoPart {  // the whole evaluation (every kernel but k_full_loglik_bal)
    static constexpr bool kOn = false;
};
struct PartRange {
    static constexpr bool kOn = true;
    int seg_begin, seg_end;  // intervals [seg_begin, seg_end) of 0..k
    PartState *in;           // this trajectory's state at seg_begin (seg_begin > 0)
    PartState *out;          // where it goes at seg_end (seg_end < k)
};

// ------------------------------------------------------------------------------------------------
// SIR (p = 2) FULL evaluation.
//   param = (R0, gamma, ch_1..ch_nch, hyper); theta0 = beta(t) = R0*prod(ch)*gamma/N, theta1 = param[1].
// Per fine step: one Euler step of F0 and Sigma at the left endpoint, then the reference's RK4
// variant.  F's second row is -(first row) - (0, gamma), which removes a third of the multiplies;
// the hazards h = (beta S I, gamma I) are shared between the covariance step and RK stage 1.
// ------------------------------------------------------------------------------------------------
template <bool kStore, class ChSrc, class EpsSrc, class Part = NoPart>
__device__ __forceinline__ void sir_full_eval(const DevProb &P, const SegDt &SD, const Tables &T, const ChSrc &chs,
                                              const EpsSrc &eps, double R0, double gam, double S, double I,
                                              const FullOut &out, long long b, double &loglik, int &status,
                                              const Part part = Part{}) {
    const double h2 = P.h2, dt6 = P.dt6;
    const double hg = h2 * gam, dg6 = dt6 * gam;
    const double invN = P.invN;
    const int g = P.g, k = P.k, nrow = P.nrow;
    const bool seg_const = k <= kSegConst;
    double R0cum = R0;
    ChCursor<ChSrc> cur(chs, T.chidx, P.nch);
    cur.advance(0, R0cum);
    double beta = (R0cum * gam) * invN, hb = h2 * beta;
    double X0 = 0.0, X1 = 0.0;
    VolzAcc va;
    int st = 0;
    int seg_begin = 0, seg_end = k;
    bool resumed = false;
    if constexpr (Part::kOn) {
        seg_begin = part.seg_begin;
        seg_end = part.seg_end;
        resumed = seg_begin > 0;
        if (resumed) {  // the state the previous piece left at this cut
            const PartState ps = *part.in;
            S = ps.S;
            I = ps.I;
            X0 = ps.X0;
            X1 = ps.X1;
            R0cum = ps.R0cum;
            va.acc = ps.acc;
            va.flags = ps.flags;
            st = ps.st;
            cur.nxt = ps.nxt;
            cur.next_s = cur.nxt < cur.nch ? cur.chidx[cur.nxt] : 0x7fffffff;
            cur.pre = chs.raw(cur.nxt < cur.nch ? cur.nxt : 0);
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
        }
    }

    // coarse row 0 = initial state
    if (!resumed) {
        if (kStore) {
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
    }

    for (int seg = seg_begin; seg < seg_end; ++seg) {
        double p00 = 1.0, p01 = 0.0, p10 = 0.0, p11 = 1.0;  // F0
        double s00 = 0.0, s01 = 0.0, s11 = 0.0;              // Sigma (symmetric)
        // latent increments of this interval: asynchronous copy into the thread's shared-memory slot now, read at
        // the end of the interval (a register load would make ptxas wait for it BEFORE the time loop: measured
        // ~1000 stalled cycles per interval, profiles/r1b_ncu_kernels.md)
        const bool staged = T.eps_slot != nullptr && eps.stage(seg, 2, T.eps_slot);
        double e0 = 0.0, e1 = 0.0;
        if (!staged) {
            e0 = eps.at(seg, 0);
            e1 = eps.at(seg, 1);
        }
        int s = seg * g;
        const int s_end = s + g;
        // The step body is a macro-like lambda over `dts`: with k <= kSegConst the per-segment dt is read
        // from the parameter block with a warp-uniform index (uniform-register operand), otherwise from
        // the shared-memory table (vector register).
        auto run = [&](const double dts, const int run_end) {
            const double c1 = fma(-dts, gam, 1.0);        // 1 - gam dts
            const double c2 = fma(-2.0 * dts, gam, 1.0);  // 1 - 2 gam dts
            LNA_DO_PRAGMA(unroll LNA_UNROLL)
            for (; s < run_end; ++s) {
                const double bI = beta * I, bS = beta * S;
                const double a1 = bS * I;   // h0 = beta*S*I
                const double gI = gam * I;  // h1
                // F0 <- F0 + (F F0) dts,  F = [[-bI, -bS], [bI, bS-gam]]:
                //   row 0 -= dts u,  row 1 = (1 - gam dts) row 1 + dts u,   u_c = bI F0[0][c] + bS F0[1][c]
                const double t0 = bS * p10, t1 = bS * p11;
                const double u0 = fma(bI, p00, t0);
                const double u1 = fma(bI, p01, t1);
                p00 = fma(-dts, u0, p00);
                p01 = fma(-dts, u1, p01);
                p10 = fma(dts, u0, c1 * p10);
                p11 = fma(dts, u1, c1 * p11);
                // Sigma <- Sigma + (Sigma F' + F Sigma + A' diag(h) A) dts  with w_c = bI S[0][c] + bS S[1][c]:
                //   s00 += dts (a1 - 2 w0);  s01 = (1 - gam dts) s01 + dts (w0 - w1 - a1)
                //   s11 = (1 - 2 gam dts) s11 + dts (2 w1 + a1 + gI)
                const double q0 = bS * s01, q1 = bS * s11;
                const double w0 = fma(bI, s00, q0);
                const double w1 = fma(bI, s01, q1);
                s00 = fma(dts, fma(-2.0, w0, a1), s00);
                s01 = fma(dts, (w0 - w1) - a1, c1 * s01);
                s11 = fma(dts, fma(2.0, w1, a1 + gI), c2 * s11);
                // RK4 variant (all stages at t[s], k4 at X0 + k3*dt/2).  Stage j needs only a_{j-1} on its
                // critical path:  beta*S_j = beta*S - (h2 beta) a_{j-1},  I_j = (I - h2 gam I_{j-1}) + h2 a_{j-1};
                //   S' = S - dt/6 QA,  I' = I + dt/6 QA - gam dt/6 (I + 2 I2 + 2 I3 + I4),  QA = a1 + 2 a2 + 2 a3 + a4
                const double I2 = fma(h2, a1, fma(-hg, I, I));
                const double a2 = fma(-hb, a1, bS) * I2;
                const double I3 = fma(h2, a2, fma(-hg, I2, I));
                const double a3 = fma(-hb, a2, bS) * I3;
                const double I4 = fma(h2, a3, fma(-hg, I3, I));
                const double a4 = fma(-hb, a3, bS) * I4;
#ifdef LNA_SHORT_CHAIN
                // latency variant: everything except the a4 terms is ready when a4 arrives
                const double QA3 = fma(2.0, a2 + a3, a1);
                const double PI = fma(2.0, I2 + I3, I) + I4;
                const double Sp = fma(-dt6, QA3, S);
                const double Ip = fma(-dg6, PI, fma(dt6, QA3, I));
                S = fma(-dt6, a4, Sp);
                I = fma(dt6, a4, Ip);
#else
                const double QA = fma(2.0, a2 + a3, a1) + a4;
                const double PI = fma(2.0, I2 + I3, I) + I4;
                S = fma(-dt6, QA, S);
                I = fma(-dg6, PI, fma(dt6, QA, I));
#endif
            }
        };
        while (s < s_end) {
            // runs of fine steps between change points carry no per-step bookkeeping at all
            if (cur.advance(s, R0cum)) {
                beta = (R0cum * gam) * invN;
                hb = h2 * beta;
            }
            const int run_end = min(s_end, cur.next_s);
            if (seg_const) run(SD.v[seg], run_end);
            else run(T.dt_seg[seg], run_end);
        }
        if (staged) {
            cp_async_wait_all();
            e0 = T.eps_slot[0];
            e1 = T.eps_slot[1];
        }
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
        // beta at the coarse time (betaTs on coarse times)
        if (cur.advance(r * g, R0cum)) {
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
        }
        const double lS = S + X0, lI = I + X1;
        if (kStore) {
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
    if constexpr (Part::kOn) {
        if (seg_end < k) {  // hand the running state to the piece that continues at this cut
            PartState ps;
            ps.S = S;
            ps.I = I;
            ps.X0 = X0;
            ps.X1 = X1;
            ps.R0cum = R0cum;
            ps.acc = va.acc;
            ps.nxt = cur.nxt;
            ps.flags = va.flags;
            ps.st = st;
            ps.pad = 0;
            *part.out = ps;
            return;
        }
    }
    if (st & ST_CHOL_FAIL) {
        status |= ST_CHOL_FAIL;
        loglik = kSentinel;
    } else {
        loglik = va.result(status);
    }
}

// ------------------------------------------------------------------------------------------------
// SEIR (p = 3) FULL evaluation: states (S, E, I), param = (R0, mu, gamma, ch.., hyper),
// theta = (beta(t), mu, gamma); F = [[-bI,0,-bS],[bI,-mu,bS],[0,mu,-gam]];
// A' diag(h) A = [[h0,-h0,0],[-h0,h0+h1,-h1],[0,-h1,h1+h2]], h = (bSI, mu E, gam I).
// ------------------------------------------------------------------------------------------------
template <bool kStore, class ChSrc, class EpsSrc>
__device__ __forceinline__ void seir_full_eval(const DevProb &P, const SegDt &SD, const Tables &T, const ChSrc &chs,
                                               const EpsSrc &eps, double R0, double mu, double gam, double S,
                                               double Ex, double I, const FullOut &out, long long b,
                                               double &loglik, int &status) {
    const double h2 = P.h2, dt6 = P.dt6;
    const int g = P.g, k = P.k, nrow = P.nrow;
    const double invN = P.invN;
    const bool seg_const = k <= kSegConst;
    double R0cum = R0;
    ChCursor<ChSrc> cur(chs, T.chidx, P.nch);
    cur.advance(0, R0cum);
    double beta = (R0cum * gam) * invN;
    double X[3] = {0.0, 0.0, 0.0};
    VolzAcc va;
    int st = 0;
    auto put_row = [&](const OutView &o, int r, double a, double c, double d) {
        o.put(b, r, T.ctimes[r]);
        o.put(b, nrow + r, a);
        o.put(b, 2 * nrow + r, c);
        o.put(b, 3 * nrow + r, d);
    };
    auto volz_row = [&](int r, double a, double c, double d) {
        // (selects, not an indexed local array: no stack frame)
        const double vS = P.idxS == 0 ? a : (P.idxS == 1 ? c : d), vI = P.idxI == 0 ? a : (P.idxI == 1 ? c : d);
        va.row(P, T, r, vS, vI, beta, (a < 0.0) | (c < 0.0) | (d < 0.0));
    };
    if (kStore) {
        if (out.ode.ok()) put_row(out.ode, 0, S, Ex, I);
        if (out.latent.ok()) put_row(out.latent, 0, S, Ex, I);
        if (out.betaN.ok()) out.betaN.put(b, 0, beta);
    }
    if (P.has_init) volz_row(0, S, Ex, I);

    for (int seg = 0; seg < k; ++seg) {
        // F0 row-major f[r][c]; Sigma symmetric: s00 s01 s02 s11 s12 s22
        double f00 = 1, f01 = 0, f02 = 0, f10 = 0, f11 = 1, f12 = 0, f20 = 0, f21 = 0, f22 = 1;
        double s00 = 0, s01 = 0, s02 = 0, s11 = 0, s12 = 0, s22 = 0;
        // latent increments of this interval: asynchronous copy into the thread's shared-memory slot now, read at the end
        // of the interval (like sir_full_eval: a register load would be waited for BEFORE the time loop)
        const bool staged = T.eps_slot != nullptr && eps.stage(seg, 3, T.eps_slot);
        double e0 = 0.0, e1 = 0.0, e2 = 0.0;
        if (!staged) {
            e0 = eps.at(seg, 0);
            e1 = eps.at(seg, 1);
            e2 = eps.at(seg, 2);
        }
        int s = seg * g;
        const int s_end = s + g;
        // per-segment dt from the parameter block (uniform-register operand) when k <= kSegConst, else shared memory
        auto run = [&](const double dts, const int run_end) {
          LNA_DO_PRAGMA(unroll LNA_UNROLL)
          for (; s < run_end; ++s) {
            const double bI = beta * I, bS = beta * S;
            const double h0 = bS * I, h1 = mu * Ex, hh2 = gam * I;
            // F F0: row0 = -(bI*F0[0] + bS*F0[2]); row1 = -row0 - mu*F0[1]; row2 = mu*F0[1] - gam*F0[2]
            const double u0 = fma(bI, f00, bS * f20), u1 = fma(bI, f01, bS * f21), u2 = fma(bI, f02, bS * f22);
            const double m0 = mu * f10, m1 = mu * f11, m2 = mu * f12;
            const double q0 = fma(-gam, f20, m0), q1 = fma(-gam, f21, m1), q2 = fma(-gam, f22, m2);
            f00 = fma(-dts, u0, f00); f01 = fma(-dts, u1, f01); f02 = fma(-dts, u2, f02);
            f10 = fma(dts, u0 - m0, f10); f11 = fma(dts, u1 - m1, f11); f12 = fma(dts, u2 - m2, f12);
            f20 = fma(dts, q0, f20); f21 = fma(dts, q1, f21); f22 = fma(dts, q2, f22);
            // M = F Sigma (rows): M0c = -(bI*S0c + bS*S2c); M1c = -M0c... = w_c - mu*S1c; M2c = mu*S1c - gam*S2c
            const double w0 = fma(bI, s00, bS * s02), w1 = fma(bI, s01, bS * s12), w2 = fma(bI, s02, bS * s22);
            const double n0 = mu * s01, n1 = mu * s11, n2 = mu * s12;
            const double M10 = w0 - n0, M11 = w1 - n1, M12 = w2 - n2;
            const double M20 = fma(-gam, s02, n0), M21 = fma(-gam, s12, n1), M22 = fma(-gam, s22, n2);
            const double d00 = fma(-2.0, w0, h0);
            const double d01 = (M10 - w1) - h0;
            const double d02 = M20 - w2;
            const double d11 = fma(2.0, M11, h0 + h1);
            const double d12 = (M12 + M21) - h1;
            const double d22 = fma(2.0, M22, h1 + hh2);
            s00 = fma(dts, d00, s00); s01 = fma(dts, d01, s01); s02 = fma(dts, d02, s02);
            s11 = fma(dts, d11, s11); s12 = fma(dts, d12, s12); s22 = fma(dts, d22, s22);
            // RK4 variant
            const double k1S = -h0, k1E = h0 - h1, k1I = h1 - hh2;
            const double S2 = fma(h2, k1S, S), E2 = fma(h2, k1E, Ex), I2 = fma(h2, k1I, I);
            const double a2 = (beta * S2) * I2, b2 = mu * E2;
            const double k2E = a2 - b2, k2I = fma(-gam, I2, b2);
            const double S3 = fma(-h2, a2, S), E3 = fma(h2, k2E, Ex), I3 = fma(h2, k2I, I);
            const double a3 = (beta * S3) * I3, b3 = mu * E3;
            const double k3E = a3 - b3, k3I = fma(-gam, I3, b3);
            const double S4 = fma(-h2, a3, S), E4 = fma(h2, k3E, Ex), I4 = fma(h2, k3I, I);
            const double a4 = (beta * S4) * I4, b4 = mu * E4;
            const double k4E = a4 - b4, k4I = fma(-gam, I4, b4);
            S = fma(-dt6, fma(2.0, a2 + a3, h0 + a4), S);
            Ex = fma(dt6, fma(2.0, k2E + k3E, k1E + k4E), Ex);
            I = fma(dt6, fma(2.0, k2I + k3I, k1I + k4I), I);
          }
        };
        while (s < s_end) {
            if (cur.advance(s, R0cum)) beta = (R0cum * gam) * invN;
            const int run_end = min(s_end, cur.next_s);
            if (seg_const) run(SD.v[seg], run_end);
            else run(T.dt_seg[seg], run_end);
        }
        if (staged) {
            cp_async_wait_all();
            e0 = T.eps_slot[0];
            e1 = T.eps_slot[1];
            e2 = T.eps_slot[2];
        }
        // 3x3 Cholesky of Sigma + 1e-8 I (upper U as dpotrf, L = U')
        const double c00 = s00 + kJitter;
        const double u00 = sqrt(c00);
        const double u01 = s01 / u00, u02 = s02 / u00;
        const double c11 = (s11 + kJitter) - u01 * u01;
        const double u11 = sqrt(c11);
        const double u12 = (s12 - u01 * u02) / u11;
        const double c22 = ((s22 + kJitter) - u02 * u02) - u12 * u12;
        const double u22 = sqrt(c22);
        if (!(c00 > 0.0) || !(c11 > 0.0) || !(c22 > 0.0)) st |= ST_CHOL_FAIL;
        const double nX0 = fma(f00, X[0], fma(f01, X[1], f02 * X[2])) + u00 * e0;
        const double nX1 = fma(f10, X[0], fma(f11, X[1], f12 * X[2])) + fma(u01, e0, u11 * e1);
        const double nX2 = fma(f20, X[0], fma(f21, X[1], f22 * X[2])) + fma(u02, e0, fma(u12, e1, u22 * e2));
        X[0] = nX0; X[1] = nX1; X[2] = nX2;
        const int r = seg + 1;
        if (cur.advance(r * g, R0cum)) beta = (R0cum * gam) * invN;
        const double lS = S + X[0], lE = Ex + X[1], lI = I + X[2];
        if (kStore) {
            if (out.A.ok()) {
                const double Am[9] = {f00, f10, f20, f01, f11, f21, f02, f12, f22};
#pragma unroll
                for (int e = 0; e < 9; ++e) out.A.put(b, seg * 9 + e, Am[e]);
            }
            if (out.L.ok()) {
                const double Lm[9] = {u00, u01, u02, 0.0, u11, u12, 0.0, 0.0, u22};
#pragma unroll
                for (int e = 0; e < 9; ++e) out.L.put(b, seg * 9 + e, Lm[e]);
            }
            if (out.ode.ok()) put_row(out.ode, r, S, Ex, I);
            if (out.latent.ok()) put_row(out.latent, r, lS, lE, lI);
            if (out.betaN.ok()) out.betaN.put(b, r, beta);
        }
        if (P.has_init) volz_row(r, lS, lE, lI);
    }
    if (st & ST_CHOL_FAIL) {
        status |= ST_CHOL_FAIL;
        loglik = kSentinel;
    } else {
        loglik = va.result(status);
    }
}

// Item-major epsilon (OriginTraj k x p column-major inside the item).
struct EpsFromView {
    View v;
    long long b;
    int k;
    __device__ __forceinline__ double at(int seg, int c) const { return v.at(b, c * k + seg); }
    // start the asynchronous copy of components 0..p-1 of interval `seg` into `slot`
    __device__ __forceinline__ bool stage(int seg, int p, double *slot) const {
        for (int c = 0; c < p; ++c) cp_async8(slot + c, v.base + b * v.sb + (long long)(c * k + seg) * v.sj);
        return true;
    }
};
// OriginTraj scaled by one factor: the self-mixing branch of ESlice_par_General (ESS_vec[6] = 1, LNA_functional.cpp:1660-1662),
// where the candidate's increments are OriginTraj * cos(theta) + (running OriginTraj_new) * sin(theta) = c * OriginTraj
struct EpsScaled {
    EpsFromView e;
    double c;
    __device__ __forceinline__ double at(int seg, int comp) const { return e.at(seg, comp) * c; }
    __device__ __forceinline__ bool stage(int, int, double *) const { return false; }
};
struct EpsZero {
    __device__ __forceinline__ double at(int, int) const { return 0.0; }
    __device__ __forceinline__ bool stage(int, int, double *) const { return false; }
};

}  // namespace lna
