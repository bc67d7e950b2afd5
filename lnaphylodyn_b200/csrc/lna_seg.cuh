// lna_seg.cuh -- ESlice_par_General (src/LNA_functional.cpp:1532-1688) for LARGE resident batches: sequential candidates with
// provable early rejection, resumable evaluation, compaction between launches (sm_100a).
//
// With tens of thousands of chains per GPU the machine is full without any speculation, so the cheapest schedule is the
// reference's own: one candidate of a chain at a time, the next one only after the previous is rejected.  What the reference
// cannot do is STOP a candidate early: its coalescent log-likelihood is a sum over the genealogy's grid cells, the cells still
// to come are bounded above by a constant of the genealogy (rowUB, see lna_async.cuh), and once  partial + rowUB[r] < logy
// the candidate is rejected whatever follows -- on the vignette problems after 40 % of the trajectory on average, which
// turns the 10.6 evaluations of a slice update into 4.9 evaluation-equivalents.
// One lane per chain.  A launch advances every unresolved chain by at most T LNA intervals ("ticks"): a lane whose candidate
// is rejected starts the chain's next candidate in the following tick; after T ticks the lane parks its trajectory state (the
// six time-evolving doubles, the likelihood accumulator, the slice bracket) in HBM and the chain is appended to the next
// launch's dense list, so warps stay full while the batch thins out.  Lanes of a warp are in different intervals of different
// candidates, but every tick is the same g-step loop (condition, as for lna_async.cuh: change points on interval boundaries).
// Same decisions, same accepted candidate, same stored outputs as the sequential loop.
#pragma once
#include "lna_eslice.cuh"

namespace lna {

constexpr int kSegD = 12;  // doubles per chain: S, I, X0, X1, R0cum, acc, th, lo, hi, th_cur, (2 spare)
constexpr int kSegI = 8;   // ints per chain: cand, next, seg, nxt, flags, st, ts_j, (1 spare)
struct SegState {
    double *d;  // [kSegD][B]
    int *i;     // [kSegI][B]
};

__global__ void __launch_bounds__(128, 3)
k_ess_par_seg(const __grid_constant__ DevProb P, const __grid_constant__ SegDt SD, long long B, CandArgs a, SegState s,
              const int *list_in, const int *count_in, int *list_out, int *count_out, int T) {
    extern __shared__ __align__(16) unsigned char smem[];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_in = count_in ? (long long)__ldg(count_in) : B;
    if ((long long)blockIdx.x * blockDim.x >= n_in) return;  // (blocks beyond the list leave before touching shared memory)
    const Tables Tb = stage_tables(P, smem, true);
    const bool have = idx < n_in;
    const long long c = have ? (list_in ? (long long)__ldg(list_in + idx) : idx) : 0;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (!__any_sync(full, have)) return;  // (whole warps beyond the list leave at once)
    const int npar = 2 + P.npt, g = P.g, k = P.k, nrow = P.nrow;
    constexpr int kMaxCand = 20;
    const double h2 = P.h2, dt6 = P.dt6, invN = P.invN;
    // per-lane copy of what a candidate START needs: u[22] | S0 I0 R0 gamma hyper | nu_I0 nu_R0 nu_gamma nu_hyper
    constexpr int kG = kMaxUnif + 9;
    double *gs = reinterpret_cast<double *>(smem + full_eval_smem_bytes(tables_bytes(P.nch, P.k, P.n0), blockDim.x)) +
                 (size_t)threadIdx.x * kG;
    double *g_par = gs + kMaxUnif, *g_nrm = g_par + 5;
    if (have) {
        for (int q = 0; q < kMaxUnif; ++q) gs[q] = a.uniforms.at(c, q);
        for (int q = 0; q < 4; ++q) g_par[q] = a.par.at(c, q);
        g_par[4] = a.par.at(c, npar - 1);
        for (int q = 0; q < 3; ++q) g_nrm[q] = a.normals.at(c, q);
        g_nrm[3] = a.normals.at(c, npar - 2);
    }
    auto unif = [&](int q) { return gs[q]; };
    const EpsFromView eps{a.origin, c, P.k};
    EssParPre zpre{};
    double logy = 0.0, logy_lo = 0.0;
    if (have) {
        zpre.z0 = __ddiv_rn(__dsub_rn(log(g_par[1]), a.pr[0]), a.pr[1]);
        zpre.z1 = __ddiv_rn(__dsub_rn(log(g_par[2]), a.pr[2]), a.pr[3]);
        zpre.z2 = __ddiv_rn(__dsub_rn(log(g_par[3]), a.pr[4]), a.pr[5]);
        zpre.zh = __ddiv_rn(__dsub_rn(log(g_par[4]), a.pr[6]), a.pr[7]);
        logy = a.coal_log[c] + log(unif(0));
        logy_lo = logy - 1e-9 * (fabs(logy) + 1.0);
    }
    // ---- resume ------------------------------------------------------------------------------------------------------
    auto SD_ = [&](int f) -> double & { return s.d[(long long)f * B + c]; };
    auto SI_ = [&](int f) -> int & { return s.i[(long long)f * B + c]; };
    int cand = 0, next = 1, seg = 0, nxt = 0, flags = 0, st = 0;
    double S = 0, I = 0, X0 = 0, X1 = 0, R0cum = 0, acc = 0, th_cur = 0, beta = 0, hb = 0, gam = 0, hg = 0, dg6 = 0;
    ThetaSeq ts;
    ts.init();
    bool done = !have;
    bool revert_cur = false;
    int started = 0, ticks = 0;
    PropScalars x{};
    ChProp<PROP_ESS_PAR> chs{a.par, a.normals, c, 4, x};
    typename ChProp<PROP_ESS_PAR>::Raw pre{}, pre0{};
    if (have) pre0 = chs.raw(0);  // (ch_1 and its normal: the same for every candidate of the chain)

    auto setup_candidate = [&]() {  // proposal scalars of candidate `cand` at angle th_cur: ess_par_scalars, same expressions
        double sn, ct;
        sincos(th_cur, &sn, &ct);
        x.ct = ct;
        x.sn = sn;
        x.hyper_old = g_par[4];
        x.I0 = a.ess[0] ? exp(__dadd_rn(__dmul_rn(rot(zpre.z0, ct, g_nrm[0], sn), a.pr[1]), a.pr[0])) : g_par[1];
        x.R0 = a.ess[1] ? exp(__dadd_rn(__dmul_rn(rot(zpre.z1, ct, g_nrm[1], sn), a.pr[3]), a.pr[2])) : g_par[2];
        x.gam = a.ess[2] ? exp(__dadd_rn(__dmul_rn(rot(zpre.z2, ct, g_nrm[2], sn), a.pr[5]), a.pr[4])) : g_par[3];
        x.hyper_new = a.ess[4] ? exp(__dadd_rn(__dmul_rn(rot(zpre.zh, ct, g_nrm[3], sn), a.pr[7]), a.pr[6])) : g_par[4];
        x.rot_ch = a.ess[5];
        x.S0 = g_par[0];
        if (revert_cur) {  // the cap: the current parameters, bit for bit (:1609-1622)
            x.I0 = g_par[1];
            x.R0 = g_par[2];
            x.gam = g_par[3];
            x.rot_ch = 0;
        }
        x.noop = 0;
        chs.x = x;
        gam = x.gam;
        hg = h2 * gam;
        dg6 = dt6 * gam;
    };
    auto apply_changes = [&](int sidx) {
        bool any = false;
        while (nxt < P.nch && Tb.chidx[nxt] <= sidx) {
            R0cum *= chs.eval(pre);
            ++nxt;
            pre = chs.raw(nxt < P.nch ? nxt : 0);
            any = true;
        }
        if (any) {
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
        }
    };
    auto volz_row = [&](int r, double lS, double lI) {
        if (r <= P.n0 && ((lS < 0.0) | (lI < 0.0))) flags |= ST_NEG_STATE;
        const int i = P.Lrow - r;
        if (i >= 0 && i < P.n0 && Tb.cellCnt[i] > 0) {
            const double ratio = lS / lI;
            const double term = fma(-2.0 * Tb.cellCD[i] * beta, ratio, Tb.cellY[i] * log(beta * ratio));
            if (!(lS > 0.0) || !(lI > 0.0) || term != term) flags |= ST_NAN;
            acc += term;
        }
    };
    if (have && SI_(1) != 0) {  // the chain was parked by the previous launch
        cand = SI_(0);
        next = SI_(1);
        seg = SI_(2);
        nxt = SI_(3);
        flags = SI_(4);
        st = SI_(5);
        ts.j = SI_(6);
        S = SD_(0);
        I = SD_(1);
        X0 = SD_(2);
        X1 = SD_(3);
        R0cum = SD_(4);
        acc = SD_(5);
        ts.th = SD_(6);
        ts.lo = SD_(7);
        ts.hi = SD_(8);
        th_cur = SD_(9);
        if (cand != 0) {
            revert_cur = cand == kMaxCand + 1;
            setup_candidate();
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
            pre = chs.raw(nxt < P.nch ? nxt : 0);
        }
    }
    const long long slot = c;  // one scratch slot per chain: the accepted candidate is the last one written

    for (int tick = 0; tick < T; ++tick) {
        // ---- 1. a chain whose candidate was rejected starts its next one ---------------------------------------------
        if (!done && cand == 0) {
            started += 1;
            cand = next;
            next += 1;
            revert_cur = cand == kMaxCand + 1;
            if (!revert_cur) ts.advance(cand, unif);
            th_cur = revert_cur ? 0.0 : ts.th;
            setup_candidate();
            S = x.S0;
            I = x.I0;
            X0 = X1 = 0.0;
            R0cum = x.R0;
            nxt = 0;
            pre = pre0;
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
            apply_changes(0);
            acc = 0.0;
            flags = 0;
            st = 0;
            seg = 0;
            a.cand.ode.base[slot] = Tb.ctimes[0];
            a.cand.ode.base[slot + (long long)nrow * a.cand.ode.sj] = S;
            a.cand.ode.base[slot + 2LL * nrow * a.cand.ode.sj] = I;
            a.cand.latent.base[slot] = Tb.ctimes[0];
            a.cand.latent.base[slot + (long long)nrow * a.cand.latent.sj] = S;
            a.cand.latent.base[slot + 2LL * nrow * a.cand.latent.sj] = I;
            a.cand.betaN.base[slot] = beta;
            volz_row(0, S, I);
        }
        const bool run = !done && cand != 0;
        if (!__any_sync(full, run)) break;
        ticks += 1;
        // ---- 2. one LNA interval (sir_full_eval's step, same expressions) ---------------------------------------------
        double p00 = 1.0, p01 = 0.0, p10 = 0.0, p11 = 1.0, s00 = 0.0, s01 = 0.0, s11 = 0.0;
        if (run) eps.stage(seg, 2, Tb.eps_slot);
        if (run) {
            const double dts = Tb.dt_seg[seg];
            const double c1 = fma(-dts, gam, 1.0);
            const double c2 = fma(-2.0 * dts, gam, 1.0);
            LNA_DO_PRAGMA(unroll LNA_UNROLL)
            for (int q = 0; q < g; ++q) {
                const double bI = beta * I, bS = beta * S;
                const double a1 = bS * I;
                const double gI = gam * I;
                const double t0_ = bS * p10, t1_ = bS * p11;
                const double u0 = fma(bI, p00, t0_);
                const double u1 = fma(bI, p01, t1_);
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
        }
        cp_async_wait_all();
        // ---- 3. interval epilogue --------------------------------------------------------------------------------------
        if (run) {
            const double e0 = Tb.eps_slot[0], e1 = Tb.eps_slot[1];
            const double c00 = s00 + kJitter;
            const double l00 = sqrt(c00);
            const double l10 = s01 / l00;
            const double c11 = (s11 + kJitter) - l10 * l10;
            const double l11 = sqrt(c11);
            if (!(c00 > 0.0) || !(c11 > 0.0)) st |= ST_CHOL_FAIL;
            const double nX0 = fma(p00, X0, p01 * X1) + l00 * e0;
            const double nX1 = fma(p10, X0, p11 * X1) + fma(l10, e0, l11 * e1);
            X0 = nX0;
            X1 = nX1;
            const int r = seg + 1;
            apply_changes(r * g);
            const double lS = S + X0, lI = I + X1;
            {
                const long long NC = a.cand.A.sj, oA = slot + (long long)seg * 4 * NC, oR = slot + (long long)r * NC;
                const long long oS = oR + (long long)nrow * NC, oI = oS + (long long)nrow * NC;
                double *pA = a.cand.A.base + oA, *pL = a.cand.L.base + oA;
                pA[0] = p00;
                pA[NC] = p10;
                pA[2 * NC] = p01;
                pA[3 * NC] = p11;
                pL[0] = l00;
                pL[NC] = l10;
                pL[2 * NC] = 0.0;
                pL[3 * NC] = l11;
                const double tr = Tb.ctimes[r];
                a.cand.ode.base[oR] = tr;
                a.cand.ode.base[oS] = S;
                a.cand.ode.base[oI] = I;
                a.cand.latent.base[oR] = tr;
                a.cand.latent.base[oS] = lS;
                a.cand.latent.base[oI] = lI;
                a.cand.betaN.base[oR] = beta;
            }
            volz_row(r, lS, lI);
            seg += 1;
            bool ended = false, reject = false;
            if (seg == k) {
                ended = true;
            } else if (!revert_cur && ((st & ST_CHOL_FAIL) || (flags & (ST_NEG_STATE | ST_NAN)) || acc + __ldg(P.rowUB + r) < logy_lo)) {
                ended = true;  // a rejection whatever follows (:1641-1651 for the failures)
                reject = true;
            }
            if (ended) {
                double ll;
                int stf = 0;
                if (reject) {
                    ll = kSentinel;
                } else if (st & ST_CHOL_FAIL) {
                    stf = ST_CHOL_FAIL;
                    ll = kSentinel;
                } else if (flags & ST_NEG_STATE) {
                    stf = ST_NEG_STATE;
                    ll = kSentinel;
                } else if (flags & ST_NAN) {
                    stf = ST_NAN;
                    ll = kSentinel;
                } else {
                    ll = acc;
                }
                const bool failed = (stf & ST_CHOL_FAIL) != 0;
                if (revert_cur || (ll > logy && !failed)) {
                    stf |= revert_cur ? ST_CAP_HIT : 0;
                    a.win_slot[c] = (stf & ST_CHOL_FAIL) ? -1 : (int)slot;
                    a.cand_ll[c] = ll;
                    a.theta[c] = th_cur;
                    a.n_evals[c] = cand;
                    a.status[c] = stf;
                    done = true;
                }
                cand = 0;
            }
        }
    }
    // ---- park the unresolved chains for the next launch -------------------------------------------------------------------
    const bool park = have && !done;
    const unsigned pm = __ballot_sync(full, park);
    if (pm) {
        int base = 0;
        if (lane == __ffs(pm) - 1) base = atomicAdd(count_out, __popc(pm));
        base = __shfl_sync(full, base, __ffs(pm) - 1);
        if (park) {
            list_out[base + __popc(pm & ((1u << lane) - 1u))] = (int)c;
            SI_(0) = cand;
            SI_(1) = next;
            SI_(2) = seg;
            SI_(3) = nxt;
            SI_(4) = flags;
            SI_(5) = st;
            SI_(6) = ts.j;
            SD_(0) = S;
            SD_(1) = I;
            SD_(2) = X0;
            SD_(3) = X1;
            SD_(4) = R0cum;
            SD_(5) = acc;
            SD_(6) = ts.th;
            SD_(7) = ts.lo;
            SD_(8) = ts.hi;
            SD_(9) = th_cur;
        }
    }
    if (a.exec_count) {  // executed work: candidates started, warp-ticks (an interval costs the pipe the same whatever the lane mask)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) started += __shfl_xor_sync(full, started, o);
        if (lane == 0) {
            atomicAdd(a.exec_count, (unsigned long long)started);
            atomicAdd(a.exec_count + 2, (unsigned long long)ticks);
        }
    }
}

}  // namespace lna
