// lna_async.cuh -- ESlice_par_General (src/LNA_functional.cpp:1532-1688) for resident chains WITHOUT wave barriers and with
// PROVABLE EARLY REJECTION of slice candidates (sm_100a).
//
// Two facts about the reference's shrink loop:
//  (1) the angles theta_1, theta_2, .. depend only on the pre-drawn uniforms, so candidates can be evaluated speculatively
//      (lna_eslice.cuh);
//  (2) a candidate is REJECTED iff its coalescent log-likelihood is <= logy, and that log-likelihood is a sum over the
//      genealogy's grid cells of  -2 a_i lambda + y_i log(lambda)  (a_i = sum C D, y_i = number of coalescences of cell i,
//      lambda = beta S / I at the cell's trajectory row) -- each term is at most  y_i (log(y_i / (2 a_i)) - 1)  whatever the
//      trajectory does.  The time loop visits the rows in forward order, so after row r the terms still to come are bounded
//      by a constant of the genealogy, rowUB[r]: as soon as  partial + rowUB[r] < logy  the candidate is rejected, exactly
//      as the reference would reject it after evaluating it to the end (a Cholesky failure or a negative state later on is
//      a rejection too in ESlice_par_General, :1641-1651).  Wide-angle candidates are thrown out after a quarter of the
//      trajectory on the vignette problems (tools/early_exit_probe.py).
// The wave kernel cannot use (2): a wave lasts as long as its slowest lane.  Here every chain owns W lanes of a warp; a lane
// whose candidate is rejected takes the chain's next unstarted candidate in the same tick (so the candidate that will be
// accepted never queues behind one particular lane's slow rejection).  All
// lanes of the warp advance in lockstep by ONE LNA INTERVAL (g fine steps + the interval epilogue) per tick, each on its own
// (candidate, interval); the chain is resolved when some lane has accepted candidate i and every lane is past i.  Same
// decisions, same accepted candidate, same stored outputs as the sequential loop.
// A tick costs the FP64 pipe the same whatever the lane mask, so idle lanes are the enemy: the kernel is PERSISTENT -- a fixed
// number of warps per SM sub-partition -- and a group of W lanes whose chain is resolved takes the next chain from a global
// counter in the same tick, so the lanes stay busy until the batch is exhausted.
//
// Condition (checked on the host, else the wave kernels run): change points on LNA-interval boundaries, so that every lane
// runs the same g-step loop whatever interval it is in (true for the reference's vignettes: changetime on the coarse grid).
#pragma once
#include "lna_eslice.cuh"

namespace lna {

constexpr int kAsyncCandDoubles = 7;
// doubles of shared memory per group of W lanes (host and device)
__host__ __device__ inline int async_group_doubles(int nch) { return kMaxUnif + 5 + 4 + 3 * nch + 20 * kAsyncCandDoubles; }

template <int W>
__global__ void __launch_bounds__(128, 3)
k_ess_par_async(const __grid_constant__ DevProb P, const __grid_constant__ SegDt SD, long long B, CandArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables T = stage_tables(P, smem, true);
    const int w = threadIdx.x % W;  // (blockDim is a multiple of 32 >= W)
    // per-group copy of what every candidate START needs -- the call's uniforms, the chain's scalars and its change ratios
    // with their normals -- so that starting a candidate touches no global memory (a start happens in almost every tick of a
    // warp; its latency would be paid by all 32 lanes):  u[22] | S0 I0 R0 gamma hyper | nu_I0 nu_R0 nu_gamma nu_hyper | ch[nch] | nu[nch]
    // ... | z[nch] = log(ch) hyper (the chain's part of every proposed change ratio) | cs[20][7] = (theta, cos, sin, I0, R0,
    // gamma, hyper') of candidates 1..20, computed ONCE per chain by the group's lanes in parallel: a candidate start -- which
    // happens in a third of a warp's ticks and is paid by all 32 lanes -- is seven shared-memory loads instead of a sincos and
    // four exponentials
    const int gs_stage = kMaxUnif + 5 + 4 + 2 * P.nch, gs_len = async_group_doubles(P.nch);
    double *gs = reinterpret_cast<double *>(smem + full_eval_smem_bytes(tables_bytes(P.nch, P.k, P.n0), blockDim.x)) +
                 (size_t)(threadIdx.x / W) * gs_len;
    double *g_par = gs + kMaxUnif, *g_nrm = g_par + 5, *g_ch = g_nrm + 4, *g_nu = g_ch + P.nch, *g_z = g_nu + P.nch,
           *g_cs = g_z + P.nch;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const unsigned gmask = (W >= 32) ? full : (((1u << W) - 1u) << (lane & ~(W - 1)));
    const int leader = lane & ~(W - 1);
    const int npar = 2 + P.npt, g = P.g, k = P.k, nrow = P.nrow;
    constexpr int kMaxCand = 20;
    const double h2 = P.h2, dt6 = P.dt6, invN = P.invN;
    long long c = 0;          // the group's chain
    long long slot = 0;       // scratch slot of (chain, lane): c * W + w -- the winner's outputs stay where the commit kernels look
    auto unif = [&](int q) { return gs[q]; };
    EpsFromView eps{a.origin, 0, P.k};
    EssParPre zpre{};
    double logy = 0.0, logy_lo = 0.0;
    bool have_chain = false, exhausted = false;

    // ---- lane state -------------------------------------------------------------------------------------------------
    int cand = 0;             // candidate being evaluated (0 = none)
    int gnext = 1;            // the chain's next unstarted candidate (uniform within the group)
    int seg = 0;              // interval the current candidate is in
    int accepted = 0;         // candidate this lane accepted (0 = none); the lane stops there
    bool resolved = true;     // the chain is decided (or the group has no chain yet)
    double S = 0, I = 0, X0 = 0, X1 = 0, R0cum = 0, beta = 0, hb = 0, gam = 0, hg = 0, dg6 = 0;
    double acc = 0, th_cur = 0, ll_acc = 0;
    int flags = 0, st = 0, nxt = 0, st_acc = 0, ticks = 0;
    bool revert_cur = false;
    PropScalars x{};

    auto apply_changes = [&](int s) {  // every change point with fine index <= s (they sit on interval boundaries)
        bool any = false;
        while (nxt < P.nch && T.chidx[nxt] <= s) {
            // ChProp<PROP_ESS_PAR>::eval with log(ch) hyper taken from the per-chain table
            double f = g_ch[nxt];
            if (x.rot_ch) f = exp(__ddiv_rn(rot(g_z[nxt], x.ct, g_nu[nxt], x.sn), x.hyper_new));  // :1299, :1310
            R0cum *= f;
            ++nxt;
            any = true;
        }
        if (any) {
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
        }
    };
    auto volz_row = [&](int r, double lS, double lI) {  // VolzAcc::row
        if (r <= P.n0 && ((lS < 0.0) | (lI < 0.0))) flags |= ST_NEG_STATE;
        const int i = P.Lrow - r;
        if (i >= 0 && i < P.n0 && T.cellCnt[i] > 0) {
            const double ratio = lS / lI;
            const double term = fma(-2.0 * T.cellCD[i] * beta, ratio, T.cellY[i] * log(beta * ratio));
            if (!(lS > 0.0) || !(lI > 0.0) || term != term) flags |= ST_NAN;
            acc += term;
        }
    };

    for (;;) {
        // ---- 0. groups whose chain is decided take the next chain of the batch -------------------------------------
        {
            const bool need = resolved && !exhausted;  // (uniform within a group)
            if (__any_sync(full, need)) {
                long long nc = -1;
                if (need && w == 0) nc = (long long)atomicAdd(a.work_counter, 1);
                nc = __shfl_sync(full, nc, leader);
                if (need) {
                    if (nc < B) {
                        c = nc;
                        slot = c * W + w;
                        eps.b = c;
                        __syncwarp(gmask);  // (the group's previous chain is not read any more)
                        for (int q = w; q < gs_stage; q += W) {
                            double v;
                            if (q < kMaxUnif) v = a.uniforms.at(c, q);
                            else if (q < kMaxUnif + 4) v = a.par.at(c, q - kMaxUnif);
                            else if (q == kMaxUnif + 4) v = a.par.at(c, npar - 1);
                            else if (q < kMaxUnif + 8) v = a.normals.at(c, q - (kMaxUnif + 5));
                            else if (q == kMaxUnif + 8) v = a.normals.at(c, npar - 2);
                            else if (q < kMaxUnif + 9 + P.nch) v = a.par.at(c, 4 + (q - (kMaxUnif + 9)));
                            else v = a.normals.at(c, 3 + (q - (kMaxUnif + 9 + P.nch)));
                            gs[q] = v;
                        }
                        __syncwarp(gmask);
                        // Param_Slice_update_all2's whitened scalars (:1287-1296) = ess_par_pre
                        zpre.z0 = __ddiv_rn(__dsub_rn(log(g_par[1]), a.pr[0]), a.pr[1]);
                        zpre.z1 = __ddiv_rn(__dsub_rn(log(g_par[2]), a.pr[2]), a.pr[3]);
                        zpre.z2 = __ddiv_rn(__dsub_rn(log(g_par[3]), a.pr[4]), a.pr[5]);
                        zpre.zh = __ddiv_rn(__dsub_rn(log(g_par[4]), a.pr[6]), a.pr[7]);
                        logy = a.coal_log[c] + log(unif(0));
                        // (rounding head-room of the bound test: the partial sum carries ~1e-13 relative error)
                        logy_lo = logy - 1e-9 * (fabs(logy) + 1.0);
                        for (int q = w; q < P.nch; q += W) g_z[q] = __dmul_rn(log(g_ch[q]), g_par[4]);  // :1296
                        {   // candidates w+1, w+1+W, ..: ess_par_scalars on the staged copy (same expressions)
                            ThetaSeq ts;
                            ts.init();
                            for (int j = w + 1; j <= kMaxCand; j += W) {
                                ts.advance(j, unif);
                                double sn, ct;
                                sincos(ts.th, &sn, &ct);
                                double *cs = g_cs + (j - 1) * kAsyncCandDoubles;
                                cs[0] = ts.th;
                                cs[1] = ct;
                                cs[2] = sn;
                                cs[3] = a.ess[0] ? exp(__dadd_rn(__dmul_rn(rot(zpre.z0, ct, g_nrm[0], sn), a.pr[1]), a.pr[0])) : g_par[1];
                                cs[4] = a.ess[1] ? exp(__dadd_rn(__dmul_rn(rot(zpre.z1, ct, g_nrm[1], sn), a.pr[3]), a.pr[2])) : g_par[2];
                                cs[5] = a.ess[2] ? exp(__dadd_rn(__dmul_rn(rot(zpre.z2, ct, g_nrm[2], sn), a.pr[5]), a.pr[4])) : g_par[3];
                                cs[6] = a.ess[4] ? exp(__dadd_rn(__dmul_rn(rot(zpre.zh, ct, g_nrm[3], sn), a.pr[7]), a.pr[6])) : g_par[4];
                            }
                        }
                        __syncwarp(gmask);
                        gnext = 1;
                        cand = 0;
                        accepted = 0;
                        resolved = false;
                        have_chain = true;
                    } else {
                        exhausted = true;
                        have_chain = false;
                    }
                }
            }
        }
        // ---- 1. lanes without a candidate take their next one -----------------------------------------------------
        // the group's earliest accepted candidate so far: nothing beyond it is worth starting
        int first_acc = accepted ? accepted : 0x7fffffff;
#pragma unroll
        for (int o = 1; o < W; o <<= 1) first_acc = min(first_acc, __shfl_xor_sync(full, first_acc, o));
        // free lanes take the chain's next unstarted candidates in lane order (candidates start in increasing order, so a
        // lane's bracket replay only ever moves forward and everything below an accepted candidate has been started)
        const bool want = !resolved && cand == 0 && !accepted;
        const unsigned wm = __ballot_sync(full, want) & gmask;
        const int mine = gnext + __popc(wm & ((1u << lane) - 1u));
        const int lim = min(kMaxCand + 2, first_acc);  // candidates < lim are worth starting
        const bool start = want && mine < lim;
        if (!resolved) gnext += max(0, min(__popc(wm), lim - gnext));
        if (a.exec_count) {
            const unsigned sm_ = __ballot_sync(full, start);
            if (sm_ && lane == 0) atomicAdd(a.exec_count, (unsigned long long)__popc(sm_));
        }
        if (start) {
            cand = mine;
            revert_cur = cand == kMaxCand + 1;  // the cap: back to the current parameters (:1609-1622)
            {
                const double *cs = g_cs + (revert_cur ? 0 : cand - 1) * kAsyncCandDoubles;
                th_cur = cs[0];
                x.ct = cs[1];
                x.sn = cs[2];
                x.I0 = cs[3];
                x.R0 = cs[4];
                x.gam = cs[5];
                x.hyper_new = cs[6];
                x.hyper_old = g_par[4];
                x.rot_ch = a.ess[5];
            }
            x.S0 = g_par[0];
            if (revert_cur) {  // current parameters, bit for bit
                th_cur = 0.0;
                x.I0 = g_par[1];
                x.R0 = g_par[2];
                x.gam = g_par[3];
                x.rot_ch = 0;
            }
            x.noop = 0;
            gam = x.gam;
            hg = h2 * gam;
            dg6 = dt6 * gam;
            S = x.S0;
            I = x.I0;
            X0 = X1 = 0.0;
            R0cum = x.R0;
            nxt = 0;
            beta = (R0cum * gam) * invN;
            hb = h2 * beta;
            apply_changes(0);
            acc = 0.0;
            flags = 0;
            st = 0;
            seg = 0;
            if (a.cand.ode.ok()) {
                a.cand.ode.put(slot, 0, T.ctimes[0]);
                a.cand.ode.put(slot, nrow, S);
                a.cand.ode.put(slot, 2 * nrow, I);
            }
            if (a.cand.latent.ok()) {
                a.cand.latent.put(slot, 0, T.ctimes[0]);
                a.cand.latent.put(slot, nrow, S);
                a.cand.latent.put(slot, 2 * nrow, I);
            }
            if (a.cand.betaN.ok()) a.cand.betaN.put(slot, 0, beta);
            volz_row(0, S, I);
        }
        const bool run = cand != 0;
        ticks += a.max_waves ? __popc(__ballot_sync(full, run)) : 1;  // (max_waves = 1: count LANE-ticks instead, experiments)
        if (!__any_sync(full, run)) {
            // nobody in the warp is evaluating: the batch is exhausted.  What the FP64 pipe was charged: one
            // warp-interval per tick whatever the lane mask, i.e. (ticks - 1) / k warp-evaluations
            if (a.exec_count && lane == 0) {
                if (a.max_waves) atomicAdd(a.exec_count + 1, (unsigned long long)((ticks + k / 2) / k));
                else atomicAdd(a.exec_count + 2, (unsigned long long)(ticks - 1));  // warp-ticks
            }
            break;
        }
        // ---- 2. one LNA interval: g fine steps (sir_full_eval's step, same expressions) ----------------------------
        double p00 = 1.0, p01 = 0.0, p10 = 0.0, p11 = 1.0, s00 = 0.0, s01 = 0.0, s11 = 0.0;
        if (run) eps.stage(seg, 2, T.eps_slot);
        if (run) {
            // SigmaF's dt is the interval's own t[1] - t[0] (A.3): equal for all intervals up to the last rounding of the time
            // grid, but not bit-equal, so every lane reads the dt of the interval it is in
            const double dts = T.dt_seg[seg];
            const double c1 = fma(-dts, gam, 1.0);        // 1 - gam dts
            const double c2 = fma(-2.0 * dts, gam, 1.0);  // 1 - 2 gam dts
            LNA_DO_PRAGMA(unroll LNA_UNROLL)
            for (int s = 0; s < g; ++s) {
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
        // ---- 3. interval epilogue ----------------------------------------------------------------------------------
        bool ended = false, reject = false;
        if (run) {
            const double e0 = T.eps_slot[0], e1 = T.eps_slot[1];
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
            apply_changes(r * g);  // beta at the coarse time (betaTs on coarse times) = beta of the next interval
            const double lS = S + X0, lI = I + X1;
            {   // candidate outputs, slot-major scratch (element j of slot s at base[s + j NC]): one 64-bit offset per array
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
                const double tr = T.ctimes[r];
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
            if (seg == k) {
                ended = true;
            } else if (!revert_cur && ((st & ST_CHOL_FAIL) || (flags & (ST_NEG_STATE | ST_NAN)) || acc + __ldg(P.rowUB + r) < logy_lo)) {
                // a failed factorisation / a negative state in the checked rows / a NaN term make the final value the -1e7
                // sentinel, i.e. a rejection (:1641-1651); so does a partial sum that the remaining cells cannot lift over logy
                ended = true;
                reject = true;
            }
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
                accepted = cand;
                ll_acc = ll;
                st_acc = stf | (revert_cur ? ST_CAP_HIT : 0);
            }
            cand = 0;
        }
        // ---- 4. is the chain decided?  some lane accepted candidate i and no lane is still working below i --------------
        int gacc = accepted ? accepted : 0x7fffffff;
#pragma unroll
        for (int o = 1; o < W; o <<= 1) gacc = min(gacc, __shfl_xor_sync(full, gacc, o));
        // a lane is "below" the accepted candidate if it is evaluating a candidate < gacc (all of those have been started)
        const bool below = !resolved && !accepted && cand != 0 && cand < gacc;
        const unsigned pend = __ballot_sync(full, below) & gmask;
        if (!resolved && gacc != 0x7fffffff && pend == 0) {
            if (accepted == gacc) {  // the winner writes the chain's result
                a.win_slot[c] = (st_acc & ST_CHOL_FAIL) ? -1 : (int)slot;
                a.cand_ll[c] = ll_acc;
                a.theta[c] = th_cur;
                a.n_evals[c] = accepted;
                a.status[c] = st_acc;
            }
            resolved = true;
            cand = 0;  // lanes evaluating candidates beyond the accepted one drop them
        }
    }
}

}  // namespace lna
