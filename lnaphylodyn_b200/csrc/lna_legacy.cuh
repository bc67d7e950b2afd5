// lna_legacy.cuh -- the reference's LEGACY copies of the pipeline, sm_100a.
//
// Reference (file:line into MingweiWilliamTang/LNAphyloDyn):
//   ESlice_general               src/LNA_NS_general.cpp:320-394
//   Foo::LNA_ESlice              src/generalLNA.cpp:296-372   (the same text as ESlice_general)
//   ESlice_SIRS                  src/SIRS_period.cpp:313-402
//   volz_loglik_nh               src/SIR_phylodyn.cpp:517-560 (on LogTraj(.), src/SIR_LNA.cpp:40-46)
//   coal_loglik                  src/SIR_phylodyn.cpp:436-468
//   Foo::rate_h / dh / ODE_ones / ODE_rk45   src/generalLNA.cpp:76-122
// The integrator / moment / centred-density functions of LNA_NS_general.cpp are arithmetic twins of the current ones
// (ODE_rk45, SigmaF, KF_param on the SIR model; k_centred with the inv2 density) and reuse their kernels.
// None of this is a hot path: one thread per item, plain loops.
#pragma once
#include "lna_kernels.cuh"

namespace lna {

struct LegacyEssArgs {
    const double *f_cur;     // B x nrow*(p+1)
    const double *ode;       // B x nrow*(p+1)
    const double *v;         // reps x B x nrow*(p+1): the centred proposals of every repetition (k_centred mode 1 / 3)
    const double *betaN;     // B x nrow (ESlice_SIRS: betaDyn(.) * N, computed by k_beta_dyn)
    const double *uniforms;  // B x reps*22, consumed contiguously: u, theta_1, one per shrink; per repetition
    const double *lambda;    // [B] Kingman scale (like = 1)
    const int *v_status;     // reps x B status of the proposal simulation (ST_CHOL_FAIL = the reference threw there)
    double *cur;             // B x nrow*(p+1) scratch: f_cur between repetitions
    double *newTraj;         // B x nrow*(p+1)
    int *n_evals, *status, *n_unif;
    int p, reps;
    int like;            // 0: volz_loglik_nh(LogTraj(.), betaN), 1: coal_loglik(LogTraj(.), lambda)
    int cap_keeps_last;  // ESlice_SIRS leaves the loop at the 20-shrink cap WITHOUT returning to f_cur (:377-380)
};

// One elliptical-slice update of the latent trajectory in the centred parametrisation, `reps` times.
__global__ void k_eslice_legacy(DevProb P, long long B, LegacyEssArgs a) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nrow = P.nrow, p = a.p, n = nrow * p;
    const long long per = (long long)nrow * (p + 1);
    const double *ode = a.ode + b * per, *bet = a.betaN + b * nrow, *unif = a.uniforms + b * (long long)a.reps * 22;
    double *cur = a.cur + b * per, *nt = a.newTraj + b * per;
    for (long long e = 0; e < per; ++e) cur[e] = a.f_cur[b * per + e];
    // L = first row whose time reaches t_correct (volz_loglik_nh :521-523, coal_loglik :438-441), on the item's own times
    int L = 0;
    while (L < nrow - 1 && cur[L] < P.t_correct) ++L;
    const int n0 = P.n0;
    const double lam = a.like == 1 ? a.lambda[b] : 1.0, llam = log(lam);
    // likelihood of trajectory T (states, standard scale); `neg` = some state of the trajectory is negative
    auto loglik = [&](const double *T, bool &neg) -> double {
        double mn = T[nrow];
        for (int e = nrow; e < nrow + n; ++e) mn = fmin(mn, T[e]);  // (fmin ignores NaN like arma's min())
        neg = mn < 0.0;
        if (a.like == 0) {
            for (int j = 1; j <= 2; ++j)
                for (int r = 0; r <= n0 && r < nrow; ++r)
                    if (log(T[r + (long long)j * nrow]) < 0.0) return kSentinel;  // a LOG state below zero (:524-526)
            double acc = 0.0;
            for (int i = 0; i < n0; ++i) {
                const int row = L - i;
                if (row < 0 || P.cellCnt[i] == 0) continue;
                const double f = log(T[row + 2LL * nrow]), s = log(T[row + nrow]), bt = bet[row];
                acc += -2.0 * (P.cellCD[i] * bt * exp(-f) * exp(s)) + P.cellY[i] * (log(bt) - f + s);
            }
            return acc;
        }
        double acc = 0.0;
        for (int c = 0; c < L && c < n0; ++c) {
            if (P.cellCnt[c] == 0) continue;
            const double f = log(T[(L - c) + 2LL * nrow]);
            acc += -lam * (P.cellCD[c] * exp(-f)) + P.cellY[c] * (llam - f);
        }
        return acc;
    };
    int iu = 0, evals = 0, st = 0;
    for (int count = 0; count < a.reps; ++count) {
        const double *v = a.v + ((long long)count * B + b) * per;
        if (a.v_status && (a.v_status[(long long)count * B + b] & ST_CHOL_FAIL)) {
            st |= ST_CHOL_FAIL;  // the reference's chol() threw inside the proposal simulation: the call ends there
            break;
        }
        bool neg;
        const double logy = loglik(cur, neg) + log(unif[iu++]);
        ++evals;
        double theta = 0.0 + (kTwoPi - 0.0) * unif[iu++];
        double th_min = theta - kTwoPi, th_max = theta;
        auto rotate = [&](double th) {
            double sn, ct;
            sincos(th, &sn, &ct);
            for (int e = nrow; e < nrow + n; ++e) nt[e] = ((cur[e] - ode[e]) * ct + (v[e] - ode[e]) * sn) + ode[e];
        };
        for (int r = 0; r < nrow; ++r) nt[r] = cur[r];
        rotate(theta);
        double ll = loglik(nt, neg);
        ++evals;
        int i = 0;
        while (neg || ll <= logy) {
            i += 1;
            if (i > 20) {
                if (!a.cap_keeps_last)
                    for (long long e = 0; e < per; ++e) nt[e] = cur[e];
                st |= ST_CAP_HIT;
                break;
            }
            if (theta < 0.0) th_min = theta;
            else th_max = theta;
            theta = th_min + (th_max - th_min) * unif[iu++];
            rotate(theta);
            ll = loglik(nt, neg);
            ++evals;
        }
        for (long long e = 0; e < per; ++e) cur[e] = nt[e];
    }
    a.n_evals[b] = evals;
    a.status[b] = st;
    a.n_unif[b] = iu;
}

// betaDyn(beta, alpha, times, period) * N (SIR_phylodyn.cpp:40-46, SIRS_period.cpp:325) on each item's own time column
__global__ void k_beta_dyn(long long B, int nrow, int ld, const double *traj, const double *params4, double *betat) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * nrow) return;
    const long long b = idx / nrow;
    const int r = (int)(idx % nrow);
    const double *q = params4 + b * 4;
    betat[idx] = q[0] * (1 + q[1] * sin(2 * 3.14159265358979323846264338327950288 * traj[b * (long long)ld + r] / q[2])) * q[3];
}

// volz_loglik_nh (SIR_phylodyn.cpp:517-560) on a LOG trajectory and the reference's per-event sum, one thread per item
__global__ void k_volz_nh_log(DevProb P, long long B, const double *f1, const double *betaN, double *out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nrow = P.nrow, n0 = P.n0;
    const double *T = f1 + b * (long long)nrow * 3, *bet = betaN + b * nrow;
    int L = 0;
    while (L < nrow - 1 && T[L] < P.t_correct) ++L;
    for (int j = 1; j <= 2; ++j)
        for (int r = 0; r <= n0 && r < nrow; ++r)
            if (T[r + (long long)j * nrow] < 0.0) {
                out[b] = kSentinel;
                return;
            }
    double acc = 0.0;
    for (int i = 0; i < n0; ++i) {
        const int row = L - i;
        if (row < 0) continue;
        const double f = T[row + 2LL * nrow], s = T[row + nrow], bt = bet[row];
        for (int e = P.evStart[i]; e < P.evStart[i + 1]; ++e)
            acc += -2.0 * (P.evC[e] * P.evD[e] * bt * exp(-f) * exp(s)) + P.evY[e] * (log(bt) - f + s);
    }
    out[b] = acc;
}

// class Foo (generalLNA.cpp): hazards h = (param0 param1 / N S I, param1 I, param2 N), effect matrix A = A_post - A_pre (3 x 2)
struct FooNet {
    double A[6];  // column-major 3 x 2
    double pn0, p1, p2N;
};
__device__ __forceinline__ void foo_rhs(const FooNet &f, const double *x, double *res) {
    const double h0 = f.pn0 * x[0] * x[1], h1 = f.p1 * x[1], h2 = f.p2N;
    res[0] = f.A[0] * h0 + f.A[1] * h1 + f.A[2] * h2;
    res[1] = f.A[3] * h0 + f.A[4] * h1 + f.A[5] * h2;
}
// mode 0: rate_h (3), 1: dh (3 x 2), 2: ODE_ones (2), 3: A' dh (2 x 2; what F_vec computes), 4: ODE_rk45 (n x 3);
// modes 5 / 6 = 2 / 3 with the infection rate given directly in param[0] (ODE_general_one / general_F of
// LNA_NS_general.cpp:52-72 on the SIR effect matrix, th1 = betaf(t))
__global__ void k_foo(long long B, int mode_in, int n, const double *A_pre, const double *A_post, const double *param,
                      const double *Nvec, const double *states_or_initial, const double *t, double *out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    FooNet f;
    for (int e = 0; e < 6; ++e) f.A[e] = A_post[b * 6 + e] - A_pre[b * 6 + e];
    const double N = Nvec[b], *par = param + b * 3, *x = states_or_initial + b * 2;
    const int mode = mode_in >= 5 ? mode_in - 3 : mode_in;
    f.pn0 = mode_in >= 5 ? par[0] : par[0] * par[1] / N;
    f.p1 = par[1];
    f.p2N = par[2] * N;
    if (mode == 0) {
        out[b * 3] = f.pn0 * x[0] * x[1];
        out[b * 3 + 1] = f.p1 * x[1];
        out[b * 3 + 2] = f.p2N;
    } else if (mode == 1 || mode == 3) {
        const double dh[6] = {f.pn0 * x[1], 0.0, 0.0, f.pn0 * x[0], f.p1, 0.0};
        if (mode == 1) {
            for (int e = 0; e < 6; ++e) out[b * 6 + e] = dh[e];
        } else {
            for (int c = 0; c < 2; ++c)
                for (int j = 0; j < 2; ++j)
                    out[b * 4 + j + 2 * c] = f.A[3 * j] * dh[3 * c] + f.A[3 * j + 1] * dh[3 * c + 1] + f.A[3 * j + 2] * dh[3 * c + 2];
        }
    } else if (mode == 2) {
        foo_rhs(f, x, out + b * 2);
    } else {
        double *O = out + b * (long long)n * 3;
        const double dt = t[1] - t[0];
        double X0[2], X1[2] = {x[0], x[1]}, k1[2], k2[2], k3[2], k4[2], xs[2];
        for (int i = 0; i < n; ++i) O[i] = t[i];
        O[n] = X1[0];
        O[2LL * n] = X1[1];
        for (int i = 1; i < n; ++i) {
            X0[0] = X1[0];
            X0[1] = X1[1];
            foo_rhs(f, X0, k1);
            for (int j = 0; j < 2; ++j) xs[j] = X0[j] + k1[j] * dt / 2;
            foo_rhs(f, xs, k2);
            for (int j = 0; j < 2; ++j) xs[j] = X0[j] + k2[j] * dt / 2;
            foo_rhs(f, xs, k3);
            for (int j = 0; j < 2; ++j) xs[j] = X0[j] + k3[j] * dt / 2;
            foo_rhs(f, xs, k4);
            for (int j = 0; j < 2; ++j) {
                X1[j] = X0[j] + (k1[j] / 6 + k2[j] / 3 + k3[j] / 3 + k4[j] / 6) * dt;
                O[i + (long long)(j + 1) * n] = X1[j];
            }
        }
    }
}

}  // namespace lna
