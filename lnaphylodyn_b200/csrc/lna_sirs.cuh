// lna_sirs.cuh -- fused evaluation of the seasonal SIRS model (p = 3) of src/SIRS_period.cpp (sm_100a).
//
// One THREAD per parameter draw: ODE2 (:174-193) -> SIRS_KOM_Filter / SIRS_IntSigma (:50-79, :147-168) -> log_like_trajSIRS
// (:82-118) in one pass over the fine grid, everything that evolves in time in registers (state 3, fundamental matrix 9,
// symmetric covariance 6), nothing but the per-interval results ever written.  What the reference composes from three
// exports and a 2001 x 4 fine trajectory in memory is here 2000 register-resident steps per draw.
//
//   param = (beta, gamma, mu, alpha), x_r[0] = period;  beta(t) = beta (1 + alpha sin(2 pi t / period))
//   rhs   dS = -b S I + mu R,  dI = b S I - gamma I,  dR = gamma I - mu R                       (SIRS_ODE :122-144)
//   F     = [[-bI, -bS, mu], [bI, bS - gamma, 0], [0, gamma, -mu]]                              (SIRS_Fm_LNA_period :6-14)
//   Q     = A' diag(h) A, A = [-1 1 0; 0 -1 1; 1 0 -1], h = (b S I, gamma I, mu R)              (:55-70)
//   F0 <- F0 + (F F0) dt,  Sigma <- Sigma + (Sigma F' + F Sigma + Q) dt  at the left endpoints, dt = t[1] - t[0] GLOBAL
//   density of a given trajectory X around the ODE: rows with time <= t_correct contribute
//       -log(DegenerateDet3(Sigma_i)) / 2 - (X_{i+1} - eta_{i+1} - A_i (X_i - eta_i))' PseudoInverse(Sigma_i) (...) / 2
//
// The two sine arguments of the reference, 2*pi*t/period in the rhs and t/period*2*pi in the moment step, differ in the
// last bit for some t; both are tabulated per problem on the host (the time grid is shared by all draws), so a step costs
// two broadcast loads and two FMAs instead of two sin() evaluations.
#pragma once
#include "lna_device.cuh"

namespace lna {

struct SirsOut {
    OutView A, S, ode;  // p*p*k (column-major slices), p*p*k, nrow*(p+1)
};

// DegenerateDet3 (:17-19) and the quadratic form through PseudoInverse (:22-48) of the rank-2 covariance, M(r,c) symmetric
__device__ __forceinline__ double sirs_interval_density(double m01, double m02, double m11, double m12, double m22,
                                                        double v0, double v1, double v2, int &flags) {
    const double tr = m11 + m22 - m01 - m02;
    const double dd = (m11 - m01) * (m22 - m02) - (m12 - m01) * (m12 - m02);  // DegenerateDet3
    const double sq = sqrt(tr * tr - 4 * dd);
    const double l1 = (tr + sq) / 2, l2 = (tr - sq) / 2;
    double e1[3], e2[3];
    e1[1] = (m12 - m01) / (m11 - m01 - l1);
    e1[0] = 1 - e1[1];
    e1[2] = -1;
    e2[1] = (m12 - m01) / (m11 - m01 - l2);
    e2[0] = 1 - e2[1];
    e2[2] = -1;
    const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    const double n2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
    const double p1 = (e1[0] / n1) * v0 + (e1[1] / n1) * v1 + (e1[2] / n1) * v2;
    const double p2 = (e2[0] / n2) * v0 + (e2[1] / n2) * v1 + (e2[2] / n2) * v2;
    const double term = -log(dd) * 0.5 - 0.5 * (p1 * p1 / l1 + p2 * p2 / l2);
    if (term != term) flags |= ST_NAN;
    return term;
}

// sinO[s] = sin(2 pi t[s] / period), sinF[s] = sin(t[s] / period * 2 pi): device tables of length nt
template <bool kStore>
__device__ __forceinline__ void sirs_full_eval(const DevProb &P, const double *__restrict__ sinO,
                                               const double *__restrict__ sinF, const Tables &T, double beta, double gam,
                                               double mu, double alpha, double S, double I, double R, const View &sde,
                                               long long bx, const SirsOut &out, long long b, double &loglik, int &status) {
    const double h2 = P.h2, dt6 = P.dt6, dt = P.dt_ode;
    const int g = P.g, k = P.k, nrow = P.nrow;
    const double ba = beta * alpha;
    double ll = 0.0;
    int fl = 0;
    if (kStore && out.ode.ok()) {
        out.ode.put(b, 0, T.ctimes[0]);
        out.ode.put(b, nrow, S);
        out.ode.put(b, 2 * nrow, I);
        out.ode.put(b, 3 * nrow, R);
    }
    // deviations of the given trajectory from the ODE at the previous coarse row
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    const bool dens = sde.ok();
    if (dens) {
        d0 = sde.at(bx, nrow) - S;
        d1 = sde.at(bx, 2 * nrow) - I;
        d2 = sde.at(bx, 3 * nrow) - R;
    }
    for (int seg = 0; seg < k; ++seg) {
        double f00 = 1, f01 = 0, f02 = 0, f10 = 0, f11 = 1, f12 = 0, f20 = 0, f21 = 0, f22 = 1;  // F0 row-major
        double s00 = 0, s01 = 0, s02 = 0, s11 = 0, s12 = 0, s22 = 0;                              // Sigma (symmetric)
        // the interval's row of the given trajectory: in flight during the time loop
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, xt = 0.0;
        const int r = seg + 1;
        if (dens) {
            xt = sde.at(bx, r);
            x0 = sde.at(bx, nrow + r);
            x1 = sde.at(bx, 2 * nrow + r);
            x2 = sde.at(bx, 3 * nrow + r);
        }
        const int s_end = (seg + 1) * g;
        // the two table entries of a step are loaded one step ahead (the tables hold nt = last step + 1 entries): with one warp
        // per SM sub-partition -- the 16 384-draw sweep -- nothing else hides a load that the whole step depends on
        // (ncu: 40 % of the kernel's stall samples sat on these two loads)
        double so = __ldg(sinO + seg * g), sf = __ldg(sinF + seg * g);
        LNA_DO_PRAGMA(unroll LNA_UNROLL)
        for (int s = seg * g; s < s_end; ++s) {
            const double bo = fma(ba, so, beta);  // rhs:    beta (1 + alpha sin(2 pi t / period))
            const double bf = fma(ba, sf, beta);  // moments: beta (1 + alpha sin(t / period * 2 pi))
            so = __ldg(sinO + s + 1);
            sf = __ldg(sinF + s + 1);
            // ---- Euler step of F0 and Sigma at the left endpoint --------------------------------------------------
            {
                const double bI = bf * I, bS = bf * S;
                const double h0 = bS * I, h1 = gam * I, hh2 = mu * R;
                // F X for X = F0 (columns c): row0 = mu X2c - u_c, row1 = u_c - gam X1c, row2 = gam X1c - mu X2c,
                // u_c = bI X0c + bS X1c
                const double u0 = fma(bI, f00, bS * f10), u1 = fma(bI, f01, bS * f11), u2 = fma(bI, f02, bS * f12);
                const double g0 = gam * f10, g1 = gam * f11, g2 = gam * f12;
                const double m0 = mu * f20, m1 = mu * f21, m2 = mu * f22;
                f00 = fma(dt, m0 - u0, f00); f01 = fma(dt, m1 - u1, f01); f02 = fma(dt, m2 - u2, f02);
                f10 = fma(dt, u0 - g0, f10); f11 = fma(dt, u1 - g1, f11); f12 = fma(dt, u2 - g2, f12);
                f20 = fma(dt, g0 - m0, f20); f21 = fma(dt, g1 - m1, f21); f22 = fma(dt, g2 - m2, f22);
                // M = F Sigma with the same pattern on Sigma's columns (s10 = s01, s20 = s02, s21 = s12)
                const double U0 = fma(bI, s00, bS * s01), U1 = fma(bI, s01, bS * s11), U2 = fma(bI, s02, bS * s12);
                const double G0 = gam * s01, G1 = gam * s11, G2 = gam * s12;
                const double W0 = mu * s02, W1 = mu * s12, W2 = mu * s22;
                const double M00 = W0 - U0, M01 = W1 - U1, M02 = W2 - U2;
                const double M10 = U0 - G0, M11 = U1 - G1, M12 = U2 - G2;
                const double M20 = G0 - W0, M21 = G1 - W1, M22 = G2 - W2;
                // Sigma += (Sigma F' + F Sigma + A' H A) dt = (M' + M + Q) dt
                s00 = fma(dt, fma(2.0, M00, h0 + hh2), s00);
                s01 = fma(dt, (M01 + M10) - h0, s01);
                s02 = fma(dt, (M02 + M20) - hh2, s02);
                s11 = fma(dt, fma(2.0, M11, h0 + h1), s11);
                s12 = fma(dt, (M12 + M21) - h1, s12);
                s22 = fma(dt, fma(2.0, M22, h1 + hh2), s22);
            }
            // ---- the reference's RK4 variant: all four stages at t[s], k4 at X0 + k3 dt/2 --------------------------
            {
                const double a1 = (bo * S) * I, c1 = gam * I, e1 = mu * R;
                const double k1S = e1 - a1, k1I = a1 - c1, k1R = c1 - e1;
                const double S2 = fma(h2, k1S, S), I2 = fma(h2, k1I, I), R2 = fma(h2, k1R, R);
                const double a2 = (bo * S2) * I2, c2 = gam * I2, e2 = mu * R2;
                const double k2S = e2 - a2, k2I = a2 - c2, k2R = c2 - e2;
                const double S3 = fma(h2, k2S, S), I3 = fma(h2, k2I, I), R3 = fma(h2, k2R, R);
                const double a3 = (bo * S3) * I3, c3 = gam * I3, e3 = mu * R3;
                const double k3S = e3 - a3, k3I = a3 - c3, k3R = c3 - e3;
                const double S4 = fma(h2, k3S, S), I4 = fma(h2, k3I, I), R4 = fma(h2, k3R, R);
                const double a4 = (bo * S4) * I4, c4 = gam * I4, e4 = mu * R4;
                const double k4S = e4 - a4, k4I = a4 - c4, k4R = c4 - e4;
                S = fma(dt6, fma(2.0, k2S + k3S, k1S + k4S), S);
                I = fma(dt6, fma(2.0, k2I + k3I, k1I + k4I), I);
                R = fma(dt6, fma(2.0, k2R + k3R, k1R + k4R), R);
            }
        }
        if (kStore) {
            if (out.A.ok()) {
                const double Am[9] = {f00, f10, f20, f01, f11, f21, f02, f12, f22};
#pragma unroll
                for (int e = 0; e < 9; ++e) out.A.put(b, seg * 9 + e, Am[e]);
            }
            if (out.S.ok()) {
                const double Sm[9] = {s00, s01, s02, s01, s11, s12, s02, s12, s22};
#pragma unroll
                for (int e = 0; e < 9; ++e) out.S.put(b, seg * 9 + e, Sm[e]);
            }
            if (out.ode.ok()) {
                out.ode.put(b, r, T.ctimes[r]);
                out.ode.put(b, nrow + r, S);
                out.ode.put(b, 2 * nrow + r, I);
                out.ode.put(b, 3 * nrow + r, R);
            }
        }
        if (dens) {
            const double n0 = x0 - S, n1 = x1 - I, n2 = x2 - R;
            if (xt <= P.t_correct) {
                const double v0 = n0 - fma(f00, d0, fma(f01, d1, f02 * d2));
                const double v1 = n1 - fma(f10, d0, fma(f11, d1, f12 * d2));
                const double v2 = n2 - fma(f20, d0, fma(f21, d1, f22 * d2));
                ll += sirs_interval_density(s01, s02, s11, s12, s22, v0, v1, v2, fl);
            }
            d0 = n0;
            d1 = n1;
            d2 = n2;
        }
    }
    status |= fl;
    loglik = ll;
}

// One draw per thread.  `group` > 1: the given trajectory is shared by `group` consecutive draws (item b / group of `sde`);
// a view with item stride 0 shares one trajectory / one initial state over the whole sweep;
// sde.base == nullptr: moments only (ODE2 + SIRS_KOM_Filter), no density.
template <bool kStore>
__global__ void __launch_bounds__(64, 6)
k_sirs_full(const __grid_constant__ DevProb P, const double *__restrict__ sinO, const double *__restrict__ sinF, long long B,
            View param, View initial, View sde, int group, SirsOut out, double *loglik, int *status) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables T = stage_tables(P, smem, false);
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long bx = group > 1 ? b / group : b;
    double ll = 0.0;
    int st = 0;
    sirs_full_eval<kStore>(P, sinO, sinF, T, param.at(b, 0), param.at(b, 1), param.at(b, 2), param.at(b, 3), initial.at(b, 0),
                           initial.at(b, 1), initial.at(b, 2), sde, bx, out, b, ll, st);
    if (loglik) loglik[b] = ll;
    if (status) status[b] = st;
}

}  // namespace lna
