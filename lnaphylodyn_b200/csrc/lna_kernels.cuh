// lna_kernels.cuh -- __global__ kernels of the LNA / coalescent path (sm_100a).  See lna_device.cuh.
#pragma once
#include "lna_device.cuh"

namespace lna {

// ================================================================================================
// HOT KERNEL: one FULL evaluation per thread (New_Param_List -> TransformTraj -> volz_loglik_nh2,
// reference LNA_functional.cpp:1217-1226), item-major inputs.
// ================================================================================================
template <int kModel, bool kStore>
__global__ void __launch_bounds__(128, 4)
k_full_loglik(const __grid_constant__ DevProb P, const __grid_constant__ SegDt SD, long long B, View param,
              View initial, View origin, FullOut out, double *coalLog, int *status, int group) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables T = stage_tables(P, smem, true);
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // chain-shared inputs: `group` consecutive items (the proposals of one chain) read the same initial state and the
    // same latent increments (item bc of `initial` / `origin`); group = 1 is the per-item layout
    const long long bc = group > 1 ? b / group : b;
    int st = 0;
    double ll;
    ChFromParam chs{param, b, P.nparam};
    if (kModel == 0) {
        const double R0 = param.at(b, P.iR0), gam = param.at(b, 1);
        if (origin.ok()) {
            EpsFromView eps{origin, bc, P.k};
            sir_full_eval<kStore>(P, SD, T, chs, eps, R0, gam, initial.at(bc, 0), initial.at(bc, 1), out, b, ll, st);
        } else {
            EpsZero eps;
            sir_full_eval<kStore>(P, SD, T, chs, eps, R0, gam, initial.at(bc, 0), initial.at(bc, 1), out, b, ll, st);
        }
    } else {
        const double R0 = param.at(b, P.iR0), mu = param.at(b, 1), gam = param.at(b, 2);
        if (origin.ok()) {
            EpsFromView eps{origin, bc, P.k};
            seir_full_eval<kStore>(P, SD, T, chs, eps, R0, mu, gam, initial.at(bc, 0), initial.at(bc, 1),
                                   initial.at(bc, 2), out, b, ll, st);
        } else {
            EpsZero eps;
            seir_full_eval<kStore>(P, SD, T, chs, eps, R0, mu, gam, initial.at(bc, 0), initial.at(bc, 1),
                                   initial.at(bc, 2), out, b, ll, st);
        }
    }
    if (coalLog) coalLog[b] = ll;
    if (status) status[b] = st;
}

// ================================================================================================
// BALANCED FULL kernel (SIR): one persistent block per SM, warp w on SM sub-partition w % 4.
// k_full_loglik's unit of work is a warp-load (32 evaluations x 2000 dependent steps) and a sub-partition's FP64 pipe is
// saturated by three of them, so a launch lasts ceil(warp-loads per sub-partition) x one load: the headline batch (65 536
// evaluations = 2048 loads on 592 sub-partitions = 3.46 each) runs 4 deep on 272 sub-partitions while 320 idle for a quarter
// of the launch.  Here the n loads of an SM are laid on a line of n x k LNA intervals and the line is cut into four equal
// segments, one per sub-partition; a load that a cut falls into is evaluated in two pieces by two warps on neighbouring
// sub-partitions, the running state (PartState, 64 B per trajectory) handed over through shared memory.  All four pipes of
// the SM then carry n / 4 loads.  Same expressions in the same order as the uncut evaluation: same bits.
// ================================================================================================
struct BalPiece {
    long long load;          // warp-load: items 32 load .. 32 load + 31
    int seg_begin, seg_end;  // LNA intervals of that load
    bool valid;
};
// piece `pidx` of sub-partition `sp` in block `j` of `nblocks`, L warp-loads in all
__host__ __device__ inline BalPiece bal_piece(long long L, int nblocks, int j, int k, int sp, int pidx) {
    const long long base = L / nblocks, rem = L % nblocks;
    const long long n = base + (j < rem ? 1 : 0), l0 = (long long)j * base + (j < rem ? j : rem);
    const long long c0 = ((long long)sp * n * k + 2) / 4, c1 = ((long long)(sp + 1) * n * k + 2) / 4;
    const long long start = pidx == 0 ? c0 : (c0 / k + pidx) * k;
    BalPiece pc{0, 0, 0, false};
    if (start >= c1) return pc;
    const long long ld = start / k, end = (ld + 1) * k < c1 ? (ld + 1) * k : c1;
    pc.load = l0 + ld;
    pc.seg_begin = (int)(start - ld * k);
    pc.seg_end = (int)(end - ld * k);
    pc.valid = true;
    return pc;
}
constexpr int kBalCuts = 3;
__host__ __device__ inline size_t bal_smem_bytes(size_t tab_bytes, int block) {
    return ((full_eval_smem_bytes(tab_bytes, block) + 15) & ~(size_t)15) + sizeof(PartState) * 32 * kBalCuts + 16;
}

template <bool kStore>
__global__ void __launch_bounds__(512, 1)
k_full_loglik_bal(const __grid_constant__ DevProb P, const __grid_constant__ SegDt SD, long long B, View param,
                  View initial, View origin, FullOut out, double *coalLog, int *status, int group) {
    extern __shared__ __align__(16) unsigned char smem[];
    PartState *hand = reinterpret_cast<PartState *>(
        smem + ((full_eval_smem_bytes(tables_bytes(P.nch, P.k, P.n0), blockDim.x) + 15) & ~(size_t)15));
    volatile int *flag = reinterpret_cast<volatile int *>(hand + 32 * kBalCuts);
    if (threadIdx.x < kBalCuts) flag[threadIdx.x] = 0;
    const Tables T = stage_tables(P, smem, true);  // (ends with __syncthreads)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sp = warp & 3;
    const BalPiece pc = bal_piece((B + 31) / 32, gridDim.x, blockIdx.x, P.k, sp, warp >> 2);
    if (!pc.valid) return;
    const long long b = pc.load * 32 + lane;
    const bool live = b < B;
    const bool waits = pc.seg_begin > 0, hands = pc.seg_end < P.k;
    if (waits) {  // the first piece of this load runs on sub-partition sp - 1
        if (lane == 0)
            while (flag[sp - 1] == 0) __nanosleep(400);
        __syncwarp();
        __threadfence_block();
    }
    if (live) {
        const long long bc = group > 1 ? b / group : b;
        int st = 0;
        double ll = 0.0;
        ChFromParam chs{param, b, P.nparam};
        const PartRange part{pc.seg_begin, pc.seg_end, hand + (sp > 0 ? sp - 1 : 0) * 32 + lane,
                             hand + (sp < kBalCuts ? sp : kBalCuts - 1) * 32 + lane};
        const double R0 = param.at(b, P.iR0), gam = param.at(b, 1);
        if (origin.ok()) {
            EpsFromView eps{origin, bc, P.k};
            sir_full_eval<kStore>(P, SD, T, chs, eps, R0, gam, initial.at(bc, 0), initial.at(bc, 1), out, b, ll, st, part);
        } else {
            EpsZero eps;
            sir_full_eval<kStore>(P, SD, T, chs, eps, R0, gam, initial.at(bc, 0), initial.at(bc, 1), out, b, ll, st, part);
        }
        if (!hands) {
            if (coalLog) coalLog[b] = ll;
            if (status) status[b] = st;
        }
    }
    if (hands) {
        __threadfence_block();
        __syncwarp();
        if (lane == 0) flag[sp] = 1;
    }
}

// MH decision of Update_Param (LNA_functional.cpp:1228-1241) on top of a finished FULL evaluation.
__global__ void k_mh_accept(long long B, const double *coalLog_new, const double *coal_log, const double *offset,
                            const double *unif, const int *status, int *accept) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double a = coalLog_new[b] - coal_log[b] + (offset ? offset[b] : 0.0);
    // a Cholesky failure is an exception in the reference: the R wrapper's tryCatch keeps the old state
    accept[b] = (!(status[b] & ST_CHOL_FAIL) && log(unif[b]) < a) ? 1 : 0;
}

// ================================================================================================
// Per-genealogy precompute: segmented reduction of the sorted events into per-cell sums.
// One warp per cell, lanes stride over the cell's events, shuffle-reduce.
// ================================================================================================
__global__ void k_cell_sums(int n0, const int *evStart, const double *C, const double *D, const double *y,
                            double *cellCD, double *cellY) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n0) return;
    double cd = 0.0, yy = 0.0;
    for (int e = evStart[warp] + lane; e < evStart[warp + 1]; e += 32) {
        cd = fma(C[e], D[e], cd);
        yy += y[e];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cd += __shfl_xor_sync(0xffffffffu, cd, o);
        yy += __shfl_xor_sync(0xffffffffu, yy, o);
    }
    if (lane == 0) {
        cellCD[warp] = cd;
        cellY[warp] = yy;
    }
}

// ================================================================================================
// Mirrors of the individual Rcpp exports (API completeness; not the fused hot path).
// ================================================================================================

// theta0(t) per param_transform (LNA_functional.cpp:63-86): cumulative product over cht_j <= t.
__device__ inline double beta_at(const DevProb &P, const View &param, long long b, double t) {
    double R0 = param.at(b, P.iR0);
    for (int i = 0; i < P.nch; ++i) {
        if (!(P.cht[i] <= t)) break;
        R0 *= param.at(b, P.nparam + i);
    }
    return R0 * param.at(b, P.iGamma) / P.N;
}

__global__ void k_betaTs(DevProb P, long long B, View param, const double *times, int m, double *betas) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // single sweep like betaTs (LNA_functional.cpp:31-58)
    double R0 = param.at(b, P.iR0);
    const double gam = param.at(b, P.iGamma);
    int i = 0;
    for (int j = 0; j < m; ++j) {
        while (i < P.nch && P.cht[i] <= times[j]) R0 *= param.at(b, P.nparam + i++);
        betas[b * m + j] = R0 * gam / P.N;
    }
}

__global__ void k_param_transform(DevProb P, long long B, const double *t, View param, double *param2) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int j = 1; j < P.npt; ++j) param2[b * P.npt + j] = param.at(b, j);
    param2[b * P.npt] = beta_at(P, param, b, t[b]);
}

// Generic small-matrix model pieces (used by the mirrors below; the hot path has fused versions).
template <int kModel>
struct Model;
template <>
struct Model<0> {  // SIR
    static constexpr int P = 2;
    __device__ static void rhs(const double *x, const double *th, double *d) {
        const double a = th[0] * x[0] * x[1];
        d[0] = -a;
        d[1] = a - th[1] * x[1];
    }
    __device__ static void F(const double *x, const double *th, double *Fm) {  // row-major Fm[r*P+c]
        Fm[0] = -th[0] * x[1]; Fm[1] = -th[0] * x[0];
        Fm[2] = th[0] * x[1];  Fm[3] = th[0] * x[0] - th[1];
    }
    __device__ static void Q(const double *x, const double *th, double *Qm) {  // A' diag(h) A, row-major
        const double h0 = th[0] * x[0] * x[1], h1 = th[1] * x[1];
        Qm[0] = h0;  Qm[1] = -h0;
        Qm[2] = -h0; Qm[3] = h0 + h1;
    }
};
template <>
struct Model<1> {  // SEIR
    static constexpr int P = 3;
    __device__ static void rhs(const double *x, const double *th, double *d) {
        const double a = th[0] * x[0] * x[2];
        d[0] = -a;
        d[1] = a - th[1] * x[1];
        d[2] = th[1] * x[1] - th[2] * x[2];
    }
    __device__ static void F(const double *x, const double *th, double *Fm) {
        Fm[0] = -th[0] * x[2]; Fm[1] = 0.0;    Fm[2] = -th[0] * x[0];
        Fm[3] = th[0] * x[2];  Fm[4] = -th[1]; Fm[5] = th[0] * x[0];
        Fm[6] = 0.0;           Fm[7] = th[1];  Fm[8] = -th[2];
    }
    __device__ static void Q(const double *x, const double *th, double *Qm) {
        const double h0 = th[0] * x[0] * x[2], h1 = th[1] * x[1], h2 = th[2] * x[2];
        Qm[0] = h0;  Qm[1] = -h0;     Qm[2] = 0.0;
        Qm[3] = -h0; Qm[4] = h0 + h1; Qm[5] = -h1;
        Qm[6] = 0.0; Qm[7] = -h1;     Qm[8] = h1 + h2;
    }
};

template <>
struct Model<2> {  // seasonal SIRS (src/SIRS_period.cpp): states (S, I, R), th = (beta(t), gamma, mu)
    static constexpr int P = 3;
    __device__ static void rhs(const double *x, const double *th, double *d) {  // SIRS_ODE :122-144
        d[0] = -th[0] * x[0] * x[1] + th[2] * x[2];
        d[1] = th[0] * x[0] * x[1] - th[1] * x[1];
        d[2] = th[1] * x[1] - th[2] * x[2];
    }
    __device__ static void F(const double *x, const double *th, double *Fm) {  // SIRS_Fm_LNA_period :6-14
        Fm[0] = -th[0] * x[1]; Fm[1] = -th[0] * x[0];        Fm[2] = th[2];
        Fm[3] = th[0] * x[1];  Fm[4] = th[0] * x[0] - th[1]; Fm[5] = 0.0;
        Fm[6] = 0.0;           Fm[7] = th[1];                Fm[8] = -th[2];
    }
    __device__ static void Q(const double *x, const double *th, double *Qm) {  // A' diag(h) A, A = [-1 1 0; 0 -1 1; 1 0 -1]
        const double h0 = th[0] * x[0] * x[1], h1 = th[1] * x[1], h2 = th[2] * x[2];
        Qm[0] = h0 + h2; Qm[1] = -h0;     Qm[2] = -h2;
        Qm[3] = -h0;     Qm[4] = h0 + h1; Qm[5] = -h1;
        Qm[6] = -h2;     Qm[7] = -h1;     Qm[8] = h1 + h2;
    }
};
template <>
struct Model<3> {  // SIR on the log scale (transX = "log"): states (log S, log I)
    static constexpr int P = 2;
    __device__ static void rhs(const double *x, const double *th, double *d) {  // ODE_SIR_one :117-119
        d[0] = -th[0] * exp(x[1]) - th[0] * exp(x[1] - x[0]) / 2;
        d[1] = th[0] * exp(x[0]) - th[1] - th[0] * exp(x[0] - x[1]) / 2 - th[1] * exp(-x[1]) / 2;
    }
    __device__ static void F(const double *x, const double *th, double *Fm) {  // SIR_F :256-258
        Fm[0] = th[0] * exp(x[1] - x[0]) / 2;
        Fm[1] = -th[0] * exp(x[1]) - th[0] / 2 * exp(x[1] - x[0]);
        Fm[2] = th[0] * exp(x[0]) - th[0] * exp(x[0] - x[1]) / 2;
        Fm[3] = th[0] * exp(x[0] - x[1]) / 2 + th[1] / 2 * exp(-x[1]);
    }
    __device__ static void Q(const double *x, const double *th, double *Qm) {  // Q A' diag(h) A Q, Q = diag(exp(-x)) :577-589
        const double h0 = th[0] * exp(x[0]) * exp(x[1]), h1 = th[1] * exp(x[1]);  // SIR_h :310-311
        const double q0 = exp(-x[0]), q1 = exp(-x[1]);
        Qm[0] = q0 * h0 * q0;  Qm[1] = q0 * -h0 * q1;
        Qm[2] = q1 * -h0 * q0; Qm[3] = q1 * (h0 + h1) * q1;
    }
};
constexpr double kPi = 3.14159265358979323846264338327950288;

// ODE_rk45 (LNA_functional.cpp:455-491): full fine trajectory, one thread per item.
template <int kModel>
__global__ void k_ode_rk45(DevProb P, long long B, View initial, View param, double *OdeTraj) {
    using M = Model<kModel>;
    constexpr int p = M::P;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double *out = OdeTraj + b * (long long)P.nt * (p + 1);
    double x[p], th[3] = {0.0, param.at(b, 1), P.npt > 2 ? param.at(b, 2) : 0.0};
    const double gam = param.at(b, P.iGamma);
    double R0 = param.at(b, P.iR0);
    int nxt = 0;
    const double dt = P.dt_ode;
    for (int c = 0; c < p; ++c) {
        x[c] = initial.at(b, c);
        out[(long long)(c + 1) * P.nt] = x[c];
    }
    out[0] = P.times[0];
    for (int i = 1; i < P.nt; ++i) {
        if (kModel == 2) {  // ODE2 (SIRS_period.cpp:174-193): beta(t) = beta (1 + alpha sin(2 pi t / period)), x_r[0] = period
            th[0] = param.at(b, 0) * (1 + param.at(b, 3) * sin(2 * kPi * P.times[i - 1] / P.N));
        } else {
            while (nxt < P.nch && P.chidx[nxt] <= i - 1) R0 *= param.at(b, P.nparam + nxt++);
            th[0] = R0 * gam / P.N;
        }
        double k1[p], k2[p], k3[p], k4[p], xs[p];
        M::rhs(x, th, k1);
        for (int c = 0; c < p; ++c) xs[c] = x[c] + k1[c] * dt / 2;
        M::rhs(xs, th, k2);
        for (int c = 0; c < p; ++c) xs[c] = x[c] + k2[c] * dt / 2;
        M::rhs(xs, th, k3);
        for (int c = 0; c < p; ++c) xs[c] = x[c] + k3[c] * dt / 2;
        M::rhs(xs, th, k4);
        out[i] = P.times[i];
        for (int c = 0; c < p; ++c) {
            x[c] = x[c] + (k1[c] / 6 + k2[c] / 3 + k3[c] / 3 + k4[c] / 6) * dt;
            out[i + (long long)(c + 1) * P.nt] = x[c];
        }
    }
}

// KF_param / KF_param_chol (LNA_functional.cpp:603-670) from a GIVEN fine trajectory:
// one thread per (item, segment) -- the k segments are independent.
template <int kModel>
__global__ void k_kf_param(DevProb P, long long B, const double *OdeTraj, View param, int chol, double *Acube,
                           double *Scube, int *status) {
    using M = Model<kModel>;
    constexpr int p = M::P;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= B * P.k) return;
    const long long b = tid / P.k;
    const int seg = (int)(tid % P.k);
    const double *tr = OdeTraj + b * (long long)P.nt * (p + 1);
    double F0[p * p], Sg[p * p], Fm[p * p], Qm[p * p], T1[p * p];
    for (int e = 0; e < p * p; ++e) F0[e] = 0.0, Sg[e] = 0.0;
    for (int c = 0; c < p; ++c) F0[c * p + c] = 1.0;
    double th[3] = {0.0, param.at(b, 1), P.npt > 2 ? param.at(b, 2) : 0.0};
    const int r0 = seg * P.g;
    // SigmaF uses the block's own first interval (:568); SIRS_KOM_Filter the global one (SIRS_period.cpp:149)
    const double dt = kModel == 2 ? tr[1] - tr[0] : tr[r0 + 1] - tr[r0];
    for (int i = 0; i < P.g; ++i) {
        const int r = r0 + i;
        th[0] = kModel == 2 ? param.at(b, 0) * (1 + param.at(b, 3) * sin(tr[r] / P.N * 2 * kPi))
                            : beta_at(P, param, b, tr[r]);
        double x[p];
        for (int c = 0; c < p; ++c) x[c] = tr[r + (long long)(c + 1) * P.nt];
        M::F(x, th, Fm);
        M::Q(x, th, Qm);
        // F0 += (F*F0)*dt
        for (int a = 0; a < p; ++a)
            for (int c = 0; c < p; ++c) {
                double s = 0.0;
                for (int q = 0; q < p; ++q) s += Fm[a * p + q] * F0[q * p + c];
                T1[a * p + c] = s;
            }
        for (int e = 0; e < p * p; ++e) F0[e] = F0[e] + T1[e] * dt;
        // Sigma += (Sigma*F' + F*Sigma + Q)*dt
        for (int a = 0; a < p; ++a)
            for (int c = 0; c < p; ++c) {
                double s1 = 0.0, s2 = 0.0;
                for (int q = 0; q < p; ++q) {
                    s1 += Sg[a * p + q] * Fm[c * p + q];
                    s2 += Fm[a * p + q] * Sg[q * p + c];
                }
                T1[a * p + c] = (s1 + s2) + Qm[a * p + c];
            }
        for (int e = 0; e < p * p; ++e) Sg[e] = Sg[e] + T1[e] * dt;
    }
    double *Ao = Acube + (b * P.k + seg) * (long long)(p * p);
    double *So = Scube + (b * P.k + seg) * (long long)(p * p);
    for (int a = 0; a < p; ++a)
        for (int c = 0; c < p; ++c) Ao[a + c * p] = F0[a * p + c];
    if (!chol) {
        for (int a = 0; a < p; ++a)
            for (int c = 0; c < p; ++c) So[a + c * p] = Sg[a * p + c];
        return;
    }
    // upper Cholesky of Sigma + 1e-8 I on the upper triangle (dpotrf 'U'), store L = U'
    double U[p * p];
    bool fail = false;
    for (int e = 0; e < p * p; ++e) U[e] = 0.0;
    for (int j = 0; j < p; ++j) {
        double ajj = Sg[j * p + j] + kJitter;
        for (int q = 0; q < j; ++q) ajj -= U[q * p + j] * U[q * p + j];
        if (!(ajj > 0.0)) fail = true;
        ajj = sqrt(ajj);
        U[j * p + j] = ajj;
        for (int c = j + 1; c < p; ++c) {
            double s = Sg[j * p + c];
            for (int q = 0; q < j; ++q) s -= U[q * p + j] * U[q * p + c];
            U[j * p + c] = s / ajj;
        }
    }
    for (int a = 0; a < p; ++a)
        for (int c = 0; c < p; ++c) So[a + c * p] = U[c * p + a];
    if (fail && status) atomicOr(&status[b], ST_CHOL_FAIL);
}

// Ode_Coarse_Slicer (LNA_functional.cpp:1179-1188)
__global__ void k_coarse_slicer(int nt, int ncol, int g, int nrow, long long B, const double *thin, double *coarse) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)nrow * ncol;
    if (tid >= B * per) return;
    const long long b = tid / per;
    const int e = (int)(tid % per), c = e / nrow, r = e % nrow;
    coarse[tid] = thin[b * (long long)nt * ncol + (long long)c * nt + (long long)r * g];
}

// TransformTraj (LNA_functional.cpp:877-900) and Traj_sim_general_noncentral (:824-872, eps supplied).
// mode 0: OriginLatent is k x p column-major; mode 1: `normals` step-major + extra outputs.
template <int p>
__global__ void k_transform_traj(int k, long long B, const double *Ode, const double *eps_in, const double *Acube,
                                 const double *Lcube, int step_major, double t_correct, double *traj,
                                 double *OriginLatent, double *logMultiNorm, double *logOrigin) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nrow = k + 1;
    const double *O = Ode + b * (long long)nrow * (p + 1);
    double *T = traj + b * (long long)nrow * (p + 1);
    const double *A = Acube + b * (long long)k * p * p, *L = Lcube + b * (long long)k * p * p;
    const double *ep = eps_in + b * (long long)k * p;
    double X[p], lm = 0.0, lo = 0.0;
    for (int c = 0; c < p; ++c) X[c] = 0.0;
    for (int c = 0; c <= p; ++c) T[(long long)c * nrow] = O[(long long)c * nrow];
    for (int i = 0; i < k; ++i) {
        double e[p], nX[p];
        for (int c = 0; c < p; ++c) e[c] = step_major ? ep[i * p + c] : ep[i + c * k];
        const double *Ai = A + i * p * p, *Li = L + i * p * p;
        for (int r = 0; r < p; ++r) {
            double ax = Ai[r] * X[0], le = Li[r] * e[0];
            for (int q = 1; q < p; ++q) {
                ax += Ai[r + q * p] * X[q];
                le += Li[r + q * p] * e[q];
            }
            nX[r] = ax + le;
        }
        T[i + 1] = O[i + 1];
        for (int r = 0; r < p; ++r) {
            X[r] = nX[r];
            T[(i + 1) + (long long)(r + 1) * nrow] = nX[r] + O[(i + 1) + (long long)(r + 1) * nrow];
        }
        if (step_major) {
            for (int c = 0; c < p; ++c) OriginLatent[b * (long long)k * p + i + c * k] = e[c];
            if (O[i + 1] <= t_correct) {
                double ee = 0.0, det;
                for (int c = 0; c < p; ++c) ee += e[c] * e[c];
                if (p == 2) det = Li[0] * Li[3] - Li[2] * Li[1];
                else
                    det = Li[0] * (Li[4] * Li[8] - Li[7] * Li[5]) - Li[3] * (Li[1] * Li[8] - Li[7] * Li[2]) +
                          Li[6] * (Li[1] * Li[5] - Li[4] * Li[2]);
                lm += -log(det);
                lo += -0.5 * ee;
            }
        }
    }
    if (step_major) {
        logMultiNorm[b] = lm;
        logOrigin[b] = lo;
    }
}

// Centred parametrisation, one thread per item:
//   mode 0: log_like_traj_general2 (LNA_functional.cpp:675-722): Gaussian transition density of a trajectory
//   mode 1: Traj_sim_general3 (:771-817): X1 = A (X0 - eta0) + eta1 + mvrnormArma(Sigma) with supplied normals
template <int p>
__device__ inline double det_small(const double *M) {
    if (p == 2) return M[0] * M[3] - M[2] * M[1];
    return M[0] * (M[4] * M[8] - M[7] * M[5]) - M[3] * (M[1] * M[8] - M[7] * M[2]) + M[6] * (M[1] * M[5] - M[4] * M[2]);
}
template <int p>
__device__ inline void solve_small(const double *S, const double *v, double *w) {
    const double d = det_small<p>(S);
    if (p == 2) {
        w[0] = (S[3] * v[0] - S[2] * v[1]) / d;
        w[1] = (-S[1] * v[0] + S[0] * v[1]) / d;
    } else {
        double C[9];
        C[0] = S[4] * S[8] - S[7] * S[5]; C[3] = -(S[3] * S[8] - S[6] * S[5]); C[6] = S[3] * S[7] - S[6] * S[4];
        C[1] = -(S[1] * S[8] - S[7] * S[2]); C[4] = S[0] * S[8] - S[6] * S[2]; C[7] = -(S[0] * S[7] - S[6] * S[1]);
        C[2] = S[1] * S[5] - S[4] * S[2]; C[5] = -(S[0] * S[5] - S[3] * S[2]); C[8] = S[0] * S[4] - S[3] * S[1];
        for (int r = 0; r < 3; ++r) w[r] = (C[r] * v[0] + C[r + 3] * v[1] + C[r + 6] * v[2]) / d;
    }
}

//   mode 3: Traj_sim_SIRS (SIRS_period.cpp:197-240): like mode 1 from a given initial state, noise = mvrnormArma2(Sigma) =
//           chol(Sigma + 1e-13 I_3)' z (basic_util.cpp:72-81) of the RANK-2 seasonal-SIRS covariance -- whether that
//           factorisation exists is decided by the last rounding of Sigma, so the outcome is reported per item (ST_CHOL_FAIL =
//           the reference's chol() exception) -- and the density through PseudoInverse / DegenerateDet3
//   mode | 8 (p = 2): the legacy copies' density, through inv2(Sigma) (basic_util.cpp:5-14) instead of solve():
//           log_like_traj_general / Traj_sim_general (LNA_NS_general.cpp:164-250), Foo::LNA_log_like_traj / LNA_Traj_sim
__device__ inline double quad_inv2(const double *S, const double *v) {  // (v' inv2(S)) v
    const double r00 = S[3], r11 = S[0], r10 = -S[1], r01 = -S[2];
    const double iD = 1 / (r00 * r11 - r10 * r01);
    const double w0 = v[0] * (r00 * iD) + v[1] * (r10 * iD), w1 = v[0] * (r01 * iD) + v[1] * (r11 * iD);
    return w0 * v[0] + w1 * v[1];
}
template <int p>
__global__ void k_centred(int k, long long B, int mode_flags, double t_correct, const double *Sde_or_normals,
                          const double *Ode, const double *Acube, const double *Scube, double *traj_out,
                          double *loglik, int *status, const double *initial) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int mode = mode_flags & 7;
    const bool inv2_density = (mode_flags & 8) && p == 2;
    const int nrow = k + 1;
    const double *O = Ode + b * (long long)nrow * (p + 1);
    const double *A = Acube + b * (long long)k * p * p, *S = Scube + b * (long long)k * p * p;
    double ll = 0.0;
    int st = 0;
    if (mode == 0 || mode == 2) {
        // mode 2 = log_like_trajSIRS (SIRS_period.cpp:82-118): rank-2 covariance, PseudoInverse / DegenerateDet3
        const double *X = Sde_or_normals + b * (long long)nrow * (p + 1);
        double d0[p], d1[p], v[p], w[p];
        for (int j = 0; j < p; ++j) d1[j] = X[(long long)(j + 1) * nrow] - O[(long long)(j + 1) * nrow];
        for (int i = 0; i < k; ++i) {
            const double *Ai = A + i * p * p, *Si = S + i * p * p;
            for (int j = 0; j < p; ++j) {
                d0[j] = d1[j];
                d1[j] = X[(i + 1) + (long long)(j + 1) * nrow] - O[(i + 1) + (long long)(j + 1) * nrow];
            }
            if (X[i + 1] <= t_correct) {
                for (int r = 0; r < p; ++r) {
                    double ax = 0.0;
                    for (int q = 0; q < p; ++q) ax += Ai[r + q * p] * d0[q];
                    v[r] = d1[r] - ax;
                }
                if (mode == 2 && p == 3) {
                    const double m01 = Si[3], m02 = Si[6], m11 = Si[4], m12 = Si[7], m22 = Si[8];
                    const double tr = m11 + m22 - m01 - m02;
                    const double dd = (m11 - m01) * (m22 - m02) - (m12 - m01) * (m12 - m02);  // DegenerateDet3
                    const double sq = sqrt(tr * tr - 4 * dd);
                    const double l1 = (tr + sq) / 2, l2 = (tr - sq) / 2;
                    double e1[3], e2[3];
                    e1[1] = (m12 - m01) / (m11 - m01 - l1); e1[0] = 1 - e1[1]; e1[2] = -1;
                    e2[1] = (m12 - m01) / (m11 - m01 - l2); e2[0] = 1 - e2[1]; e2[2] = -1;
                    const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
                    const double n2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
                    double p1 = 0.0, p2 = 0.0;  // projections on the two eigenvectors
                    for (int r = 0; r < 3; ++r) {
                        p1 += (e1[r] / n1) * v[r];
                        p2 += (e2[r] / n2) * v[r];
                    }
                    ll += -log(dd) * 0.5 - 0.5 * (p1 * p1 / l1 + p2 * p2 / l2);
                } else if (inv2_density) {
                    ll += -log(det_small<p>(Si)) / 2.0 - 0.5 * quad_inv2(Si, v);
                } else {
                    solve_small<p>(Si, v, w);
                    double q = 0.0;
                    for (int r = 0; r < p; ++r) q += v[r] * w[r];
                    ll += -log(det_small<p>(Si)) / 2.0 - 0.5 * q;
                }
            }
        }
    } else {
        const double *z = Sde_or_normals + b * (long long)k * p;
        double *T = traj_out + b * (long long)nrow * (p + 1);
        double X0[p], X1[p], e0[p], e1[p], noise[p], w[p];
        for (int j = 0; j < p; ++j) {
            X1[j] = (mode == 3 && initial) ? initial[b * p + j] : O[(long long)(j + 1) * nrow];
            e1[j] = X1[j];
            T[(long long)(j + 1) * nrow] = X1[j];
        }
        T[0] = 0.0;  // quirk of the reference (:782, SIRS_period.cpp:205)
        for (int i = 0; i < k; ++i) {
            const double *Ai = A + i * p * p, *Si = S + i * p * p;
            for (int j = 0; j < p; ++j) {
                X0[j] = X1[j];
                e0[j] = e1[j];
                e1[j] = O[(i + 1) + (long long)(j + 1) * nrow];
            }
            // mvrnormArma (basic_util.cpp:44-58): chols(Sigma) z for p = 2, chol(Sigma + 1e-13 I)^T z otherwise
            double L[p * p];
            if (p == 2) {
                const double r00 = sqrt(Si[0]), r01 = Si[1] / r00;
                L[0] = r00; L[1] = r01; L[2] = 0.0; L[3] = sqrt(Si[3] - r01 * r01);
            } else if (st & ST_CHOL_FAIL) {
                for (int e = 0; e < p * p; ++e) L[e] = 0.0;  // (the reference threw at the first failure: nothing after it counts)
            } else {
                double U[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                for (int j = 0; j < 3; ++j) {
                    double ajj = Si[j + j * 3] + 0.0000000000001;
                    for (int q = 0; q < j; ++q) ajj -= U[q + j * 3] * U[q + j * 3];
                    if (!(ajj > 0.0)) st |= ST_CHOL_FAIL;
                    ajj = sqrt(ajj);
                    U[j + j * 3] = ajj;
                    for (int c = j + 1; c < 3; ++c) {
                        double s2 = Si[j + c * 3];
                        for (int q = 0; q < j; ++q) s2 -= U[q + j * 3] * U[q + c * 3];
                        U[j + c * 3] = s2 / ajj;
                    }
                }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) L[r + c * 3] = U[c + r * 3];
            }
            for (int r = 0; r < p; ++r) {
                double s2 = 0.0;
                for (int q = 0; q < p; ++q) s2 += L[r + q * p] * z[i * p + q];
                noise[r] = s2;
            }
            for (int r = 0; r < p; ++r) {
                double ax = 0.0;
                for (int q = 0; q < p; ++q) ax += Ai[r + q * p] * (X0[q] - e0[q]);
                X1[r] = ax + e1[r] + noise[r];
            }
            if (O[i + 1] <= t_correct) {
                if (mode == 3 && p == 3) {
                    // -log(DegenerateDet3(Sig)) / 2 - noise' PseudoInverse(Sig) noise / 2 (:226-231), same closed form as mode 2
                    const double m01 = Si[3], m02 = Si[6], m11 = Si[4], m12 = Si[7], m22 = Si[8];
                    const double tr = m11 + m22 - m01 - m02;
                    const double dd = (m11 - m01) * (m22 - m02) - (m12 - m01) * (m12 - m02);
                    const double sq = sqrt(tr * tr - 4 * dd);
                    const double l1 = (tr + sq) / 2, l2 = (tr - sq) / 2;
                    double g1[3], g2[3];
                    g1[1] = (m12 - m01) / (m11 - m01 - l1); g1[0] = 1 - g1[1]; g1[2] = -1;
                    g2[1] = (m12 - m01) / (m11 - m01 - l2); g2[0] = 1 - g2[1]; g2[2] = -1;
                    const double n1 = sqrt(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
                    const double n2 = sqrt(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
                    double p1 = 0.0, p2 = 0.0;
                    for (int r = 0; r < 3; ++r) {
                        p1 += (g1[r] / n1) * noise[r];
                        p2 += (g2[r] / n2) * noise[r];
                    }
                    ll += -log(dd) * 0.5 + (-0.5) * (p1 * p1 / l1 + p2 * p2 / l2);
                } else if (inv2_density) {
                    ll += -log(det_small<p>(Si)) / 2 + (-0.5) * quad_inv2(Si, noise);
                } else {
                    solve_small<p>(Si, noise, w);
                    double q = 0.0;
                    for (int r = 0; r < p; ++r) q += noise[r] * w[r];
                    ll += -log(det_small<p>(Si)) / 2.0 + (-0.5) * q;
                }
            }
            T[i + 1] = O[i + 1];
            for (int j = 0; j < p; ++j) T[(i + 1) + (long long)(j + 1) * nrow] = X1[j];
        }
    }
    if (loglik) loglik[b] = ll;
    if (status) status[b] = st;
}

// volz_loglik_nh2 (LNA_functional.cpp:1119-1176), literal PER-EVENT form: one warp per trajectory.
// The block stages the genealogy's event arrays (C*D and y) in shared memory once; each warp first
// derives the per-cell factors of its trajectory (lane = cell), then strides over the events and
// reduces with warp shuffles.
__global__ void k_volz_events(DevProb P, long long B, const double *traj, const double *betaN, double *loglik,
                              int *status, int stage_events, int nan_guard) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *cellq = reinterpret_cast<double *>(smem);      // [warps][n0] beta*exp(-f)*exp(s)
    double *celll = cellq + (size_t)warps * P.n0;          // [warps][n0] log beta - f + s
    double *evCD = celll + (size_t)warps * P.n0;           // [E] (optional)
    double *evY = evCD + (stage_events ? P.E : 0);
    int *evCell = reinterpret_cast<int *>(evY + (stage_events ? P.E : 0));  // [E] (optional)
    if (stage_events) {
        for (int e = threadIdx.x; e < P.E; e += blockDim.x) {
            evCD[e] = P.evC[e] * P.evD[e];
            evY[e] = P.evY[e];
        }
        for (int c = 0; c < P.n0; ++c)
            for (int e = P.evStart[c] + threadIdx.x; e < P.evStart[c + 1]; e += blockDim.x) evCell[e] = c;
    }
    __syncthreads();
    const long long b = (long long)blockIdx.x * warps + w;
    if (b >= B) return;
    const int nrow = P.nrow, p = P.p;
    const double *f1 = traj + b * (long long)nrow * (p + 1);
    const double *bN = betaN + b * (long long)nrow;
    // negative-state check on rows 0..n0, all state columns
    int neg = 0;
    for (int e = lane; e < (P.n0 + 1) * p; e += 32) {
        const int r = e % (P.n0 + 1), c = e / (P.n0 + 1);
        if (f1[r + (long long)(c + 1) * nrow] < 0.0) neg = 1;
    }
    neg = __any_sync(0xffffffffu, neg);
    double *cq = cellq + (size_t)w * P.n0, *cl = celll + (size_t)w * P.n0;
    for (int i = lane; i < P.n0; i += 32) {
        const int row = P.Lrow - i;
        // transX = "log": the trajectory already holds logs (:1160-1163)
        const double vI = f1[row + (long long)(P.idxI + 1) * nrow], vS = f1[row + (long long)(P.idxS + 1) * nrow];
        const double f = P.transX ? vI : log(vI);
        const double s = P.transX ? vS : log(vS);
        const double be = bN[row];
        cq[i] = be * exp(-f) * exp(s);
        cl[i] = log(be) - f + s;
    }
    __syncwarp();
    double acc = 0.0;
    int nan = 0;
    if (stage_events) {
        for (int e = lane; e < P.E; e += 32) {
            const int c = evCell[e];
            const double v = -2.0 * (evCD[e] * cq[c]) + evY[e] * cl[c];
            nan |= (v != v);
            acc += v;
        }
    } else {
        for (int c = 0; c < P.n0; ++c)
            for (int e = P.evStart[c] + lane; e < P.evStart[c + 1]; e += 32) {
                const double v = -2.0 * (P.evC[e] * P.evD[e] * cq[c]) + P.evY[e] * cl[c];
                nan |= (v != v);
                acc += v;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) {
        int st = 0;
        double r = acc;
        if (neg) {
            r = kSentinel;
            st = ST_NEG_STATE;
        } else if (nan) {
            st = ST_NAN;
            if (nan_guard) r = kSentinel;  // volz_loglik_nh3 (:1058-1113) has no NaN guard: the sum is returned
        }
        loglik[b] = r;
        if (status) status[b] = st;
    }
}

// coal_loglik3 (LNA_functional.cpp:1016-1054) / coal_loglik (SIR_phylodyn.cpp:436-468): Kingman-lambda.
// One warp per trajectory; cell c uses row n0k - c where n0k = first row with time >= t_correct.
__global__ void k_kingman(DevProb P, long long B, const double *traj, const double *lambda, int col, int take_log,
                          double *loglik) {
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * warps + w;
    if (b >= B) return;
    const int nrow = P.nrow;
    const double *f1 = traj + b * (long long)nrow * (P.p + 1);
    int n0k = 0;
    while (n0k < nrow - 1 && f1[n0k] < P.t_correct) n0k++;
    const double lam = lambda[b], llam = log(lam);
    double acc = 0.0;
    const int ncell = n0k < P.n0 + 1 ? n0k : P.n0 + 1;
    for (int c = 0; c < ncell && c < P.n0; ++c) {
        double f = f1[(n0k - c) + (long long)col * nrow];
        if (take_log) f = log(f);
        const double ef = exp(-f), lf = llam - f;
        for (int e = P.evStart[c] + lane; e < P.evStart[c + 1]; e += 32)
            acc += -lam * (P.evC[e] * P.evD[e] * ef) + P.evY[e] * lf;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) loglik[b] = acc;
}

// Independent-DFMA microbenchmark for the FP64 roofline denominator: 8 independent chains per thread, multiplier and
// addend as kernel parameters (constant-bank operands: a DFMA with three distinct REGISTER sources issues at 2/3 rate,
// tools/fp64_microbench.cu), time loop unrolled so that loop overhead does not compete for issue slots.  With 16 resident
// warps per SM sub-partition this reads the pipe's rate: 64 lanes x 2 flop x SM clock per SM.
__global__ void __launch_bounds__(512) k_fp64_peak(int iters, double m, double c, double *sink) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}

}  // namespace lna
