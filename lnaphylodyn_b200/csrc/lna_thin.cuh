// lna_thin.cuh -- volz_thin_sir (R/coal_simulation.R:11-70) on the device: coalescent times of a genealogy simulated by
// thinning against the Volz rate on an SIR trajectory, sm_100a.
//
// The reference's loop is sequential: propose the next event after an Exp(A(A-1)/2 / lower) waiting time, accept it as a
// coalescence with probability lower / thed(time), thed = I / (2 beta S).  On the vignette-scale problems the acceptance
// probability is 1e-3 .. 1e-5, i.e. almost every proposal only advances the clock.  Between two events the proposals form a
// homogeneous Poisson process with independent marks, so ONE WARP per genealogy evaluates 32 consecutive proposals at once:
//   lane j draws proposal number (draw + j) of the genealogy's counter-based stream (Philox4x32-10: waiting time and mark),
//   the arrival times are accumulated in the sequential order of additions (32 shuffles, then a 32-deep add chain, so every
//   lane holds exactly the double the sequential loop would hold), each lane looks up its own grid cell, and a ballot finds
//   the first proposal that is an event (a coalescence or the next sampling time).  Proposals after it are discarded and
//   drawn again in the next round under the new lineage count -- the stream is indexed by proposal number, so the result is
//   bit-identical to the sequential loop consuming the same stream (checked in tests/test_gpu_thin.py).
#pragma once
#include "lna_eslice.cuh"

namespace lna {

struct ThinArgs {
    const double *t, *S, *I, *betaN;  // fine-grid trajectory of item b at + b * stride (stride 0: one trajectory for all items)
    long long stride;
    int n;
    const double *samp_times;  // [ns]
    const int *n_sampled;      // [ns]
    int ns;
    double t_correct, lower;
    unsigned long long seed;
    long long item_offset;     // stream id of item 0 (sharding across GPUs)
    double *coal_times;        // B x max_coal
    int *lineages;             // B x max_coal
    int max_coal;
    int *n_coal, *status;      // [B]; status 1 = a limit was hit (max_coal / max_draws / sampling list exhausted)
    unsigned long long *n_draws;
    unsigned long long max_draws;
};

// proposal number `draw` of genealogy `item`: a -> waiting time, b -> thinning mark.  The top bit of the last counter word
// keeps these streams apart from the resident chains' (k_draw_streams) under the same seed.
__device__ __forceinline__ void thin_draw(unsigned long long seed, unsigned long long item, unsigned long long draw, double &a,
                                          double &b) {
    uint32_t r[4];
    philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)item, (uint32_t)(item >> 32) | 0x80000000u, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    a = u01(r[0], r[1]);
    b = u01(r[2], r[3]);
}

__global__ void k_thin_draws(unsigned long long seed, unsigned long long item, unsigned long long first, long long count,
                             double *a, double *b) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    thin_draw(seed, item, first + (unsigned long long)j, a[j], b[j]);
}

constexpr int kThinWarps = 4;
__global__ void __launch_bounds__(kThinWarps * 32) k_volz_thin_sir(long long B, ThinArgs a) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * kThinWarps + (threadIdx.x >> 5);
    if (b >= B) return;  // whole warps exit together
    const unsigned full = 0xffffffffu;
    const double *t = a.t + b * a.stride, *S = a.S + b * a.stride, *I = a.I + b * a.stride, *bet = a.betaN + b * a.stride;
    const unsigned long long item = (unsigned long long)(a.item_offset + b);
    double *ct = a.coal_times + b * (long long)a.max_coal;
    int *lin = a.lineages + b * (long long)a.max_coal;
    const int n = a.n;
    // i = which.min(ts >= 0) - 1 (1-based) = the last grid point at or before t_correct; every lane scans a stripe
    int first_false = n;
    for (int r = lane; r < n; r += 32)
        if (!(a.t_correct - t[r] >= 0.0)) {
            first_false = r;
            break;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first_false = min(first_false, __shfl_xor_sync(full, first_false, o));
    int i = first_false >= n ? 0 : max(first_false - 1, 0);  // (which.min of an all-TRUE vector is 1)
    double max_samp = a.samp_times[0];
    for (int s = 1; s < a.ns; ++s) max_samp = fmax(max_samp, a.samp_times[s]);
    int curr = 0, active = a.n_sampled[0], ncoal = 0, st = 0;
    double time = a.samp_times[0];
    unsigned long long draw = 0;
    while (time <= max_samp || active > 1) {
        if (active == 1) {
            if (curr + 1 >= a.ns) {
                st = 1;
                break;
            }
            curr += 1;
            active += a.n_sampled[curr];
            time = a.samp_times[curr];
        }
        if (draw >= a.max_draws || ncoal >= a.max_coal) {
            st = 1;
            break;
        }
        const double rate = 0.5 * active * (active - 1) / a.lower;
        double ua, ub;
        thin_draw(a.seed, item, draw + (unsigned long long)lane, ua, ub);
        const double E = -log(ua) / rate;  // rexp(1, rate)
        double e[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) e[q] = __shfl_sync(full, E, q);
        double tj = time;
#pragma unroll
        for (int q = 0; q < 32; ++q) tj = q <= lane ? __dadd_rn(tj, e[q]) : tj;  // time + E_0 + .. + E_lane, left to right
        int ij = i;
        while (tj > a.t_correct - t[ij]) {
            ij -= 1;
            if (ij < 0) {
                ij = 0;
                break;
            }
        }
        const double thed = 1.0 / (bet[ij] * S[ij] / (I[ij] / 2.0));
        const bool samp = curr < a.ns - 1 && tj >= a.samp_times[curr + 1];
        const bool acc = !samp && ub <= a.lower / thed;
        const unsigned ev = __ballot_sync(full, samp || acc);
        if (ev == 0) {
            time = __shfl_sync(full, tj, 31);
            i = __shfl_sync(full, ij, 31);
            draw += 32;
            continue;
        }
        const int f = __ffs(ev) - 1;
        const double tf = __shfl_sync(full, tj, f);
        const bool is_samp = (__ballot_sync(full, samp) >> f) & 1u;
        i = __shfl_sync(full, ij, f);
        draw += (unsigned long long)f + 1;
        if (is_samp) {
            curr += 1;
            active += a.n_sampled[curr];
            time = a.samp_times[curr];
        } else {
            time = fmin(tf, a.t_correct);
            if (lane == 0) {
                ct[ncoal] = time;
                lin[ncoal] = active;
            }
            ncoal += 1;
            active -= 1;
        }
    }
    if (lane == 0) {
        a.n_coal[b] = ncoal;
        a.status[b] = st;
        a.n_draws[b] = draw;
    }
}

}  // namespace lna
