// lna_comm.inl -- the one exchange of the path behind the C ABI: an NCCL all-gather of the per-chain
// summary rows over NVLink / NVSwitch, and a single-process driver for the resident chains of several GPUs (lna_multi).
// Included at the end of lna_capi.cu.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: torch's bundled copy when the process has loaded it, the system one
// otherwise), so the library itself has no link-time dependency on it and a one-GPU session never touches it.
#include <dlfcn.h>

namespace {

// the few NCCL entry points used, declared here so that no particular nccl.h version is needed at build time
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId_;
enum { kNcclSuccess = 0, kNcclDouble = 8 };
struct NcclApi {
    void *so = nullptr;
    int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("LNA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.so) break;
        }
        if (!api.so) {
            api.err = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "?") + "); set LNA_NCCL_LIB";
            return;
        }
#define LNA_NCCL_SYM(field, sym)                                             \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.so, sym)); \
    if (!api.field) api.err = std::string("NCCL symbol missing: ") + sym;
        LNA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        LNA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        LNA_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        LNA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        LNA_NCCL_SYM(AllGather, "ncclAllGather")
        LNA_NCCL_SYM(Broadcast, "ncclBroadcast")
        LNA_NCCL_SYM(GroupStart, "ncclGroupStart")
        LNA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        LNA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        LNA_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef LNA_NCCL_SYM
    });
    return &api;
}

}  // namespace

#define NCK(call)                                                                                                   \
    do {                                                                                                            \
        const int r_ = (call);                                                                                      \
        if (r_ != kNcclSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(r_)); \
    } while (0)

struct lna_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    bool owned = true;
};

extern "C" int lna_comm_available(void) {
    NcclApi *a = nccl_api();
    if (!a->so || !a->err.empty()) {
        fail("lna_comm: %s", a->err.empty() ? "NCCL unavailable" : a->err.c_str());
        return 0;
    }
    int v = 0;
    a->GetVersion(&v);
    return v > 0 ? v : 1;
}

extern "C" int lna_comm_unique_id(void *id128) {
    if (!id128) return fail("lna_comm_unique_id: null buffer");
    if (!lna_comm_available()) return -1;
    ncclUniqueId_ id;
    NCK(nccl_api()->GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return 0;
}

extern "C" lna_comm *lna_comm_create(int nranks, int rank, const void *id128, int device) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) {
        fail("lna_comm_create: bad arguments");
        return nullptr;
    }
    if (!lna_comm_available()) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) {
        fail("lna_comm_create: cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    ncclUniqueId_ id;
    std::memcpy(&id, id128, sizeof id);
    lna_comm *c = new lna_comm();
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    const int r = nccl_api()->CommInitRank(&c->comm, nranks, id, rank);
    if (r != kNcclSuccess) {
        fail("lna_comm_create: ncclCommInitRank: %s", nccl_api()->GetErrorString(r));
        delete c;
        return nullptr;
    }
    return c;
}

extern "C" void lna_comm_destroy(lna_comm *c) {
    if (!c) return;
    if (c->comm && c->owned) {
        cudaSetDevice(c->device);
        nccl_api()->CommDestroy(c->comm);
    }
    delete c;
}

extern "C" int lna_comm_rank(const lna_comm *c) { return c ? c->rank : -1; }
extern "C" int lna_comm_size(const lna_comm *c) { return c ? c->nranks : 0; }

// all-gather of contiguous shards of doubles: rank r owns counts[r] doubles at offset sum(counts[0..r-1]) of `recv`.
// Equal shards: ONE ncclAllGather (in place when send already sits at its offset); ragged shards (lna_shard sizes differ by
// one item): one grouped launch of nranks broadcasts.  No size exchange, no host synchronisation.
static int comm_allgather_v(lna_comm *c, const double *send, double *recv, const int64_t *counts, cudaStream_t st) {
    NcclApi *a = nccl_api();
    bool equal = true;
    for (int r = 1; r < c->nranks; ++r) equal = equal && counts[r] == counts[0];
    if (equal) {
        NCK(a->AllGather(send, recv, (size_t)counts[0], kNcclDouble, c->comm, st));
        return 0;
    }
    NCK(a->GroupStart());
    int64_t off = 0;
    for (int r = 0; r < c->nranks; ++r) {
        const int rc = a->Broadcast(r == c->rank ? send : recv + off, recv + off, (size_t)counts[r], kNcclDouble, r, c->comm, st);
        if (rc != kNcclSuccess) {
            a->GroupEnd();
            return fail("ncclBroadcast: %s", a->GetErrorString(rc));
        }
        off += counts[r];
    }
    NCK(a->GroupEnd());
    return 0;
}

extern "C" int lna_comm_allgather(lna_comm *c, const double *send_dev, double *recv_dev, const int64_t *counts, void *stream) {
    if (!c || !send_dev || !recv_dev || !counts) return fail("lna_comm_allgather: null argument");
    cudaSetDevice(c->device);
    return comm_allgather_v(c, send_dev, recv_dev, counts, (cudaStream_t)stream);
}

// The summary rows (coalLog, logOrigin, n_full_evals, status) of ALL chains of the job, in global chain order, into
// out_dev (B_total x 4 doubles on this rank's device): this rank contributes the rows of its resident batch, whose size
// must be its lna_shard of B_total.  Asynchronous on the problem's stream (after joining the batch's stream groups).
extern "C" int lna_chains_allgather_summaries(lna_chains *ch, lna_comm *c, int64_t B_total, double *out_dev) {
    if (!ch || !c || !out_dev) return fail("lna_chains_allgather_summaries: null argument");
    std::vector<int64_t> counts(c->nranks);
    for (int r = 0; r < c->nranks; ++r) {
        int64_t b, e;
        lna_shard(B_total, r, c->nranks, &b, &e);
        counts[r] = 4 * (e - b);
    }
    if (counts[c->rank] != 4 * ch->B)
        return fail("lna_chains_allgather_summaries: this rank holds %lld chains, its shard of %lld over %d ranks is %lld",
                    ch->B, (long long)B_total, c->nranks, (long long)(counts[c->rank] / 4));
    cudaSetDevice(ch->pb->device);
    if (chains_join_now(ch)) return -1;
    return comm_allgather_v(c, ch->summary, out_dev, counts.data(), ch->pb->stream);
}

// ------------------------------------------------------------------------------------------------
// lna_multi: ONE process (e.g. an R session behind the Rcpp layer) driving the resident chains of several GPUs.
// Every step call only enqueues work (CUDA graph launches on the devices' streams), so one host thread keeps all devices
// busy; chains are sharded contiguously (lna_shard), the Philox streams are keyed by the GLOBAL chain index, so the result
// does not depend on the number of devices.
// ------------------------------------------------------------------------------------------------
struct lna_multi {
    int n = 0;
    long long B = 0;
    std::vector<int> device;
    std::vector<lna_problem *> pb;
    std::vector<lna_chains *> ch;
    std::vector<long long> off;       // off[n] = B
    std::vector<ncclComm_t> comm;     // ncclCommInitAll over the devices (empty when a device is listed twice / NCCL absent)
    std::vector<double *> gathered;   // per device: B x 4 summaries of all chains
};

extern "C" void lna_multi_destroy(lna_multi *m) {
    if (!m) return;
    for (int i = 0; i < (int)m->ch.size(); ++i) {
        if (m->ch[i]) lna_chains_destroy(m->ch[i]);
        if (i < (int)m->gathered.size() && m->gathered[i]) {
            cudaSetDevice(m->device[i]);
            cudaFree(m->gathered[i]);
        }
    }
    for (int i = 0; i < (int)m->comm.size(); ++i)
        if (m->comm[i]) {
            cudaSetDevice(m->device[i]);
            nccl_api()->CommDestroy(m->comm[i]);
        }
    for (lna_problem *p : m->pb)
        if (p) lna_problem_destroy(p);
    delete m;
}

extern "C" lna_multi *lna_multi_create(const lna_cfg *cfg, const lna_init *init, int ndev, const int *devices, int64_t B_total) {
    if (!cfg || ndev < 1 || !devices || B_total < ndev) {
        fail("lna_multi_create: bad arguments (need at least one chain per device)");
        return nullptr;
    }
    lna_multi *m = new lna_multi();
    m->n = ndev;
    m->B = B_total;
    m->device.assign(devices, devices + ndev);
    m->off.assign(ndev + 1, 0);
    for (int i = 0; i < ndev; ++i) {
        int64_t b, e;
        lna_shard(B_total, i, ndev, &b, &e);
        m->off[i] = b;
        m->off[i + 1] = e;
        lna_problem *p = lna_problem_create(cfg, init, devices[i]);
        m->pb.push_back(p);
        lna_chains *c = p ? lna_chains_create(p, e - b) : nullptr;
        m->ch.push_back(c);
        if (!p || !c) {
            lna_multi_destroy(m);
            return nullptr;
        }
    }
    return m;
}

extern "C" int lna_multi_devices(const lna_multi *m) { return m ? m->n : 0; }
extern "C" lna_chains *lna_multi_chains(lna_multi *m, int i) { return (m && i >= 0 && i < m->n) ? m->ch[i] : nullptr; }
extern "C" lna_problem *lna_multi_problem(lna_multi *m, int i) { return (m && i >= 0 && i < m->n) ? m->pb[i] : nullptr; }

extern "C" int lna_multi_init(lna_multi *m, const double *par, const double *OriginTraj) {
    if (!m || !par || !OriginTraj) return fail("lna_multi_init: null argument");
    const DevProb &P = m->pb[0]->dp;
    const size_t npar = 2 + P.npt, kp = (size_t)P.k * P.p;
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_init(m->ch[i], par + (size_t)m->off[i] * npar, OriginTraj + (size_t)m->off[i] * kp)) return -1;
    return 0;
}

extern "C" int lna_multi_set_free_running(lna_multi *m, int on) {
    if (!m) return fail("lna_multi_set_free_running: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_set_free_running(m->ch[i], on)) return -1;
    return 0;
}

extern "C" int lna_multi_step_rng(lna_multi *m, const lna_mcmc_opts *o, uint64_t seed, uint64_t iteration, int64_t chain_offset) {
    if (!m || !o) return fail("lna_multi_step_rng: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_step_rng(m->ch[i], o, seed, iteration, chain_offset + m->off[i])) return -1;
    return 0;
}

extern "C" int lna_multi_step_mcmc2_rng(lna_multi *m, const lna_mcmc2_opts *o, uint64_t seed, uint64_t iteration,
                                        int64_t chain_offset) {
    if (!m || !o) return fail("lna_multi_step_mcmc2_rng: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_step_mcmc2_rng(m->ch[i], o, seed, iteration, chain_offset + m->off[i])) return -1;
    return 0;
}

extern "C" int lna_multi_set_joint_transform(lna_multi *m, int d, const double *T) {
    if (!m || !T) return fail("lna_multi_set_joint_transform: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_set_joint_transform(m->ch[i], d, T + (size_t)m->off[i] * d * d)) return -1;
    return 0;
}

extern "C" int lna_multi_sync(lna_multi *m) {
    if (!m) return fail("lna_multi_sync: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_sync(m->ch[i])) return -1;
    return 0;
}

extern "C" int lna_multi_get(lna_multi *m, double *par, double *OriginTraj, double *LatentTraj, double *coalLog,
                             double *logOrigin, int32_t *counters) {
    if (!m) return fail("lna_multi_get: null argument");
    const DevProb &P = m->pb[0]->dp;
    const size_t npar = 2 + P.npt, kp = (size_t)P.k * P.p, nO = (size_t)P.nrow * (P.p + 1);
    for (int i = 0; i < m->n; ++i) {
        const size_t o = (size_t)m->off[i];
        if (lna_chains_get(m->ch[i], par ? par + o * npar : nullptr, OriginTraj ? OriginTraj + o * kp : nullptr,
                           LatentTraj ? LatentTraj + o * nO : nullptr, coalLog ? coalLog + o : nullptr,
                           logOrigin ? logOrigin + o : nullptr, counters ? counters + o * 4 : nullptr))
            return -1;
    }
    return 0;
}

extern "C" int lna_multi_history_begin(lna_multi *m, int64_t niter) {
    if (!m) return fail("lna_multi_history_begin: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_history_begin(m->ch[i], niter)) return -1;
    return 0;
}
extern "C" int lna_multi_history_record(lna_multi *m, int64_t it) {
    if (!m) return fail("lna_multi_history_record: null argument");
    for (int i = 0; i < m->n; ++i)
        if (lna_chains_history_record(m->ch[i], it)) return -1;
    return 0;
}
extern "C" int lna_multi_history_get(lna_multi *m, int64_t niter, double *par, double *l, double *l1) {
    if (!m) return fail("lna_multi_history_get: null argument");
    const size_t npar = 2 + m->pb[0]->dp.npt;
    for (int i = 0; i < m->n; ++i) {
        const size_t o = (size_t)m->off[i];
        if (lna_chains_history_get(m->ch[i], par ? par + o * niter * npar : nullptr, l ? l + o * niter : nullptr,
                                   l1 ? l1 + o * niter : nullptr))
            return -1;
    }
    return 0;
}

// All devices end up holding the summary rows of all B chains (ncclAllGather over NVLink inside one grouped call); the
// host copy comes from the first device.  With a device listed twice (tests on a one-GPU box) or without NCCL, `use_nccl`
// must be 0 and the rows are collected with device-to-host copies instead.
extern "C" int lna_multi_gather_summaries(lna_multi *m, int use_nccl, double *out_host) {
    if (!m) return fail("lna_multi_gather_summaries: null argument");
    const size_t nB = (size_t)m->B;
    if (!use_nccl) {
        if (!out_host) return 0;
        for (int i = 0; i < m->n; ++i) {
            lna_chains *c = m->ch[i];
            cudaSetDevice(m->device[i]);
            if (chains_join_now(c)) return -1;
            CK(cudaMemcpyAsync(out_host + 4 * (size_t)m->off[i], c->summary, sizeof(double) * 4 * (size_t)c->B, cudaMemcpyDeviceToHost,
                               c->pb->stream));
        }
        for (int i = 0; i < m->n; ++i) {
            cudaSetDevice(m->device[i]);
            CK(cudaStreamSynchronize(m->ch[i]->pb->stream));
        }
        return 0;
    }
    if (!lna_comm_available()) return -1;
    NcclApi *a = nccl_api();
    if (m->comm.empty()) {
        std::set<int> uniq(m->device.begin(), m->device.end());
        if ((int)uniq.size() != m->n) return fail("lna_multi_gather_summaries: NCCL needs distinct devices");
        m->comm.assign(m->n, nullptr);
        const int r = a->CommInitAll(m->comm.data(), m->n, m->device.data());
        if (r != kNcclSuccess) {
            m->comm.clear();
            return fail("ncclCommInitAll: %s", a->GetErrorString(r));
        }
        m->gathered.assign(m->n, nullptr);
        for (int i = 0; i < m->n; ++i) {
            cudaSetDevice(m->device[i]);
            CK(cudaMalloc(&m->gathered[i], sizeof(double) * 4 * nB));
        }
    }
    std::vector<int64_t> counts(m->n);
    for (int i = 0; i < m->n; ++i) counts[i] = 4 * (m->off[i + 1] - m->off[i]);
    for (int i = 0; i < m->n; ++i) {
        cudaSetDevice(m->device[i]);
        if (chains_join_now(m->ch[i])) return -1;
    }
    NCK(a->GroupStart());
    for (int i = 0; i < m->n; ++i) {
        lna_comm c;
        c.comm = m->comm[i];
        c.nranks = m->n;
        c.rank = i;
        c.device = m->device[i];
        cudaSetDevice(m->device[i]);
        if (comm_allgather_v(&c, m->ch[i]->summary, m->gathered[i], counts.data(), m->ch[i]->pb->stream)) {
            a->GroupEnd();
            return -1;
        }
    }
    NCK(a->GroupEnd());
    if (out_host) {
        cudaSetDevice(m->device[0]);
        CK(cudaMemcpyAsync(out_host, m->gathered[0], sizeof(double) * 4 * nB, cudaMemcpyDeviceToHost, m->ch[0]->pb->stream));
        CK(cudaStreamSynchronize(m->ch[0]->pb->stream));
    }
    return 0;
}

extern "C" const double *lna_multi_gathered_dev(lna_multi *m, int i) {
    return (m && i >= 0 && i < (int)m->gathered.size()) ? m->gathered[i] : nullptr;
}
