// ESlice / chains entry points (stubs until the kernels land)
#define LNA_STUB(name, ...) extern "C" int name(__VA_ARGS__) { return fail(#name ": not implemented yet"); }
LNA_STUB(lna_param_slice_update, lna_problem *, int64_t, const double *, const double *, const double *, double *)
LNA_STUB(lna_param_slice_update_all2, lna_problem *, int64_t, const double *, const double *, const double *,
         const int32_t *, const double *, double *)
LNA_STUB(lna_eslice_par_general, lna_problem *, int64_t, const double *, const double *, const double *,
         const int32_t *, const double *, const double *, const double *, int, double *, double *, double *, double *,
         double *, double *, double *, double *, int32_t *, int32_t *)
LNA_STUB(lna_eslice_general_nc, lna_problem *, int64_t, const double *, const double *, const double *, const double *,
         const double *, const double *, const double *, const double *, double *, double *, double *, double *,
         int32_t *, int32_t *)
LNA_STUB(lna_eslice_change_points, lna_problem *, int64_t, const double *, const double *, const double *,
         const double *, const double *, const double *, int, double *, double *, double *, double *, double *,
         double *, double *, double *, int32_t *, int32_t *)
struct lna_chains { lna_problem *pb; };
extern "C" lna_chains *lna_chains_create(lna_problem *, int64_t) { fail("lna_chains_create: not implemented yet"); return nullptr; }
extern "C" void lna_chains_destroy(lna_chains *) {}
LNA_STUB(lna_chains_init, lna_chains *, const double *, const double *)
LNA_STUB(lna_chains_step, lna_chains *, const lna_mcmc_opts *, const double *, const double *)
LNA_STUB(lna_chains_step_dev, lna_chains *, const lna_mcmc_opts *, const double *, const double *)
LNA_STUB(lna_chains_get, lna_chains *, double *, double *, double *, double *, double *, int32_t *)
extern "C" const double *lna_chains_summary_dev(lna_chains *) { return nullptr; }
