// pdefilter.cu -- Helmholtz PDE filter (placeholder until the stencil solver lands)
#include "filter.cuh"
namespace tob {
struct PdeFilterState { int dummy; };
int pdefilter_create(tob_filter_impl* h) { (void)h; TOB_FAIL("PDE filter (type 2) not built yet"); }
int pdefilter_apply(tob_filter_impl* h, const double* x, double* out) { (void)h; (void)x; (void)out; TOB_FAIL("PDE filter not built yet"); }
void pdefilter_destroy(tob_filter_impl* h) { if (h->pde) { delete h->pde; h->pde = nullptr; } }
}  // namespace tob
