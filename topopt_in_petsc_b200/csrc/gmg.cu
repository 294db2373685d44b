// gmg.cu -- geometric multigrid V-cycle preconditioner (placeholder until the hierarchy lands)
#include "elast.cuh"
namespace tob {
struct GmgHierarchy { int dummy; };
int gmg_setup(tob_elast_impl* h) { (void)h; TOB_FAIL("multigrid preconditioner (pc=1) not built yet"); }
int gmg_update(tob_elast_impl* h) { (void)h; return 0; }
int gmg_vcycle(tob_elast_impl* h, const double* r, double* z) { (void)h; (void)r; (void)z; TOB_FAIL("multigrid preconditioner not built yet"); }
void gmg_destroy(tob_elast_impl* h) { if (h->mg) { delete h->mg; h->mg = nullptr; } }
}  // namespace tob
