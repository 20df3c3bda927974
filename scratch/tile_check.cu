#include "../sihmpc_b200/csrc/tile_kernel.cuh"
cudaError_t chk(const sihmpc_plan_desc *d, const DevPlan &P, sih::TilePlan *tp) { return sih::tile_configure(d, P, tp); }
cudaError_t lch(const sih::TilePlan &tp, const DevPlan &P, const sihmpc_io &io) { return sih::tile_launch(tp, P, io, 8, 1, 148, 0); }
