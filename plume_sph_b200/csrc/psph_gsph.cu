// psph_gsph.cu — Godunov-SPH variant (calc_gradients + Riemann-solver momentum/energy update).
#include "psph_gather.cuh"

extern "C" int psph_calc_gradients(psph_ctx *ctx)
{
  return psph_gsph_gradients(ctx);
}
int psph_gsph_gradients(psph_ctx *ctx) { ctx->err = "GSPH gradients: not built yet"; return -4; }
int psph_gsph_mom_engr(psph_ctx *ctx) { ctx->err = "GSPH mom_engr_update: not built yet"; return -4; }
