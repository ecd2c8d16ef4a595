// psph_comm.cu — key-range slabs and ghost exchange across the GPUs of one box (NCCL).
#include "psph_internal.cuh"

void psph_comm_destroy(psph_ctx *ctx) { (void) ctx; }
extern "C" int psph_comm_init(psph_ctx *ctx, const void *nccl_unique_id, size_t id_bytes)
{
  (void) nccl_unique_id; (void) id_bytes;
  if (ctx->cfg.nranks == 1) return 0;
  ctx->err = "psph_comm_init: multi-rank not built yet";
  return -4;
}
extern "C" int psph_exchange_ghosts(psph_ctx *ctx) { return ctx->cfg.nranks == 1 ? 0 : -4; } // move_data.cc:53-54
extern "C" int psph_repartition(psph_ctx *ctx) { return ctx->cfg.nranks == 1 ? 0 : -4; }
