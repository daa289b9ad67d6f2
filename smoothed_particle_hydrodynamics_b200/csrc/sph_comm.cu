// sph_comm.cu -- multi-GPU slab decomposition along z (placeholder stubs until
// the halo / migration exchange lands; the symbols exist so the ABI is stable).
#include "sph_internal.h"

int sph_comm_exchange(sphb200_ctx* ctx) { return sph_fail(ctx, SPHB200_E_COMM, "slab exchange not initialised"); }
void sph_comm_free(sphb200_ctx*) {}

extern "C" {

int sphb200_comm_unique_id(void*) { return sph_fail(nullptr, SPHB200_E_COMM, "slab mode not built yet"); }
int sphb200_comm_init(sphb200_ctx* ctx, int, int, const void*, int, int)
{
   return sph_fail(ctx, SPHB200_E_COMM, "slab mode not built yet");
}
int sphb200_get_local_count(const sphb200_ctx* ctx, int* owned, int* ghosts)
{
   if (!ctx)
      return SPHB200_E_INVALID;
   if (owned) *owned = ctx->n_owned;
   if (ghosts) *ghosts = ctx->n_local - ctx->n_owned;
   return SPHB200_OK;
}
int sphb200_upload_slab(sphb200_ctx* ctx, int, const float*, const float*, const float*, const uint32_t*)
{
   return sph_fail(ctx, SPHB200_E_COMM, "slab mode not built yet");
}
int sphb200_download_slab(sphb200_ctx* ctx, int, void*, size_t, uint32_t*, int*)
{
   return sph_fail(ctx, SPHB200_E_COMM, "slab mode not built yet");
}
}
