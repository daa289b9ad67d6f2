// sph_grid.cu -- voxel binning, counting sort into cell order, cell tables, gather.
//
// Replaces clearGrid + voxelizeParticles (sph.cpp:429-481): instead of
// QList::push_back per particle, particles get a cell key and are counting-sorted
// by it every step; an exclusive-scan table cell_start[c] .. cell_start[c+1]
// replaces the lists.  The order inside a cell is ascending particle index == the
// reference's push_back order (sph.cpp:476-480): the histogram atomics hand out
// arbitrary slots, and the gather pass re-ranks the (few) members of each cell.
//   k_cell_keys   : key + histogram atomic (returns the particle's slot in its cell)
//   scan          : cell_start = exclusive sum of the histogram (cub::DeviceScan)
//   k_scatter     : particle -> cell_start[key] + slot
//   k_rank_gather : rank inside the cell by particle index (global id in slab mode),
//                   final index order + positions gathered into cell order
// ~100 B of traffic per particle in four streaming passes instead of a 3-pass radix
// sort of (key, index) pairs plus a gather (0.73 -> 0.4 ms at 16.7M particles).
#include <cub/cub.cuh>

#include <cstdlib>

#include "sph_math.cuh"

namespace
{

constexpr int kThreads = 256;

// SAMPLED mode: key = voxel id (sph.cpp:443-473, 1151-1154).
// FULL mode:    key = fine-cell id, fine = 2*voxel + (orientation > h) per axis
//               (the octant rule of sph.cpp:504-515), x fastest.
// Slab mode:    the z voxel is clamped against the GLOBAL box exactly like the
//               reference, then shifted into the rank's local grid; FREE slots get
//               the sentinel key `cells` and therefore sort behind every particle.
template <bool FINE>
__global__ void __launch_bounds__(kThreads) k_cell_keys(DevParams P, int cells, const float4* __restrict__ pos4,
                                                         uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ cell_count,
                                                         uint32_t* __restrict__ cell_slot,
                                                         int* __restrict__ voxel_id_out, uint32_t* __restrict__ xkey)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (P.slab)
   {
      // FREE slots all fall into the sentinel cell: one atomic per block, not per slot
      const bool is_free = i < P.n && P.slot_state[i] == SLOT_FREE;
      const unsigned f = sph_block_append(&cell_count[cells], is_free);
      if (is_free)
      {
         keys[i] = (uint32_t)cells;
         cell_slot[i] = f;
         return;
      }
   }
   if (i >= P.n)
      return;
   float4 p = pos4[i];
   int vx = sph_voxel_coord(p.x, P.h_times2_inv, P.gx);
   int vy = sph_voxel_coord(p.y, P.h_times2_inv, P.gy);
   int vz = sph_voxel_coord(p.z, P.h_times2_inv, P.slab ? P.gz_global : P.gz);
   int lz = vz;
   if (P.slab)
      lz = min(max(vz - P.vz_offset, 0), P.gz - 1);   // only a particle that jumped >1 slab is clamped
   uint32_t key;
   if (FINE)
   {
      int cx = 2 * vx + sph_upper_half(p.x, vx, P.h_times2, P.h);
      int cy = 2 * vy + sph_upper_half(p.y, vy, P.h_times2, P.h);
      int cz = 2 * lz + sph_upper_half(p.z, vz, P.h_times2, P.h);
      key = (uint32_t)((cz * P.fy + cy) * P.fx + cx);
   }
   else
      key = (uint32_t)sph_voxel_id(vx, vy, lz, P.gx, P.gy);
   keys[i] = key;
   if (FINE)
      xkey[i] = sph_x_order_key(p.x);
   if (voxel_id_out)
      voxel_id_out[i] = sph_voxel_id(vx, vy, vz, P.gx, P.gy);   // global voxel id, as the reference numbers it
   cell_slot[i] = atomicAdd(&cell_count[key], 1u);
}

// counting-sort scatter: particle i goes to position cell_start[key] + slot.  The
// slot order inside a cell is whatever the atomics produced; `pair` carries the value
// the members are ranked by afterwards: (x order key, id) as one 64-bit word -- id =
// particle index, or global id in slab mode; the x key is 0 in sampled mode.
__global__ void __launch_bounds__(kThreads) k_scatter(int n, const uint32_t* __restrict__ keys,
                                                       const uint32_t* __restrict__ cell_slot,
                                                       const uint32_t* __restrict__ cell_start,
                                                       const uint32_t* __restrict__ gid,
                                                       const uint32_t* __restrict__ xkey,
                                                       uint32_t* __restrict__ keys_sorted,
                                                       unsigned long long* __restrict__ pair,
                                                       uint32_t* __restrict__ tmp_idx)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   uint32_t key = keys[i];
   uint32_t p = __ldg(&cell_start[key]) + cell_slot[i];
   keys_sorted[p] = key;
   const unsigned long long hi = xkey ? (unsigned long long)xkey[i] << 32 : 0ull;
   if (gid)
   {
      pair[p] = hi | gid[i];
      tmp_idx[p] = (uint32_t)i;
   }
   else
      pair[p] = hi | (uint32_t)i;
}

// One thread per sorted position: rank of its particle among the members of its cell
// (cells hold ~8 particles and the members sit in consecutive, L1-resident words), then
// the final index order and -- FULL mode -- the position gathered into cell order.
// Rank = ascending (x, id) in FULL mode (every x-run of the sweeps is then ascending in x;
// oracle_order_cells_by_x), ascending id = the reference's push_back order in sampled mode
// (sph.cpp:476-480): one 64-bit comparison either way.
// Free slots (slab mode, sentinel cell `cells`) keep their arbitrary order.
template <bool GATHER>
__global__ void __launch_bounds__(kThreads) k_rank_gather(int n, int cells, const uint32_t* __restrict__ keys_sorted,
                                                           const uint32_t* __restrict__ cell_start,
                                                           const unsigned long long* __restrict__ pair,
                                                           const uint32_t* __restrict__ tmp_idx,
                                                           const float4* __restrict__ pos4,
                                                           uint32_t* __restrict__ idx_sorted,
                                                           float4* __restrict__ s_pos4)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= n)
      return;
   const uint32_t c = keys_sorted[k];
   const unsigned long long mine = pair[k];
   const uint32_t me = tmp_idx ? tmp_idx[k] : (uint32_t)mine;
   if ((int)c >= cells)
   {
      idx_sorted[k] = me;
      return;
   }
   const int s = (int)__ldg(&cell_start[c]), e = (int)__ldg(&cell_start[c + 1]);
   int rank = 0;
   for (int a = s; a < e; a++)
      rank += (pair[a] < mine) ? 1 : 0;
   idx_sorted[s + rank] = me;
   if (GATHER)
      s_pos4[s + rank] = __ldg(&pos4[me]);
}

__global__ void __launch_bounds__(kThreads) k_iota(uint32_t* a, int n)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      a[i] = (uint32_t)i;
}

// particles into cell order: one 16-byte gather per particle
__global__ void __launch_bounds__(kThreads) k_gather_pos(DevParams P, const uint32_t* __restrict__ idx_sorted,
                                                          const float4* __restrict__ pos4,
                                                          float4* __restrict__ s_pos4)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k < sph_live_count(P))
      s_pos4[k] = __ldg(&pos4[idx_sorted[k]]);
}

// Slab mode: a rank's slot order is arbitrary (arrivals land in free slots), so
// the stable sort alone does not give the canonical in-cell order.  Re-rank the
// members of every cell by GLOBAL particle id (== the single-GPU particle index
// == the reference's push_back order, sph.cpp:476-480).  One thread per cell;
// cells hold ~10 particles, so a rank-by-counting pass is cheap.
__global__ void __launch_bounds__(kThreads) k_fix_cell_order(int cells, const uint32_t* __restrict__ cell_start,
                                                              const uint32_t* __restrict__ idx_sorted,
                                                              const uint32_t* __restrict__ gid,
                                                              uint32_t* __restrict__ idx_fixed)
{
   int c = blockIdx.x * blockDim.x + threadIdx.x;
   if (c >= cells)
      return;
   int s = (int)cell_start[c], e = (int)cell_start[c + 1];
   if (s == e)
      return;
   // most cells are already in order (only arrivals of the exchange sit in arbitrary
   // slots): one linear pass decides, and copies
   bool sorted = true;
   uint32_t prev = 0;
   for (int a = s; a < e; a++)
   {
      uint32_t ia = idx_sorted[a];
      uint32_t ga = gid[ia];
      sorted = sorted && (a == s || ga > prev);
      prev = ga;
      idx_fixed[a] = ia;
   }
   if (sorted)
      return;
   for (int a = s; a < e; a++)
   {
      uint32_t ia = idx_sorted[a];
      uint32_t ga = gid[ia];
      int rank = 0;
      for (int b = s; b < e; b++)
         rank += (gid[idx_sorted[b]] < ga) ? 1 : 0;
      idx_fixed[s + rank] = ia;
   }
}

__global__ void k_reset_scalars(StepScalars* s)
{
   s->e_kin = 0.0;
   s->e_pot = 0.0;
   s->nbr_total = 0ull;
   s->nbr_max = -1;
   s->nbr_min = 0x7fffffff;
   s->overflow = 0;
   s->finish_ticket = 0u;
}

// second and third stage of the deterministic energy reduction.  Up to 128 blocks each sum a
// fixed, contiguous share of the per-block partials (in a fixed order) and leave one value
// each behind the partials; the block that delivers last (ticket) adds those up in block
// order.  Same result on every run, and the 2 MB of partials of a 16.7 M-particle step are
// read by many SMs instead of one (30 -> ~4 us).
constexpr int kFinishThreads = 256;

__device__ __forceinline__ void finish_block_sum(double& ek, double& ep, double* sk, double* sp)
{
   sk[threadIdx.x] = ek;
   sp[threadIdx.x] = ep;
   __syncthreads();
   for (int o = kFinishThreads / 2; o > 0; o >>= 1)
   {
      if (threadIdx.x < o)
      {
         sk[threadIdx.x] += sk[threadIdx.x + o];
         sp[threadIdx.x] += sp[threadIdx.x + o];
      }
      __syncthreads();
   }
   ek = sk[0];
   ep = sp[0];
   __syncthreads();
}

__global__ void __launch_bounds__(kFinishThreads) k_finish_scalars(double* __restrict__ partials, int blocks,
                                                                   StepScalars* s)
{
   __shared__ double sk[kFinishThreads], sp[kFinishThreads];
   __shared__ bool last;
   double* stage2 = partials + 2 * (size_t)blocks;
   const int per = (blocks + (int)gridDim.x - 1) / (int)gridDim.x;
   const int b0 = (int)blockIdx.x * per, b1 = min(b0 + per, blocks);
   double ek = 0.0, ep = 0.0;
   for (int b = b0 + (int)threadIdx.x; b < b1; b += kFinishThreads)
   {
      ek += partials[2 * b];
      ep += partials[2 * b + 1];
   }
   finish_block_sum(ek, ep, sk, sp);
   if (threadIdx.x == 0)
   {
      stage2[2 * blockIdx.x] = ek;
      stage2[2 * blockIdx.x + 1] = ep;
      __threadfence();
      last = atomicAdd(&s->finish_ticket, 1u) == gridDim.x - 1;
   }
   __syncthreads();
   if (!last)
      return;
   __threadfence();
   ek = ep = 0.0;
   if (threadIdx.x < gridDim.x)
   {
      ek = ((volatile double*)stage2)[2 * threadIdx.x];
      ep = ((volatile double*)stage2)[2 * threadIdx.x + 1];
   }
   finish_block_sum(ek, ep, sk, sp);
   if (threadIdx.x == 0)
   {
      s->e_kin = ek;
      s->e_pot = ep;
      s->finish_ticket = 0u;
   }
}

int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

}  // namespace

int sph_reset_scalars(sphb200_ctx* ctx)
{
   k_reset_scalars<<<1, 1, 0, ctx->stream>>>(ctx->d_scalars);
   ctx->launches++;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   return SPHB200_OK;
}

int sph_finish_scalars(sphb200_ctx* ctx, int blocks)
{
   // the partials array has room for `blocks` pairs + 256 second-stage pairs (sphb200_create)
   const int grid = blocks >= 8192 ? 128 : blocks >= 1024 ? 16 : 1;
   k_finish_scalars<<<grid, kFinishThreads, 0, ctx->stream>>>(ctx->d_block_partials, blocks, ctx->d_scalars);
   ctx->launches++;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   return SPHB200_OK;
}

// keys -> histogram -> exclusive scan -> stable radix sort of (key, index) ->
// (slab) in-cell order by global id -> (FULL) gather positions into cell order.
// Afterwards ctx->idx_order is the particle order every later kernel uses.
int sph_bin_and_sort(sphb200_ctx* ctx, bool fine)
{
   DevParams P = sph_dev_params(ctx);
   const int n = ctx->n_local;              // slab mode: the slot capacity
   const int cells = fine ? ctx->cells_fine : ctx->cells_voxel;
   const int table = cells + 2;             // [cells] = live count, [cells+1] = slots (slab sentinel cell)
   cudaStream_t st = ctx->stream;
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->cell_count, 0, sizeof(uint32_t) * (size_t)table, st));
   if (n > 0)
   {
      if (fine)
         k_cell_keys<true><<<blocks_for(n), kThreads, 0, st>>>(P, cells, ctx->pos4, ctx->keys, ctx->cell_count,
                                                               ctx->cell_slot, ctx->voxel_id, ctx->xkey);
      else
         k_cell_keys<false><<<blocks_for(n), kThreads, 0, st>>>(P, cells, ctx->pos4, ctx->keys, ctx->cell_count,
                                                                ctx->cell_slot, ctx->voxel_id, nullptr);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
      ctx->voxel_ids_valid = true;
   }
   size_t temp = ctx->cub_temp_bytes;
   SPH_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_temp, temp, ctx->cell_count, ctx->cell_start, table,
                                                     st));
   ctx->launches += 2;
   ctx->idx_order = ctx->idx_sorted;
   if (n > 0 && (fine || !ctx->use_radix_sort))   // (the radix-sort A/B path only knows the index order of sampled mode)
   {
      const uint32_t* gid = ctx->comm ? ctx->gid : nullptr;
      k_scatter<<<blocks_for(n), kThreads, 0, st>>>(n, ctx->keys, ctx->cell_slot, ctx->cell_start, gid,
                                                    fine ? ctx->xkey : nullptr, ctx->keys_sorted, ctx->tmp_pair,
                                                    ctx->tmp_idx);
      if (fine)
         k_rank_gather<true><<<blocks_for(n), kThreads, 0, st>>>(n, cells, ctx->keys_sorted, ctx->cell_start,
                                                                 ctx->tmp_pair, gid ? ctx->tmp_idx : nullptr,
                                                                 ctx->pos4, ctx->idx_sorted, ctx->s_pos4);
      else
         k_rank_gather<false><<<blocks_for(n), kThreads, 0, st>>>(n, cells, ctx->keys_sorted, ctx->cell_start,
                                                                  ctx->tmp_pair, gid ? ctx->tmp_idx : nullptr,
                                                                  ctx->pos4, ctx->idx_sorted, ctx->s_pos4);
      ctx->launches += 2;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   else if (n > 0)
   {
      // A/B path (SPHB200_RADIX_SORT=1): stable radix sort of (key, index) from index order
      int bits = 1;
      while (bits < 32 && (1ll << bits) < (long long)cells + (ctx->comm ? 1 : 0))
         bits++;
      temp = ctx->cub_temp_bytes;
      SPH_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_temp, temp, ctx->keys, ctx->keys_sorted,
                                                          ctx->idx_iota, ctx->idx_sorted, n, 0, bits, st));
      ctx->launches += 1 + 2 * ((bits + 7) / 8);
      if (ctx->comm)
      {
         k_fix_cell_order<<<blocks_for(cells), kThreads, 0, st>>>(cells, ctx->cell_start, ctx->idx_sorted, ctx->gid,
                                                                  ctx->idx_fixed);
         ctx->launches++;
         SPH_CUDA_CHECK(ctx, cudaGetLastError());
         ctx->idx_order = ctx->idx_fixed;
      }
      if (fine)
      {
         k_gather_pos<<<blocks_for(n), kThreads, 0, st>>>(P, ctx->idx_order, ctx->pos4, ctx->s_pos4);
         ctx->launches++;
         SPH_CUDA_CHECK(ctx, cudaGetLastError());
      }
   }
   return SPHB200_OK;
}

// cell tables + CUB scratch for the current grid (re-done when a context becomes a slab)
int sph_grid_alloc(sphb200_ctx* ctx)
{
   size_t table = (size_t)ctx->cells_alloc + 2;
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->cell_count, sizeof(uint32_t) * table));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->cell_start, sizeof(uint32_t) * table));
   size_t t_scan = 0, t_sort = 0;
   cub::DeviceScan::ExclusiveSum(nullptr, t_scan, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)table);
   cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                   (uint32_t*)nullptr, ctx->capacity > 0 ? ctx->capacity : 1, 0, 32);
   ctx->cub_temp_bytes = (t_scan > t_sort ? t_scan : t_sort) + 256;
   SPH_CUDA_CHECK(ctx, cudaMalloc(&ctx->cub_temp, ctx->cub_temp_bytes));
   size_t slots = (size_t)(ctx->capacity > 0 ? ctx->capacity : 1);
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->cell_slot, sizeof(uint32_t) * slots));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->tmp_pair, sizeof(unsigned long long) * slots));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->tmp_idx, sizeof(uint32_t) * slots));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->xkey, sizeof(uint32_t) * slots));
   const char* radix = getenv("SPHB200_RADIX_SORT");
   ctx->use_radix_sort = radix && radix[0] == '1';
   return SPHB200_OK;
}

void sph_grid_free(sphb200_ctx* ctx)
{
   if (ctx->cell_count) cudaFree(ctx->cell_count);
   if (ctx->cell_start) cudaFree(ctx->cell_start);
   if (ctx->cub_temp) cudaFree(ctx->cub_temp);
   if (ctx->cell_slot) cudaFree(ctx->cell_slot);
   if (ctx->tmp_pair) cudaFree(ctx->tmp_pair);
   if (ctx->tmp_idx) cudaFree(ctx->tmp_idx);
   if (ctx->xkey) cudaFree(ctx->xkey);
   ctx->cell_count = ctx->cell_start = nullptr;
   ctx->cell_slot = ctx->tmp_idx = ctx->xkey = nullptr;
   ctx->tmp_pair = nullptr;
   ctx->cub_temp = nullptr;
}

// scratch sizing + iota; called once from sphb200_create
int sph_grid_setup(sphb200_ctx* ctx)
{
   int rc = sph_grid_alloc(ctx);
   if (rc)
      return rc;
   if (ctx->capacity > 0)
   {
      k_iota<<<blocks_for(ctx->capacity), kThreads, 0, ctx->stream>>>(ctx->idx_iota, ctx->capacity);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   return SPHB200_OK;
}

// ---- on-demand views of the voxel grid for the reference-facing getters -----
// SPH::mVoxelIds / mGrid (sph.h:142-146, 172) as of the LAST binning (the
// reference's members also hold the pre-integration binning after step()).
// voxel ids are written by every step; the per-voxel membership lists are only
// rebuilt here when a caller asks -- the GL view's drawVoxels needs just
// mGrid[c].count() (visualization.cpp:188-193).
namespace
{
__global__ void __launch_bounds__(kThreads) k_voxel_ids(DevParams P, const float4* __restrict__ pos4,
                                                         int* __restrict__ voxel_id)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= P.n)
      return;
   float4 p = pos4[i];
   voxel_id[i] = sph_voxel_id(sph_voxel_coord(p.x, P.h_times2_inv, P.gx), sph_voxel_coord(p.y, P.h_times2_inv, P.gy),
                              sph_voxel_coord(p.z, P.h_times2_inv, P.gz), P.gx, P.gy);
}

__global__ void __launch_bounds__(kThreads) k_histogram(int n, const int* __restrict__ keys,
                                                         uint32_t* __restrict__ count)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      atomicAdd(&count[keys[i]], 1u);
}
}  // namespace

int sph_refresh_voxel_ids(sphb200_ctx* ctx)
{
   if (ctx->voxel_ids_valid || ctx->n_local == 0)
      return SPHB200_OK;
   DevParams P = sph_dev_params(ctx);
   k_voxel_ids<<<blocks_for(ctx->n_local), kThreads, 0, ctx->stream>>>(P, ctx->pos4, ctx->voxel_id);
   ctx->launches++;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   ctx->voxel_ids_valid = true;
   return SPHB200_OK;
}

// mGrid[c].count() for every voxel, into a device array (enqueued on the context's stream)
int sph_grid_voxel_histogram(sphb200_ctx* ctx, uint32_t* d_counts)
{
   int rc = sph_refresh_voxel_ids(ctx);
   if (rc)
      return rc;
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * (size_t)ctx->cells_voxel, ctx->stream));
   if (ctx->n_local > 0)
   {
      k_histogram<<<blocks_for(ctx->n_local), kThreads, 0, ctx->stream>>>(ctx->n_local, ctx->voxel_id, d_counts);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   return SPHB200_OK;
}

int sph_download_grid(sphb200_ctx* ctx, int field, void* dst, size_t bytes)
{
   const int n = ctx->n_local;
   const int cells = ctx->cells_voxel;
   cudaStream_t st = ctx->stream;
   int rc = sph_refresh_voxel_ids(ctx);
   if (rc)
      return rc;
   if (!ctx->vg_count)
   {
      SPH_CUDA_CHECK(ctx, cudaMalloc(&ctx->vg_count, sizeof(uint32_t) * ((size_t)cells + 1)));
      SPH_CUDA_CHECK(ctx, cudaMalloc(&ctx->vg_start, sizeof(uint32_t) * ((size_t)cells + 1)));
      SPH_CUDA_CHECK(ctx, cudaMalloc(&ctx->vg_members, sizeof(uint32_t) * (size_t)(ctx->capacity + 1)));
      SPH_CUDA_CHECK(ctx, cudaMalloc(&ctx->vg_keys, sizeof(uint32_t) * (size_t)(ctx->capacity + 1)));
   }
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->vg_count, 0, sizeof(uint32_t) * ((size_t)cells + 1), st));
   if (n > 0)
   {
      k_histogram<<<blocks_for(n), kThreads, 0, st>>>(n, ctx->voxel_id, ctx->vg_count);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   size_t need = 0;
   const void* src = nullptr;
   if (field == SPHB200_F_CELL_COUNT)
   {
      need = sizeof(int) * (size_t)cells;
      src = ctx->vg_count;
   }
   else
   {
      size_t temp = ctx->cub_temp_bytes;
      SPH_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_temp, temp, ctx->vg_count, ctx->vg_start,
                                                        cells + 1, st));
      ctx->launches += 2;
      if (field == SPHB200_F_GRID_START)
      {
         need = sizeof(int) * ((size_t)cells + 1);
         src = ctx->vg_start;
      }
      else if (field == SPHB200_F_GRID_MEMBERS)
      {
         int bits = 1;
         while (bits < 32 && (1ll << bits) < (long long)cells)
            bits++;
         temp = ctx->cub_temp_bytes;
         if (n > 0)
            SPH_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_temp, temp, (const uint32_t*)ctx->voxel_id,
                                                                ctx->vg_keys, ctx->idx_iota, ctx->vg_members, n, 0,
                                                                bits, st));
         ctx->launches += 1 + 2 * ((bits + 7) / 8);
         need = sizeof(uint32_t) * (size_t)n;
         src = ctx->vg_members;
      }
      else
         return sph_fail(ctx, SPHB200_E_INVALID, "sph_download_grid: bad field");
   }
   if (bytes < need)
      return sph_fail(ctx, SPHB200_E_INVALID, "download: destination too small");
   if (need)
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, st));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
   return SPHB200_OK;
}
