// sph_sampled.cu -- REFERENCE_SAMPLED neighbour policy: the reference's own
// findNeighbors sub-sampler (sph.cpp:484-692) and the list-driven physics loops
// of SPH::step() (sph.cpp:242-289), one thread per particle.
//
// This mode is the bit-exact parity vehicle for the reference's default scene
// (ordered neighbour lists, voxel ids, membership); it is inherently serial per
// particle (LCG offset, early exit) so it is kept simple -- the throughput path
// is the FULL mode in sph_full.cu.  All FP is single-rounded in the reference's
// operation order, so density / acceleration / state agree with the reference's
// IEEE build to the last bit in practice.
#include "sph_math.cuh"

namespace
{

constexpr int kThreads = 128;

// findNeighbors (sph.cpp:484-692), quirks kept:
//  * the 8 octant slots with slot 3 overwritten (536-543): (0,0,sz) is never
//    visited and slot 4 is never assigned -> skipped;
//  * low-side bounds are strict: voxels on a 0-face are skipped (578-582);
//  * windows of 8 consecutive members from a wrapped-int LCG offset (590-604),
//    whole window dropped when any member index is out of range (609-620);
//  * only lanes 0..3 of a window are distance-tested (651-663);
//  * stop once more than E-8 neighbours are held (679-688).
// cell lists: members of voxel c are idx_sorted[cell_start[c] .. cell_start[c+1])
// in ascending particle index (stable sort == push_back order, 476-480).
__global__ void __launch_bounds__(kThreads)
   k_find_sampled(DevParams P, const float4* __restrict__ pos4, const int* __restrict__ voxel_id,
                  const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ idx_sorted,
                  uint32_t* __restrict__ nbr_idx, float* __restrict__ nbr_dist, int* __restrict__ nbr_count)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= P.n)
      return;
   float4 pi = pos4[i];
   int vid = voxel_id[i];
   int v[3];
   v[0] = vid % P.gx;
   v[1] = (vid / P.gx) % P.gy;
   v[2] = vid / (P.gx * P.gy);
   int s[3];
   s[0] = sph_upper_half(pi.x, v[0], P.h_times2, P.h) ? 1 : -1;
   s[1] = sph_upper_half(pi.y, v[1], P.h_times2, P.h) ? 1 : -1;
   s[2] = sph_upper_half(pi.z, v[2], P.h_times2, P.h) ? 1 : -1;
   const int E = P.examine;
   uint32_t* out_n = nbr_idx + (size_t)i * E;
   float* out_d = nbr_dist + (size_t)i * E;
   int found = 0;
   int visited = 0;
   bool done = false;
   // slot -> which axes move: bit0 x, bit1 y, bit2 z
   const int slot_axes[8] = {0, 1, 2, 3, -1, 5, 6, 7};
   for (int slot = 0; slot < 8 && !done; slot++)
   {
      int m = slot_axes[slot];
      if (m < 0)
         continue;
      int cx = v[0] + ((m & 1) ? s[0] : 0);
      int cy = v[1] + ((m & 2) ? s[1] : 0);
      int cz = v[2] + ((m & 4) ? s[2] : 0);
      if (!(cx > 0 && cx < P.gx && cy > 0 && cy < P.gy && cz > 0 && cz < P.gz))
         continue;
      int cell = sph_voxel_id(cx, cy, cz, P.gx, P.gy);
      int first_member = (int)cell_start[cell];
      int len = (int)cell_start[cell + 1] - first_member;
      if (len == 0)
         continue;
      int lcg = (int)(1664525u * (uint32_t)(i + visited) + 1013904223u);   // wraps like the int in sph.cpp:590
      int off = lcg % len;                                                 // C truncation keeps the sign
      visited++;
      int dir = (i & 1) ? -1 : 1;
      int base = 0;
      int windows = (len + 7) / 8;
      // The window positions do not depend on what the earlier windows found, only the
      // stopping rule does: fetch kAhead windows' members and positions at once (16
      // independent index loads, then 16 independent position loads) and apply the
      // reference's sequential rules to them.  One thread walks a whole voxel list
      // serially, so the dependent index -> position round trips are what this kernel
      // waits for.
      constexpr int kAhead = 4;
      for (int w0 = 0; w0 < windows && !done; w0 += kAhead)
      {
         uint32_t q[kAhead][4];
         float4 pq[kAhead][4];
         int valid = 0;                      // windows of this batch that pass the range test
#pragma unroll
         for (int a = 0; a < kAhead; a++)
         {
            int first = off + (base + 8 * a) * dir;
            bool ok = (w0 + a < windows) && valid == a && !(first < 0 || first + 7 >= len);
            if (ok)
            {
               valid = a + 1;
#pragma unroll
               for (int j = 0; j < 4; j++)
                  q[a][j] = idx_sorted[first_member + first + j];
            }
         }
#pragma unroll
         for (int a = 0; a < kAhead; a++)
            if (a < valid)
            {
#pragma unroll
               for (int j = 0; j < 4; j++)
                  pq[a][j] = pos4[q[a][j]];
            }
#pragma unroll
         for (int a = 0; a < kAhead; a++)
         {
            if (done || a >= valid)
               break;
            base += 8;
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
               if ((int)q[a][j] == i)
                  continue;
               float d2 = sph_dist2_exact(pi.x, pi.y, pi.z, pq[a][j].x, pq[a][j].y, pq[a][j].z);
               if (d2 < P.h2)
               {
                  out_n[found] = q[a][j];
                  out_d[found] = __fmul_rn(__fsqrt_rn(d2), P.scale);
                  found++;
               }
            }
            if (found > E - 8)
               done = true;
         }
         if (valid < kAhead)
            break;                           // a window fell out of range: this voxel is finished
      }
   }
   nbr_count[i] = found;
}

// computeDensity (sph.cpp:721-766) over the stored list
__global__ void __launch_bounds__(kThreads)
   k_density_list(DevParams P, const float4* __restrict__ pos4, const uint32_t* __restrict__ nbr_idx,
                  const float* __restrict__ nbr_dist, const int* __restrict__ nbr_count, float* __restrict__ rho)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= P.n)
      return;
   const int E = P.examine;
   int cnt = min(nbr_count[i], E);
   float sum = 0.0f;
   for (int k = 0; k < cnt; k++)
   {
      uint32_t q = nbr_idx[(size_t)i * E + k];
      if (q >= (uint32_t)P.n)
         break;
      if ((int)q == i)
         continue;
      float d = nbr_dist[(size_t)i * E + k];
      if (d > P.hs)
         continue;
      float t = __fsub_rn(P.hs2, __fmul_rn(d, d));
      t = __fmul_rn(__fmul_rn(t, t), t);
      float w = __fmul_rn(P.k1, t);
      sum = __fadd_rn(sum, __fmul_rn(pos4[q].w, w));
   }
   rho[i] = sum;
}

// computeAcceleration (sph.cpp:778-934) over the stored list, in list order
// (the in-loop `vt *= mu*rhoiInv`, 880-882, makes the order significant).
__global__ void __launch_bounds__(kThreads)
   k_accel_list(DevParams P, const float4* __restrict__ pos4, const float4* __restrict__ vel4,
                const float* __restrict__ rho, const uint32_t* __restrict__ nbr_idx,
                const float* __restrict__ nbr_dist, const int* __restrict__ nbr_count, float4* __restrict__ acc4)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= P.n)
      return;
   const int E = P.examine;
   float4 ri = pos4[i];
   float4 vi = vel4[i];
   float pi = __fmul_rn(__fsub_rn(rho[i], P.rho0), P.stiffness);
   float rhoi_inv = (pi > 0.0f) ? __fdiv_rn(1.0f, pi) : 1.0f;
   float pi_div = __fmul_rn(pi, __fmul_rn(rhoi_inv, rhoi_inv));
   float s = __fmul_rn(P.viscosity, rhoi_inv);
   Vec3 pg = {0.0f, 0.0f, 0.0f}, vt = {0.0f, 0.0f, 0.0f};
   int cnt = min(nbr_count[i], E);
   for (int k = 0; k < cnt; k++)
   {
      uint32_t q = nbr_idx[(size_t)i * E + k];
      float d = nbr_dist[(size_t)i * E + k];
      float rhoj = rho[q];
      float pj = __fmul_rn(__fsub_rn(rhoj, P.rho0), P.stiffness);
      float rhoj_inv = (rhoj > 0.0f) ? __fdiv_rn(1.0f, rhoj) : 1.0f;
      float rhoj_inv2 = __fmul_rn(rhoj_inv, rhoj_inv);
      float4 rj = pos4[q];
      float4 vj = vel4[q];
      float mj = rj.w;
      // grad W = f32( (K2 * rel) / (double)(d + 0.01) )   (854-856)
      double den = (double)d + 0.01;
      float gx = (float)((double)__fmul_rn(P.k2, __fmul_rn(__fsub_rn(ri.x, rj.x), P.scale)) / den);
      float gy = (float)((double)__fmul_rn(P.k2, __fmul_rn(__fsub_rn(ri.y, rj.y), P.scale)) / den);
      float gz = (float)((double)__fmul_rn(P.k2, __fmul_rn(__fsub_rn(ri.z, rj.z), P.scale)) / den);
      float c = __fsub_rn(P.hs, d);
      c = __fmul_rn(c, c);
      c = __fmul_rn(c, __fmul_rn(__fmul_rn(mj, pi_div), __fmul_rn(pj, rhoj_inv2)));
      pg.x = __fadd_rn(pg.x, __fmul_rn(gx, c));
      pg.y = __fadd_rn(pg.y, __fmul_rn(gy, c));
      pg.z = __fadd_rn(pg.z, __fmul_rn(gz, c));
      float cv = __fsub_rn(P.hs, d);
      cv = __fmul_rn(cv, __fmul_rn(__fmul_rn(rhoj_inv, mj), P.k3));
      vt.x = __fmul_rn(__fadd_rn(vt.x, __fmul_rn(__fsub_rn(vj.x, vi.x), cv)), s);
      vt.y = __fmul_rn(__fadd_rn(vt.y, __fmul_rn(__fsub_rn(vj.y, vi.y), cv)), s);
      vt.z = __fmul_rn(__fadd_rn(vt.z, __fmul_rn(__fsub_rn(vj.z, vi.z), cv)), s);
   }
   Vec3 a = sph_finish_acceleration(P, vt, pg, ri.x, ri.y, ri.z);
   acc4[i] = make_float4(a.x, a.y, a.z, 0.0f);
}

// integrate (sph.cpp:937-1022), in place; energies and neighbour statistics
// are block-reduced (the reference sums them serially in f32, 1004-1007, 226-230)
__global__ void __launch_bounds__(kThreads)
   k_integrate(DevParams P, float4* __restrict__ pos4, float4* __restrict__ vel4, const float4* __restrict__ acc4,
               const int* __restrict__ nbr_count, double* __restrict__ block_partials, StepScalars* scal)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   double ek = 0.0, ep = 0.0;
   unsigned long long cnt = 0;
   int cmax = -1, cmin = 0x7fffffff;
   if (i < P.n)
   {
      float4 p = pos4[i];
      float4 v = vel4[i];
      float4 a4 = acc4[i];
      float r[3] = {p.x, p.y, p.z};
      float vv[3] = {v.x, v.y, v.z};
      Vec3 a = {a4.x, a4.y, a4.z};
      float e_kin, e_pot;
      sph_integrate(P, r, vv, a, p.w, e_kin, e_pot);
      pos4[i] = make_float4(r[0], r[1], r[2], p.w);
      vel4[i] = make_float4(vv[0], vv[1], vv[2], 0.0f);
      ek = e_kin;
      ep = e_pot;
      int c = nbr_count[i];
      cnt = (unsigned long long)c;
      cmax = c;
      cmin = c;
   }
   sph_block_reduce_scalars(ek, ep, cnt, cmax, cmin, block_partials, scal);
}

int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

}  // namespace

int sph_step_sampled(sphb200_ctx* ctx)
{
   DevParams P = sph_dev_params(ctx);
   const int n = ctx->n_local;
   cudaStream_t st = ctx->stream;
   const bool timed = ctx->params.enable_timers != 0;
   if (timed) cudaEventRecord(ctx->ev[0], st);
   int rc = sph_bin_and_sort(ctx, false);
   if (rc)
      return rc;
   if (timed) cudaEventRecord(ctx->ev[1], st);
   rc = sph_reset_scalars(ctx);
   if (rc)
      return rc;
   int blocks = blocks_for(n);
   if (n > 0)
   {
      k_find_sampled<<<blocks, kThreads, 0, st>>>(P, ctx->pos4, ctx->voxel_id, ctx->cell_start, ctx->idx_sorted,
                                                  ctx->nbr_idx, ctx->nbr_dist, ctx->nbr_count);
      if (timed) cudaEventRecord(ctx->ev[2], st);
      k_density_list<<<blocks, kThreads, 0, st>>>(P, ctx->pos4, ctx->nbr_idx, ctx->nbr_dist, ctx->nbr_count,
                                                  ctx->rho);
      if (timed) cudaEventRecord(ctx->ev[3], st);
      if (timed) cudaEventRecord(ctx->ev[4], st);   // pressure loop is empty (sph.cpp:253-263)
      k_accel_list<<<blocks, kThreads, 0, st>>>(P, ctx->pos4, ctx->vel4, ctx->rho, ctx->nbr_idx, ctx->nbr_dist,
                                                ctx->nbr_count, ctx->acc4);
      if (timed) cudaEventRecord(ctx->ev[5], st);
      k_integrate<<<blocks, kThreads, 0, st>>>(P, ctx->pos4, ctx->vel4, ctx->acc4, ctx->nbr_count,
                                               ctx->d_block_partials, ctx->d_scalars);
      ctx->launches += 4;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
      rc = sph_finish_scalars(ctx, blocks);
      if (rc)
         return rc;
      if (timed) cudaEventRecord(ctx->ev[6], st);
   }
   ctx->lists_valid = true;
   ctx->snapshot_valid = false;
   return SPHB200_OK;
}
