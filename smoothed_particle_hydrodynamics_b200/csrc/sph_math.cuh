// sph_math.cuh -- per-particle device math shared by every kernel.
//
// Integer-producing arithmetic (binning, octant direction, the d2 < h2 test) uses
// explicitly rounded intrinsics (__fmul_rn / __fadd_rn / __fsub_rn): nvcc's
// default -fmad=true would contract `a - b*c` or `x*x + y*y` into FMAs and flip
// borderline cells / neighbours relative to the reference's IEEE build
// (SURVEY 7 "bit-exact integer outputs").  Each function cites the reference
// lines whose semantics it reproduces (/root/reference/src/sph.cpp).
#ifndef SPHB200_MATH_CUH
#define SPHB200_MATH_CUH

#include "sph_internal.h"

// (int)floor(x) as x86 evaluates it (cvttss2si): NaN and out-of-range give
// INT_MIN, which the clamp below turns into voxel 0 -- the GPU's saturating
// conversion would send +inf to the last voxel instead.
__device__ __forceinline__ int sph_floor_to_int_x86(float x)
{
   float f = floorf(x);
   if (!(f >= -2147483648.0f && f < 2147483648.0f))
      return (int)0x80000000;
   return (int)f;
}

// voxelizeParticles pass 1 (sph.cpp:452-463): one f32 multiply, floor, clamp.
__device__ __forceinline__ int sph_voxel_coord(float pos, float inv2h, int cells)
{
   int v = sph_floor_to_int_x86(__fmul_rn(pos, inv2h));
   if (v < 0) v = 0;
   if (v >= cells) v = cells - 1;
   return v;
}

// orientation inside the voxel and the octant direction (sph.cpp:504-515):
// returns 1 when (pos - v*2h) > h, else 0.
__device__ __forceinline__ int sph_upper_half(float pos, int v, float h_times2, float h)
{
   float o = __fsub_rn(pos, __fmul_rn((float)v, h_times2));
   return (o > h) ? 1 : 0;
}

// order-preserving map of a float to unsigned (-0 < +0): FULL mode orders the members of a fine cell by
// ascending (x, particle index) -- oracle/sph_oracle.c:x_order_key.  A NaN comes first: it is binned into
// cell 0 of its row, so it then sits in front of every run that contains it, where it only shifts the
// candidates behind it by one slot -- the packed sums (alternate candidates into two accumulators that
// are added at the end) do not change under such a shift, a NaN in the middle of a run would split it.
__device__ __forceinline__ uint32_t sph_x_order_key(float x)
{
   const uint32_t b = __float_as_uint(x);
   if (x != x)
      return 0u;
   return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// computeVoxelId (sph.cpp:1151-1154)
__device__ __forceinline__ int sph_voxel_id(int vx, int vy, int vz, int gx, int gy)
{
   return (vz * gy + vy) * gx + vx;
}

// squared distance exactly as sph.cpp:633-641 evaluates it in the IEEE build:
// ((dx*dx) + (dy*dy)) + (dz*dz), every operation rounded once.
__device__ __forceinline__ float sph_dist2_exact(float xi, float yi, float zi, float xj, float yj, float zj)
{
   float dx = __fsub_rn(xi, xj);
   float dy = __fsub_rn(yi, yj);
   float dz = __fsub_rn(zi, zj);
   return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// per-particle force coefficients derived from the density (sph.cpp:829-834,
// 860, 871), folded so the force pass reads two floats per neighbour:
//   fA = m_j * (p_j * rhojInv^2)      pressure factor
//   fB = rhojInv * m_j * K3           viscosity factor
__device__ __forceinline__ void sph_force_coeffs(const DevParams& P, float rho, float mass, float& fA, float& fB)
{
   float pj = (rho - P.rho0) * P.stiffness;
   float rinv = (rho > 0.0f) ? __fdiv_rn(1.0f, rho) : 1.0f;
   fA = mass * (pj * (rinv * rinv));
   fB = rinv * mass * P.k3;
}

struct Vec3
{
   float x, y, z;
};

// central point mass (sph.cpp:893-915 and 973-989): g = rel / (|rel| + eps)^3
__device__ __forceinline__ void sph_central_term(const DevParams& P, float rx, float ry, float rz, Vec3& g, float& d3)
{
   float ex = __fmul_rn(__fsub_rn(rx, P.cx), P.scale);
   float ey = __fmul_rn(__fsub_rn(ry, P.cy), P.scale);
   float ez = __fmul_rn(__fsub_rn(rz, P.cz), P.scale);
   float dot = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
   dot = __fsqrt_rn(dot);
   float sd = __fadd_rn(dot, P.softening);
   d3 = __fmul_rn(__fmul_rn(sd, sd), sd);
   g.x = __fdiv_rn(ex, d3);
   g.y = __fdiv_rn(ey, d3);
   g.z = __fdiv_rn(ez, d3);
}

// tail of computeAcceleration (sph.cpp:888-929) + the uniform-gravity switch:
// a = vt - pg + central gravity, CFL clamp, then + g.
__device__ __forceinline__ Vec3 sph_finish_acceleration(const DevParams& P, Vec3 vt, Vec3 pg, float rx, float ry,
                                                        float rz)
{
   Vec3 a;
   a.x = __fsub_rn(vt.x, pg.x);
   a.y = __fsub_rn(vt.y, pg.y);
   a.z = __fsub_rn(vt.z, pg.z);
   Vec3 g;
   float d3;
   sph_central_term(P, rx, ry, rz, g, d3);
   a.x = __fadd_rn(a.x, __fmul_rn(P.neg_gm, g.x));
   a.y = __fadd_rn(a.y, __fmul_rn(P.neg_gm, g.y));
   a.z = __fadd_rn(a.z, __fmul_rn(P.neg_gm, g.z));
   float dot = __fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z));
   if (dot > P.cfl2)
   {
      float len = __fsqrt_rn(dot);
      float sc = __fdiv_rn(P.cfl, len);
      a.x = __fmul_rn(a.x, sc);
      a.y = __fmul_rn(a.y, sc);
      a.z = __fmul_rn(a.z, sc);
   }
   if (P.use_gravity)
   {
      a.x = __fadd_rn(a.x, P.gvx);
      a.y = __fadd_rn(a.y, P.gvy);
      a.z = __fadd_rn(a.z, P.gvz);
   }
   return a;
}

// one axis of handleBoundaryConditions + applyBoundary (sph.cpp:1025-1148,
// dead code in the reference): reflect about the wall hit between the PRE-step
// position and the new one.
__device__ __forceinline__ void sph_wall_axis(const DevParams& P, const float old_pos[3], int axis, float wall_max,
                                              float np[3], float nv[3])
{
   float n[3] = {0.0f, 0.0f, 0.0f};
   float t;
   if (np[axis] < 0.0f)
   {
      n[axis] = 1.0f;
      t = __fdiv_rn(-old_pos[axis], nv[axis]);
   }
   else if (np[axis] > wall_max)
   {
      n[axis] = -1.0f;
      t = __fdiv_rn(__fsub_rn(wall_max, old_pos[axis]), nv[axis]);
   }
   else
      return;
   float hit[3], refl[3];
#pragma unroll
   for (int k = 0; k < 3; k++)
      hit[k] = __fadd_rn(old_pos[k], __fmul_rn(nv[k], t));
   float dot = __fadd_rn(__fadd_rn(__fmul_rn(nv[0], n[0]), __fmul_rn(nv[1], n[1])), __fmul_rn(nv[2], n[2]));
#pragma unroll
   for (int k = 0; k < 3; k++)
      refl[k] = __fsub_rn(nv[k], __fmul_rn(__fmul_rn(n[k], dot), 2.0f));
   float f = __fmul_rn(__fsub_rn(P.dt, t), P.damping);
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      nv[k] = refl[k];
      np[k] = __fadd_rn(hit[k], __fmul_rn(refl[k], f));
   }
}

// integrate (sph.cpp:937-1022) + the two switches (second half kick of the
// uniform field; wall reflection).  Returns the particle's energy terms.
__device__ __forceinline__ void sph_integrate(const DevParams& P, float r[3], float v[3], Vec3 a, float mass,
                                              float& e_kin, float& e_pot)
{
   float old_pos[3] = {r[0], r[1], r[2]};
   float acc[3] = {a.x, a.y, a.z};
   float vh[3], np[3], nv[3];
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      vh[k] = __fadd_rn(v[k], __fmul_rn(__fmul_rn(acc[k], P.dt), 0.5f));
      np[k] = __fadd_rn(r[k], __fmul_rn(vh[k], P.pos_dt));
   }
   Vec3 g;
   float d3;
   sph_central_term(P, np[0], np[1], np[2], g, d3);
   nv[0] = __fadd_rn(vh[0], __fmul_rn(__fmul_rn(P.neg_gm, g.x), P.dt));
   nv[1] = __fadd_rn(vh[1], __fmul_rn(__fmul_rn(P.neg_gm, g.y), P.dt));
   nv[2] = __fadd_rn(vh[2], __fmul_rn(__fmul_rn(P.neg_gm, g.z), P.dt));
   float dot = __fadd_rn(__fadd_rn(__fmul_rn(nv[0], nv[0]), __fmul_rn(nv[1], nv[1])), __fmul_rn(nv[2], nv[2]));
   e_kin = 0.0f;
   e_pot = 0.0f;
   if (dot > 0.0f)
   {
      e_kin = __fmul_rn(__fmul_rn(0.5f, mass), dot);
      e_pot = -__fdiv_rn(__fmul_rn(P.gm, mass), d3);
   }
   if (P.use_gravity)
   {
      nv[0] = __fadd_rn(nv[0], __fmul_rn(__fmul_rn(P.gvx, P.dt), 0.5f));
      nv[1] = __fadd_rn(nv[1], __fmul_rn(__fmul_rn(P.gvy, P.dt), 0.5f));
      nv[2] = __fadd_rn(nv[2], __fmul_rn(__fmul_rn(P.gvz, P.dt), 0.5f));
   }
   if (P.use_walls)
   {
      sph_wall_axis(P, old_pos, 0, P.max_x, np, nv);
      sph_wall_axis(P, old_pos, 1, P.max_y, np, nv);
      sph_wall_axis(P, old_pos, 2, P.max_z, np, nv);
   }
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      r[k] = np[k];
      v[k] = nv[k];
   }
}

// ---- slab halo messages (multi-GPU, sph_comm.cu) --------------------------------
__device__ __forceinline__ SlabEntry* sph_msg_entries(unsigned char* msg)
{
   return reinterpret_cast<SlabEntry*>(msg + sizeof(SlabMsgHeader));
}

// slot in a message section for every lane with `want`, one atomic per warp.  Must be
// called by all 32 lanes of a converged warp.
__device__ __forceinline__ unsigned sph_warp_append(unsigned* counter, bool want)
{
   const unsigned m = __ballot_sync(0xffffffffu, want);
   if (m == 0u)
      return 0u;
   const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
   unsigned base = 0;
   if (lane == leader)
      base = atomicAdd(counter, (unsigned)__popc(m));
   base = __shfl_sync(0xffffffffu, base, leader);
   return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}

// the same for a whole CTA: one global atomic per block (free slots come by the million
// and all append to one counter).  Must be called by every thread of the block.
__device__ __forceinline__ unsigned sph_block_append(unsigned* counter, bool want)
{
   __shared__ unsigned s_total, s_base;
   if (threadIdx.x == 0)
      s_total = 0;
   __syncthreads();
   const unsigned m = __ballot_sync(0xffffffffu, want);
   const int lane = threadIdx.x & 31;
   unsigned warp_off = 0;
   if (lane == 0 && m != 0u)
      warp_off = atomicAdd(&s_total, (unsigned)__popc(m));
   warp_off = __shfl_sync(0xffffffffu, warp_off, 0);
   __syncthreads();
   if (threadIdx.x == 0 && s_total != 0u)
      s_base = atomicAdd(counter, s_total);
   __syncthreads();
   return s_base + warp_off + (unsigned)__popc(m & ((1u << lane) - 1u));
}

// Exchange rules for an OWNED particle at position z (sph_comm.cu header): appends it to
// the outgoing messages as a migrant or as a boundary-layer ghost and returns the state
// its slot takes.  All 32 lanes of the warp call this; `owned` selects the real ones.  `slot` is
// the particle's slot: its global id (P.slot_gid, a scattered load) is fetched only for the few
// particles that actually go into a message.
__device__ __forceinline__ unsigned char sph_slab_emit(const DevParams& P, bool owned, float4 pos, float4 vel,
                                                       uint32_t slot)
{
   const int vz = sph_voxel_coord(pos.z, P.h_times2_inv, P.gz_global);
   // almost every warp is far from the slab faces: one vote instead of four appends
   if (!__any_sync(0xffffffffu, owned && ((P.has_up && vz >= P.own_z1 - 1) || (P.has_down && vz <= P.own_z0))))
      return SLOT_OWNED;
   const bool mig_up = owned && vz >= P.own_z1 && P.has_up;
   const bool mig_down = owned && !mig_up && vz < P.own_z0 && P.has_down;
   const bool migrant = mig_up || mig_down;
   const bool ghost_up = owned && !migrant && vz == P.own_z1 - 1 && P.has_up;
   const bool ghost_down = owned && !migrant && vz == P.own_z0 && P.has_down;   // 1-layer slabs ghost both ways
   // slots come from this rank's own counters (k_slab_finalize publishes them); the entries
   // go wherever the message lives -- in put mode that is the neighbour GPU's memory
   SlabEntry e;
   e.pos = pos;
   e.vel = vel;
   if (migrant || ghost_up || ghost_down)
      e.vel.w = __uint_as_float(P.slot_gid[slot]);
   bool overflow = false;
   unsigned s = sph_warp_append(&P.send_cnt[2], mig_up);
   if (mig_up)
   {
      if (s < (unsigned)P.mig_cap) sph_msg_entries(P.msg_up)[s] = e; else overflow = true;
   }
   s = sph_warp_append(&P.send_cnt[0], mig_down);
   if (mig_down)
   {
      if (s < (unsigned)P.mig_cap) sph_msg_entries(P.msg_down)[s] = e; else overflow = true;
   }
   s = sph_warp_append(&P.send_cnt[3], ghost_up);
   if (ghost_up)
   {
      if (s < (unsigned)P.ghost_cap) sph_msg_entries(P.msg_up)[P.mig_cap + s] = e; else overflow = true;
   }
   s = sph_warp_append(&P.send_cnt[1], ghost_down);
   if (ghost_down)
   {
      if (s < (unsigned)P.ghost_cap) sph_msg_entries(P.msg_down)[P.mig_cap + s] = e; else overflow = true;
   }
   if (overflow)
      atomicMax(&P.comm_counters[1], 1u);
   if (!migrant)
      return SLOT_OWNED;
   const bool keep_ghost = mig_up ? (vz == P.own_z1) : (vz == P.own_z0 - 1);
   return keep_ghost ? SLOT_LEAVING_GHOST : SLOT_LEAVING_FREE;
}

// block-wide sum of (e_kin, e_pot) in double + neighbour statistics, one
// partial per block (deterministic two-stage reduction; finished by
// sph_finish_scalars).  blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ void sph_block_reduce_scalars(double ek, double ep, unsigned long long cnt, int cmax,
                                                         int cmin, double* block_partials, StepScalars* scal,
                                                         int partial_index = -1)
{
   if (partial_index < 0)
      partial_index = (int)blockIdx.x;
   __shared__ double s_ek[32], s_ep[32];
   __shared__ unsigned long long s_cnt[32];
   __shared__ int s_max[32], s_min[32];
#pragma unroll
   for (int o = 16; o > 0; o >>= 1)
   {
      ek += __shfl_down_sync(0xffffffffu, ek, o);
      ep += __shfl_down_sync(0xffffffffu, ep, o);
      cnt += __shfl_down_sync(0xffffffffu, cnt, o);
      cmax = max(cmax, __shfl_down_sync(0xffffffffu, cmax, o));
      cmin = min(cmin, __shfl_down_sync(0xffffffffu, cmin, o));
   }
   int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   if (lane == 0)
   {
      s_ek[warp] = ek;
      s_ep[warp] = ep;
      s_cnt[warp] = cnt;
      s_max[warp] = cmax;
      s_min[warp] = cmin;
   }
   __syncthreads();
   if (warp == 0)
   {
      int nw = (blockDim.x + 31) >> 5;
      ek = lane < nw ? s_ek[lane] : 0.0;
      ep = lane < nw ? s_ep[lane] : 0.0;
      cnt = lane < nw ? s_cnt[lane] : 0ull;
      cmax = lane < nw ? s_max[lane] : -1;
      cmin = lane < nw ? s_min[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
         ek += __shfl_down_sync(0xffffffffu, ek, o);
         ep += __shfl_down_sync(0xffffffffu, ep, o);
         cnt += __shfl_down_sync(0xffffffffu, cnt, o);
         cmax = max(cmax, __shfl_down_sync(0xffffffffu, cmax, o));
         cmin = min(cmin, __shfl_down_sync(0xffffffffu, cmin, o));
      }
      if (lane == 0)
      {
         block_partials[2 * partial_index] = ek;
         block_partials[2 * partial_index + 1] = ep;
         atomicAdd(&scal->nbr_total, cnt);      // integer: order independent
         atomicMax(&scal->nbr_max, cmax);
         atomicMin(&scal->nbr_min, cmin);
      }
   }
}

#endif
