// sph_full.cu -- FULL neighbour policy: the throughput path.
//
// Two fused sweeps over particles in fine-cell order (cell edge h, 27-cell
// neighbourhood == the reference's 2x2x2 voxel octant, sph.cpp:504-556):
//   density sweep : poly6 density (computeDensity, sph.cpp:721-766) + the
//                   equation of state folded into two per-particle force
//                   coefficients (sph.cpp:785, 829-834) + velocity gather
//   force sweep   : spiky pressure + viscosity (computeAcceleration,
//                   sph.cpp:778-934) + integrate (937-1022) + wall collision
//                   (1025-1148) + energy / neighbour statistics
// Density sweep: a 512-thread CTA owns an 8x8x4 block of fine cells and stages the
// block plus its one-cell halo (60 x-contiguous row segments of the cell-sorted
// positions) in shared memory, in groups of four candidates with the coordinates
// split (SoA inside the group); each thread walks the 9 x-runs of its particle there,
// two candidates per instruction (packed FP32: fma.rn.f32x2 / add.rn.f32x2, sm_100a).
//
// The density sweep has to touch every candidate anyway, so it also records which
// candidates passed a (slightly enlarged) radius test as a HIT-MASK STREAM: one
// 8-byte record {32-bit mask, sorted index of the chunk's first candidate} per
// 32-candidate chunk that has a hit.  The force sweep is flat -- one thread per
// cell-sorted particle, nothing staged -- and simply walks the set bits of its
// particle's records: no second scan; it gathers those neighbours through L1 and
// applies the exact reference test (sph.cpp:633-653) and the pair body in ascending
// cell order (the order the in-loop viscosity scaling of sph.cpp:880-882 depends on).
// No per-particle neighbour list is stored.
// Blocks too dense for shared memory are split (8x8x2, 8x4x2, 4x4x2) and, past
// that, processed straight from global memory (scalar arithmetic, scan based).
// In slab mode (multi-GPU) the force sweep also appends its boundary-layer particles
// and migrants to the next halo messages (sph_slab_emit, sph_math.cuh).
#include <type_traits>

#include "sph_math.cuh"

namespace
{

// Tile shape, CTA size and staging capacity.  Overridable at compile time for A/B runs
// (tools/build_variant.py); the defaults are the measured best (profiles/r01_history.md).
#ifndef SPH_STREAM_HINTS
#define SPH_STREAM_HINTS 0       // 1: read-once / write-once data bypasses L1 allocation (ld.cs / st.cs) (A/B)
#endif
#if SPH_STREAM_HINTS
#define SPH_LD_ONCE(p) __ldcs(p)
#define SPH_ST_ONCE(p, v) __stcs(p, v)
#else
#define SPH_LD_ONCE(p) __ldg(p)
#define SPH_ST_ONCE(p, v) (*(p) = (v))
#endif
#ifndef SPH_DENS_SPLIT_ACC
#define SPH_DENS_SPLIT_ACC 0     // 1: separate poly6 accumulators for the two halves of a group (A/B)
#endif
#ifndef SPH_DENS_UNROLL
#define SPH_DENS_UNROLL 3        // unroll factor of the loop over the 9 runs: 3 = one z-plane per trip (1: 2.28 ms, 3: 2.24, 9: 2.37)
#endif
#ifndef SPH_TBY
#define SPH_TBY 8
#endif
#ifndef SPH_TBZ
#define SPH_TBZ 4
#endif
#ifndef SPH_TILE_THREADS
#define SPH_TILE_THREADS 512
#endif
#ifndef SPH_TILE_CTAS
#define SPH_TILE_CTAS 2
#endif
#ifndef SPH_CAP
#define SPH_CAP 6600
#endif
constexpr int kDensUnroll = SPH_DENS_UNROLL;
constexpr int TBX = 8, TBY = SPH_TBY, TBZ = SPH_TBZ;   // tile extent in fine cells (x rows are contiguous in memory)
constexpr int HROWS = (TBY + 2) * (TBZ + 2);   // halo rows (y,z) of a full tile
constexpr int CSW = TBX + 3;                // cell_start entries per halo row
constexpr int TROWS = TBY * TBZ;            // target rows of a full tile (<= 32: one warp scans them)
constexpr int kTileThreads = SPH_TILE_THREADS;
constexpr int kTileCtas = SPH_TILE_CTAS;    // resident CTAs per SM the kernel is built for
constexpr int kCap = SPH_CAP;               // staged particles per (sub-)tile, equal masses (12 B each)
constexpr int kCapMass = (12 * (kCap + 4)) / 16 - 4;   // ... with per-particle masses (16 B each): same bytes
#ifndef SPH_DENS_XTRIM
#define SPH_DENS_XTRIM 1         // 1: the density sweep cuts every x-run down to |dx| < h before testing candidates (the runs
                                 // are ascending in x since the in-cell order is by x): per halo row a table of the first
                                 // staged slot at or above each quarter-cell threshold, two lookups per run
#endif
#ifndef SPH_XQ
#define SPH_XQ 4
#endif
constexpr int XQ = SPH_XQ;                      // x thresholds per cell edge
constexpr int XT = (TBX + 2) * XQ + 2;          // table entries per halo row: thresholds 0 .. (TBX+2)*XQ, + the end
constexpr int WCAP = 32;                    // hit-mask records per particle (32 candidates each)
constexpr unsigned kNoStream = 0xffu;       // info.nw value: no stream, scan instead
constexpr unsigned kRunShift = 28;          // record base: sorted index | x-run number << 28
constexpr unsigned kBaseMask = (1u << kRunShift) - 1u;
#ifndef SPH_FORCE_RSM
#define SPH_FORCE_RSM 10   // 16: 2.49 ms, 12: 2.42, 10 and 8: 2.41, 6: 2.47, 4: 2.56 -- shared memory given up here is L1 for the gathers
#endif
constexpr int RSM = SPH_FORCE_RSM;                     // records per thread kept in shared memory by the force sweep
#ifndef SPH_FORCE_ILP
#define SPH_FORCE_ILP 2
#endif
#ifndef SPH_FORCE_CTAS
#define SPH_FORCE_CTAS 10  // resident CTAs per SM the force sweep is built for: 10 -> 48 registers (18 values spilled), 40 warps
                           // per SM.  With the velocity records on the texture path the sweep is latency bound (ncu r02:
                           // long-scoreboard stall 5 warps per issue, LSU 62 %, TEX 43 %, issue 68 %): 4 / 8 (64 regs):
                           // 2.31 ms, 9 / 10: 2.25 / 2.24, 12 (40 regs): 2.39; ILP 3 at 64 regs: 2.25, ILP 3 / 4 at 77 / 87: 2.75 / 3.14
#endif
constexpr int kForceIlp = SPH_FORCE_ILP;     // neighbours in flight per lane of the force sweep
#ifndef SPH_FORCE_THREADS
#define SPH_FORCE_THREADS 128
#endif
constexpr int kForceThreads = SPH_FORCE_THREADS;
#ifndef SPH_FORCE_TEX
#define SPH_FORCE_TEX 1    // neighbour gathers through the texture path: 1 = (v, fB) records, 2 = (x, fA), 3 = both
                           // (0: 2.41 ms, 1: 2.31, 2: 2.38, 3: 2.42 at 16.7M particles; the TEX path shares the L1 tag
                           // stage with LDG -- tools/ubench_gather.cu -- but copes better with scattered lanes)
#endif
constexpr int kFlatThreads = 128;
#ifndef SPH_FORCE_AOS
#define SPH_FORCE_AOS 0    // 1: the force-sweep records are 32-byte structures {x,y,z,fA, vx,vy,vz,fB} fetched with one
                           // 256-bit load per neighbour instead of two 128-bit loads from two arrays (A/B)
#endif

// force-sweep record access: two float4 arrays (A: x,y,z,fA  B: vx,vy,vz,fB), or -- SPH_FORCE_AOS -- one array
// of 32-byte records laid over the same allocation (B follows A, sph_capi.cu)
struct __align__(32) ForceRec
{
   float4 a, b;
};

__device__ __forceinline__ void rec_load(const float4* __restrict__ A, const float4* __restrict__ B, int j, float4& a,
                                         float4& b)
{
#if SPH_FORCE_AOS
   const ForceRec* r = reinterpret_cast<const ForceRec*>(A) + j;
   asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                : "l"(r));
#else
   a = __ldg(&A[j]);
   b = __ldg(&B[j]);
#endif
}

__device__ __forceinline__ void rec_store(float4* __restrict__ A, float4* __restrict__ B, int k, float4 a, float4 b)
{
#if SPH_FORCE_AOS
   ForceRec* r = reinterpret_cast<ForceRec*>(A) + k;
   r->a = a;
   r->b = b;
#else
   A[k] = a;
   B[k] = b;
#endif
}

struct TileLayout
{
   int row_g0[HROWS];        // first global (sorted) index of the halo row segment
   int row_delta[HROWS];     // smem index = global index + row_delta (0 when not staged)
   int row_len[HROWS];
   int cs[HROWS][CSW];       // cell_start of cells x0-1 .. x0+bx and the end, per halo row
   int tgt_off[TROWS + 1];   // prefix of target counts per target row
   int total;                // staged particles
   int ntargets;
   int rowk[TROWS];          // sorted index of target t of row r = rowk[r] + t
#if SPH_DENS_XTRIM
   ushort2 xtab[HROWS][XT];          // per halo row and x threshold t: .x = where a run may start when its lower bound
                                     // has index t, .y = where it must end when its upper bound has index t.  Both are the
                                     // first staged slot whose index is >= t -- unless the row is not ascending in x
                                     // (stage_rows_packed): then .x is the row's start and .y its end (no trimming)
#endif
   unsigned short tcell[kCap];   // per target: lx | r << 4 | hr0 << 9  (its cell; see locate_target)
};

// ---- pair arithmetic ---------------------------------------------------------

// poly6 term of one candidate, branch-free: t = max(hs2 - d2*scale^2, 0);
// sum += m_j * t^3  (K1 is applied once at the end).  Candidates outside h add 0.
__device__ __forceinline__ float density_term(float sum, float xi, float yi, float zi, float4 pj, float hs2,
                                              float scale2)
{
   float dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
   float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
   float t = fmaxf(fmaf(-d2, scale2, hs2), 0.0f);
   return fmaf(pj.w * t, t * t, sum);
}

// single-instruction MUFU approximations (relative error <= 2^-22, far inside the
// 1e-5 field tolerance; .ftz avoids the denormal pre/post scaling the default
// sqrtf / __fdividef expand to).  sqrt(0) = 0 for coincident particles.
__device__ __forceinline__ float sph_sqrt_approx(float x)
{
   float r;
   asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
   return r;
}

__device__ __forceinline__ float sph_rcp_approx(float x)
{
   float r;
   asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
   return r;
}

// ---- packed FP32 (sm_100a FFMA2 / FADD2): two candidates per instruction -------
// A 64-bit register pair holds the same coordinate of two consecutive candidates.
// tools/ubench_f32x2.cu: the packed density loop costs 8.9 cycles per candidate and
// scheduler against 14.0 for the scalar one (B200), same arithmetic.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c)
{
   f32x2 r;
   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
   return r;
}

__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b)
{
   f32x2 r;
   asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
   return r;
}

__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b)
{
   f32x2 r;
   asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
   return r;
}

__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
   f32x2 r;
   asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
   return r;
}

__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi)
{
   asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

struct ForceI     // per-target constants of computeAcceleration (sph.cpp:785-798)
{
   float x, y, z, vx, vy, vz;
   float pi_div;   // p_i * rhoiInv^2
   float s;        // mu * rhoiInv   (in-loop viscosity scale, sph.cpp:880-882)
};

// one neighbour of computeAcceleration's loop (sph.cpp:825-884).  (dx,dy,dz) =
// r_i - r_j and d2 are the exactly rounded values of the neighbour test; fA, fB are
// the neighbour's coefficients from sph_force_coeffs; (vx,vy,vz) its velocity.
template <bool UNIT_SCALE>
__device__ __forceinline__ void force_pair(const DevParams& P, const ForceI& I, float dx, float dy, float dz,
                                           float d2, float fA, float vx, float vy, float vz, float fB, Vec3& pg,
                                           Vec3& vt)
{
   float d = sph_sqrt_approx(d2);
   if (!UNIT_SCALE)
   {
      d *= P.scale;
      dx *= P.scale;
      dy *= P.scale;
      dz *= P.scale;
   }
   float inv = P.k2 * sph_rcp_approx(d + 0.01f);
   float hd = P.hs - d;
   float c = (hd * hd) * (I.pi_div * fA) * inv;
   pg.x = fmaf(dx, c, pg.x);
   pg.y = fmaf(dy, c, pg.y);
   pg.z = fmaf(dz, c, pg.z);
   float cv = hd * fB;
   vt.x = fmaf(vx - I.vx, cv, vt.x) * I.s;
   vt.y = fmaf(vy - I.vy, cv, vt.y) * I.s;
   vt.z = fmaf(vz - I.vz, cv, vt.z) * I.s;
}

// The same pair arithmetic split into an order-free part (force_term) and the
// ordered accumulation (force_accumulate) so that two candidates can be in flight.
struct PairTerm
{
   float px, py, pz;   // pressure-gradient contribution
   float wx, wy, wz;   // viscous contribution before the in-loop scaling
   int hit;
};

template <bool UNIT_SCALE>
__device__ __forceinline__ PairTerm force_term(const DevParams& P, const ForceI& I, float4 pj, float4 vj,
                                               bool not_self)
{
   float dx = __fsub_rn(I.x, pj.x), dy = __fsub_rn(I.y, pj.y), dz = __fsub_rn(I.z, pj.z);
   float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));   // sph.cpp:641
   PairTerm t;
   t.hit = (d2 < P.h2 && not_self) ? 1 : 0;
   float d = sph_sqrt_approx(d2);
   if (!UNIT_SCALE)
   {
      d *= P.scale;
      dx *= P.scale;
      dy *= P.scale;
      dz *= P.scale;
   }
   float inv = P.k2 * sph_rcp_approx(d + 0.01f);
   float hd = P.hs - d;
   float c = (hd * hd) * (I.pi_div * pj.w) * inv;
   t.px = dx * c;
   t.py = dy * c;
   t.pz = dz * c;
   float cv = hd * vj.w;
   t.wx = (vj.x - I.vx) * cv;
   t.wy = (vj.y - I.vy) * cv;
   t.wz = (vj.z - I.vz) * cv;
   return t;
}

__device__ __forceinline__ void force_accumulate(const ForceI& I, const PairTerm& t, Vec3& pg, Vec3& vt, int& count)
{
   if (t.hit)
   {
      pg.x += t.px;
      pg.y += t.py;
      pg.z += t.pz;
      vt.x = (vt.x + t.wx) * I.s;      // sph.cpp:875-882: scaled inside the loop
      vt.y = (vt.y + t.wy) * I.s;
      vt.z = (vt.z + t.wz) * I.s;
      count++;
   }
}

// exact test + pair body for candidate (pj, vj); returns 1 when it is a neighbour
template <bool UNIT_SCALE>
__device__ __forceinline__ int force_candidate(const DevParams& P, const ForceI& I, float4 pj, float4 vj,
                                               bool not_self, Vec3& pg, Vec3& vt)
{
   float dx = __fsub_rn(I.x, pj.x), dy = __fsub_rn(I.y, pj.y), dz = __fsub_rn(I.z, pj.z);
   float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));   // sph.cpp:641
   if (d2 < P.h2 && not_self)
   {
      force_pair<UNIT_SCALE>(P, I, dx, dy, dz, d2, pj.w, vj.x, vj.y, vj.z, vj.w, pg, vt);
      return 1;
   }
   return 0;
}

__device__ __forceinline__ ForceI make_force_i(const DevParams& P, float4 pi, float4 vi, float rho)
{
   ForceI I;
   I.x = pi.x; I.y = pi.y; I.z = pi.z;
   I.vx = vi.x; I.vy = vi.y; I.vz = vi.z;
   float p = (rho - P.rho0) * P.stiffness;
   float rinv = (p > 0.0f) ? __fdiv_rn(1.0f, p) : 1.0f;   // 1/p_i, not 1/rho_i (sph.cpp:786)
   I.pi_div = p * (rinv * rinv);
   I.s = P.viscosity * rinv;
   return I;
}

// density sweep epilogue for sorted particle k
__device__ __forceinline__ void density_store(const DevParams& P, int k, float4 pi, float rho,
                                              const uint32_t* __restrict__ idx_sorted,
                                              const float4* __restrict__ vel4, float4* __restrict__ s_posA4,
                                              float4* __restrict__ s_velB4, float* __restrict__ s_rho)
{
   float fA, fB;
   sph_force_coeffs(P, rho, pi.w, fA, fB);
   // step_host uploads the velocities while this sweep runs: then k_gather_vel fills them in
   float4 v = P.defer_velocity ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : __ldg(&vel4[idx_sorted[k]]);
   s_rho[k] = rho;
   rec_store(s_posA4, s_velB4, k, make_float4(pi.x, pi.y, pi.z, fA), make_float4(v.x, v.y, v.z, fB));
}

// force sweep epilogue: tail of computeAcceleration, integrate, walls, write-back
__device__ __forceinline__ void force_store(const DevParams& P, int k, const ForceI& I, Vec3 vt, Vec3 pg, int count,
                                            const float4* __restrict__ s_pos4,
                                            const uint32_t* __restrict__ idx_sorted, float4* __restrict__ pos4,
                                            float4* __restrict__ vel4, float4* __restrict__ s_acc4,
                                            int* __restrict__ s_count, double& ek, double& ep, float4& new_pos,
                                            float4& new_vel)
{
   Vec3 a = sph_finish_acceleration(P, vt, pg, I.x, I.y, I.z);
   float mass = s_pos4[k].w;
   float r[3] = {I.x, I.y, I.z};
   float v[3] = {I.vx, I.vy, I.vz};
   float e_kin, e_pot;
   sph_integrate(P, r, v, a, mass, e_kin, e_pot);
   uint32_t o = idx_sorted[k];
   new_pos = make_float4(r[0], r[1], r[2], mass);
   new_vel = make_float4(v[0], v[1], v[2], 0.0f);
   SPH_ST_ONCE(&pos4[o], new_pos);
   SPH_ST_ONCE(&vel4[o], new_vel);
   SPH_ST_ONCE(&s_acc4[k], make_float4(a.x, a.y, a.z, 0.0f));
   SPH_ST_ONCE(&s_count[k], count);
   ek = e_kin;
   ep = e_pot;
}

// the 9 x-runs of sorted particle k straight from the global cell table
__device__ __forceinline__ void global_runs(const DevParams& P, uint32_t key, const uint32_t* __restrict__ cell_start,
                                            int b[9], int e[9])
{
   int cx = (int)(key % (uint32_t)P.fx);
   int t = (int)(key / (uint32_t)P.fx);
   int cy = t % P.fy, cz = t / P.fy;
   int x0 = max(cx - 1, 0), x1 = min(cx + 1, P.fx - 1);
#pragma unroll
   for (int dz = 0; dz < 3; dz++)
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
      {
         int z = cz + dz - 1, y = cy + dy - 1;
         int r = dz * 3 + dy;
         if (z < 0 || z >= P.fz || y < 0 || y >= P.fy)
         {
            b[r] = 0;
            e[r] = 0;
         }
         else
         {
            int row = (z * P.fy + y) * P.fx;
            b[r] = (int)cell_start[row + x0];
            e[r] = (int)cell_start[row + x1 + 1];
         }
      }
}

// ---- untiled kernels (kernel_variant = 1; also the arithmetic of the dense
//      fallback): one thread per sorted particle, candidates through L1/L2 -----

__global__ void __launch_bounds__(kFlatThreads)
   k_density_flat(DevParams P, const float4* __restrict__ s_pos4, const uint32_t* __restrict__ keys_sorted,
                  const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ idx_sorted,
                  const float4* __restrict__ vel4, float4* __restrict__ s_posA4, float4* __restrict__ s_velB4,
                  float* __restrict__ s_rho, unsigned* __restrict__ hit_info)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= sph_live_count(P))
      return;
   hit_info[k] = kNoStream;   // the force sweep scans for these particles
   float4 pi = s_pos4[k];
   int b[9], e[9];
   global_runs(P, keys_sorted[k], cell_start, b, e);
   float scale2 = P.scale * P.scale;
   float sum = 0.0f;
#pragma unroll
   for (int r = 0; r < 9; r++)
      for (int j = b[r]; j < e[r]; j++)
         sum = density_term(sum, pi.x, pi.y, pi.z, __ldg(&s_pos4[j]), P.hs2, scale2);
   float self = density_term(0.0f, pi.x, pi.y, pi.z, pi, P.hs2, scale2);   // own term (0 for a NaN position)
   density_store(P, k, pi, P.k1 * (sum - self), idx_sorted, vel4, s_posA4, s_velB4, s_rho);
}

// ---- tiled kernels -----------------------------------------------------------

struct SubTile
{
   int x0, y0, z0;      // origin in fine cells
   int bx, by, bz;      // extent in fine cells
};

__device__ __forceinline__ int warp_sum(int v)
{
   return __reduce_add_sync(0xffffffffu, v);
}

// number of particles in the sub-tile plus its one-cell halo (warp-cooperative)
__device__ int halo_population(const DevParams& P, const SubTile& t, const uint32_t* __restrict__ cell_start)
{
   int lane = threadIdx.x & 31;
   int rows = (t.by + 2) * (t.bz + 2);
   int sum = 0;
   for (int hr = lane; hr < rows; hr += 32)
   {
      int y = t.y0 - 1 + hr % (t.by + 2);
      int z = t.z0 - 1 + hr / (t.by + 2);
      if (y >= 0 && y < P.fy && z >= 0 && z < P.fz)
      {
         int row = (z * P.fy + y) * P.fx;
         int xa = max(t.x0 - 1, 0), xb = min(t.x0 + t.bx + 1, P.fx);
         sum += (int)cell_start[row + xb] - (int)cell_start[row + xa];
      }
   }
   return warp_sum(sum);
}

// fills the layout for one sub-tile.  Called by all threads; ends with a barrier.
// staged = true lays the halo rows out back to back in shared memory.
__device__ void setup_layout(const DevParams& P, const SubTile& t, const uint32_t* __restrict__ cell_start,
                             bool staged, TileLayout& L)
{
   const int rows = (t.by + 2) * (t.bz + 2);
   const int csw = t.bx + 3;
   for (int i = threadIdx.x; i < rows * csw; i += blockDim.x)
   {
      int hr = i / csw, c = i % csw;
      int y = t.y0 - 1 + hr % (t.by + 2);
      int z = t.z0 - 1 + hr / (t.by + 2);
      int v = 0;
      if (y >= 0 && y < P.fy && z >= 0 && z < P.fz)
      {
         int x = min(max(t.x0 - 1 + c, 0), P.fx);   // x == fx addresses the end of the row
         v = (int)cell_start[(z * P.fy + y) * P.fx + x];
      }
      L.cs[hr][c] = v;
   }
   __syncthreads();
   if (threadIdx.x < 32)
   {
      // row lengths and their exclusive prefix (rows <= 36: two lanes-wide passes)
      int lane = threadIdx.x;
      int carry = 0;
      for (int base = 0; base < rows; base += 32)
      {
         int hr = base + lane;
         int len = 0, g0 = 0;
         if (hr < rows)
         {
            g0 = L.cs[hr][0];
            len = L.cs[hr][csw - 1] - g0;
         }
         // staged rows start on a group boundary (4 candidates): the packed sweep widens
         // its runs to whole groups and must stay inside the row (stage_rows_packed
         // fills the slack with far-away sentinels)
         const int plen = (len + 3) & ~3;
         int incl = plen;
#pragma unroll
         for (int o = 1; o < 32; o <<= 1)
         {
            int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
               incl += up;
         }
         if (hr < rows)
         {
            L.row_g0[hr] = g0;
            L.row_len[hr] = len;
            L.row_delta[hr] = staged ? (carry + incl - plen) - g0 : 0;
         }
         carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0)
         L.total = carry;
      // targets per target row (by*bz <= 16 rows)
      int trows = t.by * t.bz;
      int tc = 0;
      if (lane < trows)
      {
         int hr = (lane / t.by + 1) * (t.by + 2) + (lane % t.by + 1);
         tc = L.cs[hr][t.bx + 1] - L.cs[hr][1];
      }
      int incl = tc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
         int up = __shfl_up_sync(0xffffffffu, incl, o);
         if (lane >= o)
            incl += up;
      }
      if (lane < trows)
         L.tgt_off[lane] = incl - tc;
      if (lane == 31)
      {
         L.tgt_off[trows] = incl;
         L.ntargets = incl;
      }
   }
   __syncthreads();
   // target -> cell table (staged tiles hold at most kCap particles): one thread per
   // target cell writes the code of its particles
   const int trows = t.by * t.bz;
   for (int c = threadIdx.x; staged && c < trows * t.bx; c += blockDim.x)
   {
      int r = c / t.bx, x = c - r * t.bx + 1;
      int hr0 = (r / t.by + 1) * (t.by + 2) + (r % t.by + 1);
      int first = L.cs[hr0][1];
      int t0 = L.tgt_off[r] + (L.cs[hr0][x] - first);
      int cnt = L.cs[hr0][x + 1] - L.cs[hr0][x];
      unsigned short code = (unsigned short)(x | (r << 4) | (hr0 << 9));
      for (int i = 0; i < cnt; i++)
         L.tcell[t0 + i] = code;
      if (x == 1)
         L.rowk[r] = first - L.tgt_off[r];
   }
   __syncthreads();
}

struct Target
{
   int k;        // global sorted index
   int hr0;      // halo row of the target's own cell row
   int lx;       // cell column inside the halo row table (1..bx)
};

// maps flat target number -> particle and its cell by searching the prefix tables
// (unstaged fallback: any number of targets)
__device__ __forceinline__ Target locate_target_search(const SubTile& t, const TileLayout& L, int tnum)
{
   int trows = t.by * t.bz;
   int r = 0;
#pragma unroll 4
   for (int i = 1; i < trows; i++)
      r += (tnum >= L.tgt_off[i]) ? 1 : 0;
   Target T;
   T.hr0 = (r / t.by + 1) * (t.by + 2) + (r % t.by + 1);
   T.k = L.cs[T.hr0][1] + (tnum - L.tgt_off[r]);
   int lx = 1;
   for (int i = 2; i <= t.bx; i++)
      lx += (T.k >= L.cs[T.hr0][i]) ? 1 : 0;
   T.lx = lx;
   return T;
}

// maps flat target number -> particle and its cell (table built by setup_layout)
__device__ __forceinline__ Target locate_target(const TileLayout& L, int tnum)
{
   unsigned code = L.tcell[tnum];
   Target T;
   T.lx = (int)(code & 15u);
   T.hr0 = (int)(code >> 9);
   T.k = L.rowk[(code >> 4) & 31u] + tnum;
   return T;
}

// picks the sub-division level of this CTA's tile: 0 = whole tile ... 3 = every axis halved,
// 4 = nothing fits (process from global memory).  Evaluated by warp 0.
__device__ int choose_level(const DevParams& P, int X0, int Y0, int Z0, const uint32_t* __restrict__ cell_start,
                            int cap)
{
   for (int level = 0; level < 4; level++)
   {
      int bz = level >= 1 ? TBZ / 2 : TBZ;
      int by = level >= 2 ? TBY / 2 : TBY;
      int bx = level >= 3 ? TBX / 2 : TBX;
      bool fits = true;
      for (int z = 0; z < TBZ && fits; z += bz)
         for (int y = 0; y < TBY && fits; y += by)
            for (int x = 0; x < TBX && fits; x += bx)
            {
               SubTile t = {X0 + x, Y0 + y, Z0 + z, bx, by, bz};
               fits = halo_population(P, t, cell_start) + 3 * (by + 2) * (bz + 2) <= cap;   // + row padding
            }
      if (fits)
         return level;
   }
   return 4;
}

__device__ __forceinline__ void tile_origin(int& X0, int& Y0, int& Z0)
{
   X0 = blockIdx.x * TBX;     // 3D launch grid: no per-thread divisions
   Y0 = blockIdx.y * TBY;
   Z0 = blockIdx.z * TBZ;
}

// particles of the un-haloed tile (quick exit for empty space); warp 0
__device__ int tile_population(const DevParams& P, int X0, int Y0, int Z0, const uint32_t* __restrict__ cell_start)
{
   int lane = threadIdx.x & 31;
   int sum = 0;
   if (lane < TROWS)
   {
      int y = Y0 + lane % TBY, z = Z0 + lane / TBY;
      if (y < P.fy && z < P.fz)
      {
         int row = (z * P.fy + y) * P.fx;
         sum = (int)cell_start[row + min(X0 + TBX, P.fx)] - (int)cell_start[row + X0];
      }
   }
   return warp_sum(sum);
}

// hit-mask stream addressing: the records of 32 consecutive sorted particles are
// interleaved so that a warp's lanes write / read 256 contiguous bytes per record
__device__ __forceinline__ size_t stream_base(int k)
{
   return ((size_t)(k >> 5) * WCAP) * 32 + (size_t)(k & 31);
}

// Scalar density sweep over one (sub-)tile: the fallback for blocks too dense to stage
// (STAGED = false: candidates from global memory, particles flagged "scan me" for the
// force sweep).  The staged path is density_targets_packed below.  STAGED: candidates come from shared memory
// and the hit-mask stream is written; otherwise candidates come from global
// memory and the particle is flagged "scan me" for the force sweep.
//
// The inner loop is bound by the FMA pipe (ncu r1-final: 24.7 cycles per candidate
// and scheduler for 12 FMA-class instructions -- 3-operand FP32 ops issue every other
// cycle), so the fast path is written for the fewest FMA-class operations:
//   UNIT (simulation scale == 1): t = hs2 - dx^2 - dy^2 - dz^2 as three chained FMAs
//   instead of mul + 2 fma (d2) + fma (t) + fma (hit test); the hit bit comes from a
//   compare (ALU pipe) against -1e-5 hs2;
//   UMASS (all masses equal, as in the reference, sph.cpp:88,105-108): the mass is
//   factored out of the sum (one multiply less per candidate).
// 8 FMA-class instructions per candidate instead of 12.
template <bool STAGED, bool UNIT, bool UMASS>
__device__ __forceinline__ void density_targets(const DevParams& P, const SubTile& t, const TileLayout& L,
                                                const float4* __restrict__ src,
                                                const uint32_t* __restrict__ idx_sorted,
                                                const float4* __restrict__ vel4, float4* __restrict__ s_posA4,
                                                float4* __restrict__ s_velB4, float* __restrict__ s_rho,
                                                uint2* __restrict__ hit_rec, unsigned* __restrict__ hit_info)
{
   const float scale2 = P.scale * P.scale;
   const float tmin = -1e-5f * P.hs2;         // enlarged radius: a superset of the exact d2 < h2 test
   for (int tnum = threadIdx.x; tnum < L.ntargets; tnum += blockDim.x)
   {
      Target T = locate_target_search(t, L, tnum);
      const float4 pi = src[T.k + L.row_delta[T.hr0]];
      uint2* rec = hit_rec + stream_base(T.k);
      float sum = 0.0f;
      int nw = 0, nhits = 0;
#pragma unroll 1
      for (int r = 0; r < 9; r++)
      {
         int hr = T.hr0 + (r / 3 - 1) * (t.by + 2) + (r - (r / 3) * 3 - 1);
         int delta = L.row_delta[hr];
         int b = L.cs[hr][T.lx - 1] + delta;
         int e = L.cs[hr][T.lx + 2] + delta;
#pragma unroll 1
         for (int c0 = b; c0 < e; c0 += 32)
         {
            const float4* p = src + c0;
            const int n = min(32, e - c0);
            const float4* pe = p + n;
            // one hit bit per candidate (the sign of tmin - t), shifted in at the LSB: after
            // the chunk candidate i sits at bit 31 - i
            unsigned mask = 0;
#pragma unroll 4
            for (; p < pe; p++)
            {
               float4 pj = STAGED ? *p : __ldg(p);
               float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
               float tt;
               if (UNIT)
                  tt = fmaf(-dz, dz, fmaf(-dy, dy, fmaf(-dx, dx, P.hs2)));
               else
                  tt = fmaf(-fmaf(dz, dz, fmaf(dy, dy, dx * dx)), scale2, P.hs2);
               mask = __funnelshift_l(__float_as_uint(tmin - tt), mask, 1);   // sign bit: tt > tmin
               float tc = fmaxf(tt, 0.0f);
               if (UMASS)
                  sum = fmaf(tc * tc, tc, sum);
               else
                  sum = fmaf(pj.w * tc, tc * tc, sum);
            }
            mask <<= (32 - n);
            if (STAGED && mask != 0u)
            {
               if (nw < WCAP)
                  rec[(size_t)nw * 32] = make_uint2(mask, (unsigned)(c0 - delta) | ((unsigned)r << 28));
               nw++;
               nhits += __popc(mask);
            }
         }
      }
      if (STAGED)
         hit_info[T.k] = nw <= WCAP ? ((unsigned)nw | ((unsigned)nhits << 8)) : kNoStream;
      else
         hit_info[T.k] = kNoStream;
      // the particle itself sat in the centre run with t = hs2: remove its own term (the
      // reference skips realIndex == particleIndex, sph.cpp:737); a NaN or infinite position has term 0
      float t_self = (pi.x - pi.x == 0.0f && pi.y - pi.y == 0.0f && pi.z - pi.z == 0.0f) ? P.hs2 : 0.0f;   // (inf - inf is NaN too)
      float self = UMASS ? (t_self * t_self) * t_self : __fmul_rn(pi.w * t_self, t_self * t_self);
      float rho = UMASS ? (P.k1 * pi.w) * (sum - self) : P.k1 * (sum - self);
      density_store(P, T.k, pi, rho, idx_sorted, vel4, s_posA4, s_velB4, s_rho);
   }
}

// ---- packed density sweep (the staged path) -----------------------------------
// Shared-memory layout: candidates in groups of four, structure of arrays inside the
// group -- {x0 x1 x2 x3}{y0..y3}{z0..z3}[{m0..m3}] -- so that one LDS.128 per
// coordinate feeds two packed instructions (candidates 0,1 and 2,3).  Staged index s
// lives in group s >> 2, lane s & 3.  Equal masses (UMASS) drop the fourth row.
// x threshold index of a coordinate: floor(x * XQ / h) relative to the tile's first halo cell, clamped to the
// table.  ONE function for candidates (table build) and targets (lookups): it is monotone in x, so "every
// candidate with x >= lo has an index >= index(lo)" holds whatever the rounding.  NaN -> 0 (they come first).
__device__ __forceinline__ int x_threshold(float x, float inv_q, int m0, int hi)
{
   // the offset is subtracted in float (exact for every value near the table: both operands are below 2^24
   // quarter-cells and the difference is smaller than either) so that the conversion saturates for the
   // far-out-of-box values instead of an integer subtraction wrapping around; NaN converts to 0
   return min(max(__float2int_rd(x * inv_q - (float)m0), 0), hi);
}

template <bool UMASS>
__device__ __forceinline__ void stage_rows_packed(const DevParams& P, const SubTile& t, TileLayout& L,
                                                  const float4* __restrict__ src, float* __restrict__ sg)
{
   constexpr int GF = UMASS ? 12 : 16;   // floats per group
   const int rows = (t.by + 2) * (t.bz + 2);
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#if SPH_DENS_XTRIM
   const float inv_q = P.h_times2_inv * (2.0f * XQ);     // XQ / h
   const int m0 = (t.x0 - 1) * XQ, mt = (t.bx + 2) * XQ; // thresholds 0 .. mt cover the halo row's cells
#endif
   for (int hr = warp; hr < rows; hr += nwarps)
   {
      int g0 = L.row_g0[hr], len = L.row_len[hr], off = g0 + L.row_delta[hr];
#if SPH_DENS_XTRIM
      int m_carry = -1;      // threshold index of the previous particle of the row
      bool descends = false; // a row is ascending in x unless a position overflowed the reference's (int)floor: such a
                             // particle (x * 1/(2h) >= 2^31) is binned into cell 0 like a NaN but sorts last there
#endif
      for (int j0 = 0; j0 < len; j0 += 32)
      {
         const int j = j0 + lane;
         float4 p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
         if (j < len)
         {
            p = __ldg(&src[g0 + j]);
            int s = off + j;
            float* d = sg + (s >> 2) * GF + (s & 3);
            d[0] = p.x;
            d[4] = p.y;
            d[8] = p.z;
            if (!UMASS)
               d[12] = p.w;
         }
#if SPH_DENS_XTRIM
         // xtab[t] = first staged slot of the row whose index is >= t: a particle fills the thresholds
         // between its left neighbour's index (exclusive) and its own (inclusive); the row is ascending in x
         int m = j < len ? x_threshold(p.x, inv_q, m0, mt) : mt;
         int left = __shfl_up_sync(0xffffffffu, m, 1);
         if (lane == 0)
            left = m_carry;
         if (j < len)
         {
            descends |= m < left;
            for (int q = left + 1; q <= m; q++)
               L.xtab[hr][q] = make_ushort2((unsigned short)(off + j), (unsigned short)(off + j));
         }
         m_carry = __shfl_sync(0xffffffffu, m, min(31, len - 1 - j0));   // the last particle of this block of 32
#endif
      }
#if SPH_DENS_XTRIM
      // thresholds above the last particle (all of them in an empty row) point at the row's end
      for (int q = m_carry + 1 + lane; q <= mt + 1; q += 32)
         L.xtab[hr][q] = make_ushort2((unsigned short)(off + len), (unsigned short)(off + len));
      if (__any_sync(0xffffffffu, descends))
      {
         __syncwarp();
         for (int q = lane; q <= mt + 1; q += 32)
            L.xtab[hr][q] = make_ushort2((unsigned short)off, (unsigned short)(off + len));
      }
#endif
      // runs are widened to whole groups: the unused lanes of the row's last group hold
      // a position that is far from everything (and no mass)
      int s = off + len + lane;
      if (s < off + ((len + 3) & ~3))
      {
         float* d = sg + (s >> 2) * GF + (s & 3);
         d[0] = 1e30f;
         d[4] = 0.0f;
         d[8] = 0.0f;
         if (!UMASS)
            d[12] = 0.0f;
      }
   }
}

// Per target: the 9 x-runs, widened to group boundaries, in chunks of 32 candidates
// (8 groups).  With e = d2 * scale^2 - hs2 per candidate (two per instruction):
//   hit bit = sign of (e - 1e-5 hs2): a superset of the exact d2 < h2 test, the force
//             sweep applies the exact one to the survivors;
//   poly6   : sum += min(e, 0)^3 = -(hs2 - d2)^3 inside the radius, 0 outside.
// Rows start on group boundaries, so the candidates a widened run picks up outside its
// three cells are particles of the SAME row two or more columns away (or sentinels):
// they are masked out of the hit bits, and are at least h away, so their poly6 term is
// 0 (or one ulp of nothing at exactly h).
// Per pair of candidates: 1.5 LDS.128 + 9 packed FMA-pipe instructions + 2 FMNMX +
// 2 SHF, against 2 x (1 + 9 + 1 + 1) for the scalar loop.
// 128-bit shared-memory load at a compile-time byte offset from a shared-space address
// (keeps the address arithmetic out of the unrolled groups: one register + immediate)
template <int OFF>
__device__ __forceinline__ ulonglong2 lds128(unsigned addr)
{
   ulonglong2 v;
   asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+%3];" : "=l"(v.x), "=l"(v.y) : "r"(addr), "n"(OFF));
   return v;
}

// one group of four candidates (two packed pairs) of the density sweep; the group
// starts OFF bytes from shared address `a`
template <bool UNIT, bool UMASS, int OFF>
__device__ __forceinline__ void density_group(unsigned a, f32x2 NX, f32x2 NY, f32x2 NZ, f32x2 NH, f32x2 NTHR,
                                              f32x2 S2, unsigned& mask, f32x2& sum2, f32x2& sum2b)
{
   const ulonglong2 X = lds128<OFF>(a), Y = lds128<OFF + 16>(a), Z = lds128<OFF + 32>(a);
   ulonglong2 M;
   if (!UMASS)
      M = lds128<OFF + 48>(a);
#pragma unroll
   for (int half = 0; half < 2; half++)
   {
      f32x2 dx = fadd2(half ? X.y : X.x, NX);
      f32x2 dy = fadd2(half ? Y.y : Y.x, NY);
      f32x2 dz = fadd2(half ? Z.y : Z.x, NZ);
      f32x2 ee;
      if (UNIT)
         ee = ffma2(dz, dz, ffma2(dy, dy, ffma2(dx, dx, NH)));
      else
         ee = ffma2(ffma2(dz, dz, ffma2(dy, dy, fmul2(dx, dx))), S2, NH);
      float s0, s1, e0, e1;
      unpack2(fadd2(ee, NTHR), s0, s1);
      mask = __funnelshift_l(__float_as_uint(s0), mask, 1);
      mask = __funnelshift_l(__float_as_uint(s1), mask, 1);
      unpack2(ee, e0, e1);
      f32x2 u = pack2(fminf(e0, 0.0f), fminf(e1, 0.0f));
      f32x2& acc = (SPH_DENS_SPLIT_ACC && half) ? sum2b : sum2;
      if (UMASS)
         acc = ffma2(fmul2(u, u), u, acc);
      else
         acc = ffma2(fmul2(half ? M.y : M.x, u), fmul2(u, u), acc);
   }
}

template <bool UNIT, bool UMASS>
__device__ __forceinline__ void density_targets_packed(const DevParams& P, const SubTile& t, TileLayout& L,
                                                       const float* __restrict__ sg,
                                                       const float4* __restrict__ s_pos4,
                                                       const uint32_t* __restrict__ idx_sorted,
                                                       const float4* __restrict__ vel4, float4* __restrict__ s_posA4,
                                                       float4* __restrict__ s_velB4, float* __restrict__ s_rho,
                                                       uint2* __restrict__ hit_rec, unsigned* __restrict__ hit_info)
{
   constexpr int GF = UMASS ? 12 : 16;   // floats per group
   const float scale2 = P.scale * P.scale;
   const float thr = 1e-5f * P.hs2;
   const f32x2 NH = pack2(-P.hs2, -P.hs2), NTHR = pack2(-thr, -thr), S2 = pack2(scale2, scale2);
   const unsigned sbase = (unsigned)__cvta_generic_to_shared(sg);
   const int rowstep = t.by + 2;
#if SPH_DENS_XTRIM
   const float inv_q = P.h_times2_inv * (2.0f * XQ);     // XQ / h
   const int m0 = (t.x0 - 1) * XQ, mt = (t.bx + 2) * XQ;
#endif
   for (int tnum = threadIdx.x; tnum < L.ntargets; tnum += blockDim.x)
   {
      Target T = locate_target(L, tnum);
      const float4 pi = __ldg(&s_pos4[T.k]);
      const f32x2 NX = pack2(-pi.x, -pi.x), NY = pack2(-pi.y, -pi.y), NZ = pack2(-pi.z, -pi.z);
      uint2* rec = hit_rec + stream_base(T.k);
      // Every run sums into its own pair of accumulators (alternate candidates) that is folded into
      // the total when the run ends: a run's contribution then does not depend on where in shared
      // memory the run starts (a shift by one slot only swaps the two accumulators), so particles
      // in front of a run that are nobody's neighbour -- NaN positions parked in cell 0 of a row,
      // e.g. in transit through a slab -- cannot change any sum.
      float total = 0.0f;
      f32x2 sum2b = pack2(0.0f, 0.0f);
      int nw = 0, nhits = 0;
      // the 9 runs in ascending row order: (z-1: y-1, y, y+1), (z: ...), (z+1: ...)
      const int* csp = &L.cs[T.hr0 - rowstep - 1][T.lx - 1];
      const int* dlp = &L.row_delta[T.hr0 - rowstep - 1];
#if SPH_DENS_XTRIM
      // a run is ascending in x: only its candidates with |dx| < h (1.001 h: the margin covers the rounding
      // of x -+ w by orders of magnitude) can be neighbours.  Two table lookups per run give the first staged
      // slot at / above the quarter-cell threshold below x - w and the first one above x + w.
      const float w = 1.001f * P.h;
      // (index 0 also holds everything left of the table, index mt everything right of it: the upper bound is
      // at least threshold 1, the lower one at most mt)
      const int mlo = x_threshold(pi.x - w, inv_q, m0, mt), mhi = max(x_threshold(pi.x + w, inv_q, m0 - 1, mt + 1), 1);
      const ushort2* xtp = &L.xtab[T.hr0 - rowstep - 1][0];
#endif
#pragma unroll kDensUnroll
      for (int r = 0; r < 9; r++)
      {
         f32x2 sum2 = pack2(0.0f, 0.0f);
         const int delta = dlp[0];
#if SPH_DENS_XTRIM
         const int b = max(csp[0] + delta, (int)xtp[mlo].x);
         const int e = min(csp[3] + delta, (int)xtp[mhi].y);
#else
         const int b = csp[0] + delta;
         const int e = csp[3] + delta;
#endif
         const bool last_of_plane = (r == 2 || r == 5);
         csp += last_of_plane ? (rowstep - 2) * CSW : CSW;
         dlp += last_of_plane ? rowstep - 2 : 1;
#if SPH_DENS_XTRIM
         xtp += last_of_plane ? (rowstep - 2) * XT : XT;
#endif
#pragma unroll 1
         for (int c0 = b & ~3; c0 < e; c0 += 32)
         {
            const int ng = min(8, (e - c0 + 3) >> 2);
            const unsigned a0 = sbase + (unsigned)((c0 >> 2) * (GF * 4));   // the chunk's first group
            // one hit bit per candidate, shifted in at the LSB: after the chunk candidate i
            // sits at bit 31 - i.  Nested ifs instead of a loop: no bookkeeping, lanes with
            // fewer groups drop out and the warp reconverges once (a fall-through switch
            // would serialise the lanes by entry point).
            unsigned mask = 0;
#define SPH_GROUP(N) density_group<UNIT, UMASS, (N) * GF * 4>(a0, NX, NY, NZ, NH, NTHR, S2, mask, sum2, sum2b)
            SPH_GROUP(0);
            if (ng > 1)
            {
               SPH_GROUP(1);
               if (ng > 2)
               {
                  SPH_GROUP(2);
                  if (ng > 3)
                  {
                     SPH_GROUP(3);
                     if (ng > 4)
                     {
                        SPH_GROUP(4);
                        if (ng > 5)
                        {
                           SPH_GROUP(5);
                           if (ng > 6)
                           {
                              SPH_GROUP(6);
                              if (ng > 7)
                                 SPH_GROUP(7);
                           }
                        }
                     }
                  }
               }
            }
#undef SPH_GROUP
            mask <<= 32 - 4 * ng;
            // keep the candidates of the run proper, [b, e): clear the first b - c0 bits (first
            // chunk only) and everything past bit e - c0 (last chunk only; the clamped funnel
            // shift yields the top min(e - c0, 32) bits)
            mask &= (0xffffffffu >> max(b - c0, 0)) & __funnelshift_rc(0u, 0xffffffffu, e - c0);
            if (mask != 0u)
            {
               if (nw < WCAP)
                  SPH_ST_ONCE(&rec[(size_t)nw * 32], make_uint2(mask, (unsigned)(c0 - delta) | ((unsigned)r << 28)));
               nw++;
               nhits += __popc(mask);
            }
         }
         float ra, rb;
         unpack2(sum2, ra, rb);
         total += ra + rb;
      }
      SPH_ST_ONCE(&hit_info[T.k], nw <= WCAP ? ((unsigned)nw | ((unsigned)nhits << 8)) : kNoStream);
      float sa, sb;
      unpack2(sum2b, sa, sb);
      const float sum = -(total + (SPH_DENS_SPLIT_ACC ? sa + sb : 0.0f));
      // the particle itself sat in the centre run with e = -hs2: remove its own term (the
      // reference skips realIndex == particleIndex, sph.cpp:737); a NaN or infinite position has term 0
      float t_self = (pi.x - pi.x == 0.0f && pi.y - pi.y == 0.0f && pi.z - pi.z == 0.0f) ? P.hs2 : 0.0f;   // (inf - inf is NaN too)
      float self = UMASS ? (t_self * t_self) * t_self : __fmul_rn(pi.w * t_self, t_self * t_self);
      float rho = UMASS ? (P.k1 * pi.w) * (sum - self) : P.k1 * (sum - self);
      density_store(P, T.k, pi, rho, idx_sorted, vel4, s_posA4, s_velB4, s_rho);
   }
}



// One 8x8x4 tile of the density sweep.  Fast path: the layout of the whole tile is set up at once and, when its
// halo fits the staging buffer (always, outside dense clumps), staged and swept; otherwise warp 0 picks the
// sub-division level from the cell table (choose_level) and the sub-tiles are done one after the other.
template <bool UNIT, bool UMASS>
__device__ __forceinline__ void density_tile(const DevParams& P, int X0, int Y0, int Z0, TileLayout& L, int& s_level,
                                             float* sg, const float4* __restrict__ s_pos4,
                                             const uint32_t* __restrict__ cell_start,
                                             const uint32_t* __restrict__ idx_sorted, const float4* __restrict__ vel4,
                                             float4* __restrict__ s_posA4, float4* __restrict__ s_velB4,
                                             float* __restrict__ s_rho, uint2* __restrict__ hit_rec,
                                             unsigned* __restrict__ hit_info)
{
   const int cap = UMASS ? kCap : kCapMass;
   {
      const SubTile t = {X0, Y0, Z0, TBX, TBY, TBZ};
      setup_layout(P, t, cell_start, true, L);
      if (L.ntargets == 0)
         return;
      if (L.total <= cap)
      {
         stage_rows_packed<UMASS>(P, t, L, s_pos4, sg);
         __syncthreads();
         density_targets_packed<UNIT, UMASS>(P, t, L, sg, s_pos4, idx_sorted, vel4, s_posA4, s_velB4, s_rho, hit_rec,
                                             hit_info);
         return;
      }
   }
   // dense tile: sub-divide (levels 1-3) or, past that, sweep from global memory (level 4)
   __syncthreads();
   if (threadIdx.x < 32)
   {
      int level = choose_level(P, X0, Y0, Z0, cell_start, cap);
      if (threadIdx.x == 0)
         s_level = level < 1 ? 1 : level;      // (level 0 was just ruled out with the exact padded size)
   }
   __syncthreads();
   const int level = s_level;
   const int bz = (level >= 1 && level < 4) ? TBZ / 2 : TBZ;
   const int by = (level >= 2 && level < 4) ? TBY / 2 : TBY;
   const int bx = (level >= 3 && level < 4) ? TBX / 2 : TBX;
   for (int z = 0; z < TBZ; z += bz)
      for (int y = 0; y < TBY; y += by)
         for (int x = 0; x < TBX; x += bx)
         {
            SubTile t = {X0 + x, Y0 + y, Z0 + z, bx, by, bz};
            setup_layout(P, t, cell_start, level < 4, L);
            if (L.ntargets > 0)
            {
               if (level < 4)
               {
                  stage_rows_packed<UMASS>(P, t, L, s_pos4, sg);
                  __syncthreads();
                  density_targets_packed<UNIT, UMASS>(P, t, L, sg, s_pos4, idx_sorted, vel4, s_posA4, s_velB4, s_rho,
                                                      hit_rec, hit_info);
               }
               else
                  density_targets<false, UNIT, UMASS>(P, t, L, s_pos4, idx_sorted, vel4, s_posA4, s_velB4, s_rho,
                                                      hit_rec, hit_info);
            }
            __syncthreads();   // smem and layout are reused by the next sub-tile
         }
}

// the non-empty tiles of this step: one warp per tile, one lane per cell row of the tile (cell-table lookups)
__global__ void __launch_bounds__(256)
   k_tile_list(DevParams P, int tx, int ty, int tz, const uint32_t* __restrict__ cell_start,
               uint32_t* __restrict__ tile_list, int* __restrict__ tile_ctl)
{
   const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   if (i >= tx * ty * tz)
      return;
   const int x = i % tx, y = (i / tx) % ty, z = i / (tx * ty);
   const int pop = tile_population(P, x * TBX, y * TBY, z * TBZ, cell_start);
   if (pop > 0 && (threadIdx.x & 31) == 0)
      tile_list[atomicAdd(&tile_ctl[0], 1)] = (uint32_t)x | ((uint32_t)y << 10) | ((uint32_t)z << 20);
}

#ifndef SPH_DENS_PERSIST
#define SPH_DENS_PERSIST 1       // 1: persistent CTAs take the non-empty tiles from a list (k_tile_list) instead of one CTA
                                 // per tile of the grid (83 % of the 40 960 tiles of the 16.7 M dam-break are empty)
#endif

// Density sweep, persistent form: two CTAs per SM, each takes the next non-empty tile from the list until it is
// exhausted.  The order in which tiles are taken does not matter: every particle's result is its own.
template <bool UNIT, bool UMASS>
__global__ void __launch_bounds__(kTileThreads, kTileCtas)
   k_density_persist(DevParams P, const float4* __restrict__ s_pos4, const uint32_t* __restrict__ cell_start,
                     const uint32_t* __restrict__ idx_sorted, const float4* __restrict__ vel4,
                     float4* __restrict__ s_posA4, float4* __restrict__ s_velB4, float* __restrict__ s_rho,
                     uint2* __restrict__ hit_rec, unsigned* __restrict__ hit_info,
                     const uint32_t* __restrict__ tile_list, int* __restrict__ tile_ctl)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   float* sg = reinterpret_cast<float*>(smem_raw);
   __shared__ TileLayout L;
   __shared__ int s_level, s_tile;
   for (;;)
   {
      if (threadIdx.x == 0)
         s_tile = atomicAdd(&tile_ctl[1], 1);
      __syncthreads();
      const int ti = s_tile;
      if (ti >= tile_ctl[0])      // (final: k_tile_list ran before this kernel)
         return;
      const uint32_t code = __ldg(&tile_list[ti]);
      density_tile<UNIT, UMASS>(P, (int)(code & 1023u) * TBX, (int)((code >> 10) & 1023u) * TBY,
                                (int)(code >> 20) * TBZ, L, s_level, sg, s_pos4, cell_start, idx_sorted, vel4,
                                s_posA4, s_velB4, s_rho, hit_rec, hit_info);
      __syncthreads();            // the layout, the staging buffer and s_tile are reused
   }
}

template <bool UNIT, bool UMASS>
__global__ void __launch_bounds__(kTileThreads, kTileCtas)
   k_density_tiled(DevParams P, const float4* __restrict__ s_pos4, const uint32_t* __restrict__ cell_start,
                   const uint32_t* __restrict__ idx_sorted, const float4* __restrict__ vel4,
                   float4* __restrict__ s_posA4, float4* __restrict__ s_velB4, float* __restrict__ s_rho,
                   uint2* __restrict__ hit_rec, unsigned* __restrict__ hit_info)
{
   extern __shared__ __align__(16) unsigned char smem_raw[];
   float* sg = reinterpret_cast<float*>(smem_raw);
   __shared__ TileLayout L;
   __shared__ int s_level, s_pop;
   int X0, Y0, Z0;
   tile_origin(X0, Y0, Z0);
   if (threadIdx.x < 32)
   {
      int pop = tile_population(P, X0, Y0, Z0, cell_start);
      if (threadIdx.x == 0)
         s_pop = pop;
   }
   __syncthreads();
   if (s_pop == 0)
      return;
   density_tile<UNIT, UMASS>(P, X0, Y0, Z0, L, s_level, sg, s_pos4, cell_start, idx_sorted, vel4, s_posA4, s_velB4,
                             s_rho, hit_rec, hit_info);
}

// Force sweep: one thread per cell-sorted particle, no shared-memory staging and
// no block barriers before the final reduction.  The hit-mask stream of the
// density sweep names the candidates that passed the enlarged radius test; each
// lane walks the set bits of its records (MSB first = ascending cell order),
// gathers that neighbour's (x,y,z,fA) / (vx,vy,vz,fB) through L1 and applies the
// exact reference test and the pair body.  Consecutive lanes are consecutive
// particles of a cell row, so their neighbour sets overlap and the gathers hit L1.
// Particles without a stream (dense fallback tiles, record overflow, or the flat
// density kernel) scan their 27 cells instead.
template <bool UNIT_SCALE>
__global__ void __launch_bounds__(kForceThreads, SPH_FORCE_CTAS)
   k_force_stream(DevParams P, const float4* __restrict__ s_pos4, const float4* __restrict__ s_posA4,
                  const float4* __restrict__ s_velB4, const float* __restrict__ s_rho,
                  const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ cell_start,
                  const uint32_t* __restrict__ idx_sorted, const uint2* __restrict__ hit_rec,
                  const unsigned* __restrict__ hit_info, float4* __restrict__ pos4, float4* __restrict__ vel4,
                  float4* __restrict__ s_acc4, int* __restrict__ s_count, double* __restrict__ block_partials,
                  StepScalars* scal, cudaTextureObject_t tex_posA, cudaTextureObject_t tex_velB)
{
   // the first RSM records of every thread, {mask, sorted index of the chunk's first candidate}, as one
   // 8-byte word each: record w of thread t at rrec_all[w * kForceThreads + t]
   __shared__ uint2 rrec_all[RSM * kForceThreads];
   const int k = blockIdx.x * kForceThreads + threadIdx.x;
   bool active = k < sph_live_count(P);
   if (P.slab && active)
   {
      // ghost-layer particles only lend their density / velocity to owned neighbours;
      // the rank that owns them integrates them
      int cz = (int)(keys_sorted[k] / (uint32_t)(P.fx * P.fy));
      active = cz >= 2 * P.ghost_lo && cz < P.fz - 2 * P.ghost_hi;
   }
   const int kk = active ? k : 0;
   double ek = 0.0, ep = 0.0;
   unsigned long long cnt = 0;
   int cmax = -1, cmin = 0x7fffffff;
   // per-particle operands and the records: independent loads, issued together
   const unsigned info = active ? SPH_LD_ONCE(&hit_info[kk]) : 0u;
   float4 pi, vi;
   rec_load(s_posA4, s_velB4, kk, pi, vi);
   const float rho_i = SPH_LD_ONCE(&s_rho[kk]);
   const bool scan = (info & 0xffu) == kNoStream;
   const int nw = scan ? 0 : (int)(info & 0xffu);
   const int nhits = scan ? 0 : (int)(info >> 8);
   const uint2* rec = hit_rec + stream_base(kk);
#pragma unroll 4
   for (int w = 0; w < min(nw, RSM); w++)
   {
      uint2 r2 = SPH_LD_ONCE(rec + (size_t)w * 32);
      r2.y &= kBaseMask;
      rrec_all[w * kForceThreads + threadIdx.x] = r2;
   }
   ForceI I = make_force_i(P, pi, vi, rho_i);
   Vec3 pg = {0.0f, 0.0f, 0.0f}, vt = {0.0f, 0.0f, 0.0f};
   int count = 0;
   int w = 0;
   unsigned m = 0;
   int base = 0;
   const uint2* rnext = rrec_all + threadIdx.x;     // this thread's next record in shared memory
   // sorted index of this lane's next hit (records with an empty mask are never stored).  ALL_CACHED: every
   // record of every lane of the warp sits in shared memory -- the refill is then a handful of predicated
   // instructions instead of a divergent branch that some lane takes on almost every trip
   auto next_hit = [&](auto all_cached) -> int {
      if (m == 0u)
      {
         uint2 r2;
         if (decltype(all_cached)::value || w < RSM)
            r2 = *rnext;
         else
         {
            r2 = SPH_LD_ONCE(rec + (size_t)w * 32);
            r2.y &= kBaseMask;
         }
         m = r2.x;
         base = (int)r2.y;
         rnext += kForceThreads;
         w++;
      }
      int lead = __clz((int)m);
      m &= ~(0x80000000u >> lead);
      return base + lead;
   };
   // kForceIlp hits per trip: their gathers and arithmetic are independent (ILP); only the
   // accumulation is ordered (the in-loop viscosity scaling, sph.cpp:880-882).  CHECKED = false:
   // every lane of the warp still has kForceIlp hits left (no per-hit bounds test).
   auto trip = [&](int it, auto checked, auto all_cached) {
      constexpr bool CHECKED = decltype(checked)::value;
      int j[kForceIlp];
      float4 pj[kForceIlp], vj[kForceIlp];
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
      {
         if (CHECKED)
         {
            j[q] = kk;
            if (it + q < nhits)
               j[q] = next_hit(all_cached);
         }
         else
            j[q] = next_hit(all_cached);
      }
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
      {
#if SPH_FORCE_AOS
         rec_load(s_posA4, s_velB4, j[q], pj[q], vj[q]);
#else
         // SPH_FORCE_TEX 4: the two pipes share the work evenly -- even neighbours fetch (x, fA) through LSU and
         // (v, fB) through TEX, odd neighbours the other way round
         const bool swap = (SPH_FORCE_TEX & 4) && (q & 1);
         const bool tex_p = swap ? true : (SPH_FORCE_TEX & 2) != 0, tex_v = swap ? false : (SPH_FORCE_TEX & 5) != 0;
         pj[q] = tex_p ? tex1Dfetch<float4>(tex_posA, j[q]) : __ldg(&s_posA4[j[q]]);
         vj[q] = tex_v ? tex1Dfetch<float4>(tex_velB, j[q]) : __ldg(&s_velB4[j[q]]);
#endif
      }
      PairTerm t[kForceIlp];
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
         t[q] = force_term<UNIT_SCALE>(P, I, pj[q], vj[q], j[q] != kk);
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
         force_accumulate(I, t[q], pg, vt, count);
   };
   const int nmax = __reduce_max_sync(0xffffffffu, nhits);
   const int nmin = (__reduce_min_sync(0xffffffffu, nhits) / kForceIlp) * kForceIlp;
   int it = 0;
   if (__all_sync(0xffffffffu, nw <= RSM))
   {
#pragma unroll 1
      for (; it < nmin; it += kForceIlp)
         trip(it, std::false_type(), std::true_type());
#pragma unroll 1
      for (; it < nmax; it += kForceIlp)
         trip(it, std::true_type(), std::true_type());
   }
   else
   {
      // (dense scenes: more records per particle than the shared-memory cache holds)
#pragma unroll 1
      for (; it < nmin; it += kForceIlp)
         trip(it, std::false_type(), std::false_type());
#pragma unroll 1
      for (; it < nmax; it += kForceIlp)
         trip(it, std::true_type(), std::false_type());
   }
   if (scan && active)
   {
      int b[9], e[9];
      global_runs(P, keys_sorted[k], cell_start, b, e);
#pragma unroll 1
      for (int r = 0; r < 9; r++)
         for (int j = b[r]; j < e[r]; j++)
         {
            float4 pj, vj;
            rec_load(s_posA4, s_velB4, j, pj, vj);
            count += force_candidate<UNIT_SCALE>(P, I, pj, vj, j != k, pg, vt);
         }
   }
   float4 new_pos = make_float4(0.0f, 0.0f, 0.0f, 0.0f), new_vel = new_pos;
   if (active)
   {
      force_store(P, k, I, vt, pg, count, s_pos4, idx_sorted, pos4, vel4, s_acc4, s_count, ek, ep, new_pos, new_vel);
      cnt = (unsigned long long)count;
      cmax = count;
      cmin = count;
   }
   if (P.slab)
   {
      // Multi-GPU: the new position decides, right here, whether the particle goes into
      // the next halo message (boundary layer) or migrates; the messages are complete
      // when this kernel ends and the next exchange needs no pass over the particles.
      bool owned = active;
      uint32_t o = 0;
      if (k < sph_live_count(P))
      {
         o = idx_sorted[k];
         if (!active)
         {
            // ghosts and particles in transit are not integrated here: defined outputs
            // instead of whatever an earlier step left at this sorted position
            s_count[k] = 0;
            s_acc4[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
         }
         if (!active && P.slot_state[o] == SLOT_OWNED)
         {
            // an owned particle parked in a ghost layer (it crossed more than one slab in a
            // step): not integrated here, forwarded one rank per step
            new_pos = pos4[o];
            new_vel = vel4[o];
            owned = true;
         }
      }
      const unsigned char st = sph_slab_emit(P, owned, new_pos, new_vel, o);
      if (owned && st != SLOT_OWNED)
         P.slot_state[o] = st;
   }
   sph_block_reduce_scalars(ek, ep, cnt, cmax, cmin, block_partials, scal);
}


// ---- force sweep, tiled: neighbour records staged in shared memory by TMA bulk copies ----
// k_force_stream above gathers every neighbour's two 16-byte records through L1; ncu showed
// that kernel bound by the L1 data pipe (86 % of its wavefront rate: ~14 distinct 128-byte
// lines per warp-wide LDG.128).  tools/ubench_gather.cu on B200: the same fetch costs 9.4
// clocks per warp as an LDS.128 at lane-dependent slots against ~14 through L1 (the texture
// path shares the L1 tag stage: no extra throughput).  So this kernel gives a CTA an 8x4x2
// block of fine cells, copies the block's particles plus the one-cell halo -- 24 x-contiguous
// row segments of (x,y,z,fA) and of (vx,vy,vz,fB), 32 bytes per particle -- into shared
// memory with one cp.async.bulk per row and array (no registers, no transposition; an
// mbarrier counts the bytes), and every lane walks the set bits of its particle's hit-mask
// records against shared memory.  A record names its x-run (bits 28-31 of the base), which
// gives the halo row and with it the displacement global sorted index -> staged slot.
// Blocks whose halo does not fit (dense clumps) walk the same records against global
// memory; particles without a stream scan their 27 cells as in k_force_stream.
#ifndef SPH_FTX
#define SPH_FTX 8
#endif
#ifndef SPH_FTY
#define SPH_FTY 4
#endif
#ifndef SPH_FTZ
#define SPH_FTZ 2
#endif
#ifndef SPH_FCAP
#define SPH_FCAP 2944
#endif
#ifndef SPH_FTHREADS
#define SPH_FTHREADS 320
#endif
#ifndef SPH_FCTAS
#define SPH_FCTAS 2
#endif
constexpr int FTX = SPH_FTX, FTY = SPH_FTY, FTZ = SPH_FTZ;
constexpr int FHROWS = (FTY + 2) * (FTZ + 2);   // halo rows of a force tile (<= 32: one warp lays them out)
constexpr int FTROWS = FTY * FTZ;
constexpr int kFCap = SPH_FCAP;                 // staged particles per tile, 32 bytes each
constexpr int kFThreads = SPH_FTHREADS;
static_assert(FHROWS <= 32, "one warp computes the layout of a force tile");

struct ForceLayout
{
   int row_delta[FHROWS];     // staged slot = sorted index + row_delta (0 when the tile is not staged)
   int tgt_off[FTROWS + 1];   // prefix of the target counts per target row
   int tgt_k0[FTROWS];        // sorted index of the first target of the row
   int ntargets, staged;
   unsigned long long mbar;
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                "r"(bytes)
                : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
   const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
   asm volatile("{\n"
                ".reg .pred p;\n"
                "WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DONE_%=;\n"
                "bra WAIT_%=;\n"
                "DONE_%=:\n"
                "}" ::"r"(a),
                "r"(parity)
                : "memory");
}

// 1-D TMA bulk copy global -> shared; completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst_smem)),
                "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                : "memory");
}

// the hit walk of one target against `A` / `B` (shared memory when STAGED, else the global
// arrays with a zero displacement).  Same visiting order and arithmetic as k_force_stream.
template <bool UNIT_SCALE, bool STAGED>
__device__ __forceinline__ void force_walk(const DevParams& P, const ForceI& I, const ForceLayout& L, int hrb,
                                           int self_slot, const float4* A, const float4* B, const uint2* rec, int nw,
                                           int nhits, Vec3& pg, Vec3& vt, int& count)
{
   int w = 0;
   unsigned m = 0;
   int base = 0;
   // the next record is fetched one ahead of its use (its latency hides behind ~3 neighbours)
   uint2 nxt = nw > 0 ? SPH_LD_ONCE(rec) : make_uint2(0u, 0u);
   auto next_hit = [&]() -> int {
      if (m == 0u)
      {
         m = nxt.x;
         const unsigned run = nxt.y >> kRunShift;
         base = (int)(nxt.y & kBaseMask);
         if (STAGED)
         {
            const unsigned rz = (run * 11u) >> 5;   // run / 3, run < 9
            base += L.row_delta[hrb + (int)(rz * (FTY + 2) + (run - 3u * rz))];
         }
         w++;
         if (w < nw)
            nxt = SPH_LD_ONCE(rec + (size_t)w * 32);
      }
      const int lead = __clz((int)m);
      m &= ~(0x80000000u >> lead);
      return base + lead;
   };
   const int nmax = __reduce_max_sync(0xffffffffu, nhits);
#pragma unroll 1
   for (int it = 0; it < nmax; it += kForceIlp)
   {
      int j[kForceIlp];
      float4 pj[kForceIlp], vj[kForceIlp];
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
      {
         j[q] = self_slot;
         if (it + q < nhits)
            j[q] = next_hit();
      }
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
      {
         pj[q] = STAGED ? A[j[q]] : __ldg(&A[j[q]]);
         vj[q] = STAGED ? B[j[q]] : __ldg(&B[j[q]]);
      }
      PairTerm t[kForceIlp];
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
         t[q] = force_term<UNIT_SCALE>(P, I, pj[q], vj[q], j[q] != self_slot);
#pragma unroll
      for (int q = 0; q < kForceIlp; q++)
         force_accumulate(I, t[q], pg, vt, count);
   }
}

template <bool UNIT_SCALE>
__global__ void __launch_bounds__(kFThreads, SPH_FCTAS)
   k_force_tiled(DevParams P, const float4* __restrict__ s_pos4, const float4* __restrict__ s_posA4,
                 const float4* __restrict__ s_velB4, const float* __restrict__ s_rho,
                 const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ cell_start,
                 const uint32_t* __restrict__ idx_sorted, const uint2* __restrict__ hit_rec,
                 const unsigned* __restrict__ hit_info, float4* __restrict__ pos4, float4* __restrict__ vel4,
                 float4* __restrict__ s_acc4, int* __restrict__ s_count, double* __restrict__ block_partials,
                 StepScalars* scal)
{
   extern __shared__ __align__(128) unsigned char smem_raw[];
   float4* sA = reinterpret_cast<float4*>(smem_raw);
   float4* sB = sA + kFCap;
   __shared__ ForceLayout L;
   const int X0 = blockIdx.x * FTX, Y0 = blockIdx.y * FTY, Z0 = blockIdx.z * FTZ;
   const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
   const int lane = threadIdx.x & 31;
   if (threadIdx.x < 32)
   {
      if (lane == 0)
         mbar_init(&L.mbar, 1);
      // one lane per halo row: its segment [g0, g1) of the sorted arrays and, for the rows of
      // the block itself, the targets [t0, t1)
      const int hy = lane % (FTY + 2), hz = lane / (FTY + 2);
      const int y = Y0 - 1 + hy, z = Z0 - 1 + hz;
      int g0 = 0, g1 = 0, t0 = 0, t1 = 0;
      if (lane < FHROWS && y >= 0 && y < P.fy && z >= 0 && z < P.fz)
      {
         const uint32_t* row = cell_start + (size_t)(z * P.fy + y) * P.fx;
         g0 = (int)__ldg(row + max(X0 - 1, 0));
         g1 = (int)__ldg(row + min(X0 + FTX + 1, P.fx));
         if (hy >= 1 && hy <= FTY && hz >= 1 && hz <= FTZ)
         {
            t0 = (int)__ldg(row + X0);
            t1 = (int)__ldg(row + min(X0 + FTX, P.fx));
         }
      }
      const int len = g1 - g0, tc = t1 - t0;
      int incl = len, tincl = tc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
         const int up = __shfl_up_sync(0xffffffffu, incl, o), tup = __shfl_up_sync(0xffffffffu, tincl, o);
         if (lane >= o)
         {
            incl += up;
            tincl += tup;
         }
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31), ntargets = __shfl_sync(0xffffffffu, tincl, 31);
      const bool staged = total <= kFCap;
      if (lane < FHROWS)
      {
         L.row_delta[lane] = staged ? (incl - len) - g0 : 0;
         if (hy >= 1 && hy <= FTY && hz >= 1 && hz <= FTZ)
         {
            const int r = (hz - 1) * FTY + (hy - 1);   // target rows ascend with the halo row number
            L.tgt_off[r] = tincl - tc;
            L.tgt_k0[r] = t0;
         }
      }
      if (lane == 0)
      {
         L.tgt_off[FTROWS] = ntargets;
         L.ntargets = ntargets;
         L.staged = staged ? 1 : 0;
      }
      if (staged && ntargets > 0)
      {
         __syncwarp();
         if (lane == 0)
            mbar_expect_tx(&L.mbar, (unsigned)total * 32u);
         __syncwarp();
         if (len > 0)
         {
            bulk_g2s(sA + (incl - len), s_posA4 + g0, (unsigned)len * 16u, &L.mbar);
            bulk_g2s(sB + (incl - len), s_velB4 + g0, (unsigned)len * 16u, &L.mbar);
         }
      }
   }
   __syncthreads();
   const int ntargets = L.ntargets;
   if (ntargets == 0)
   {
      if (threadIdx.x == 0)
      {
         block_partials[2 * bid] = 0.0;
         block_partials[2 * bid + 1] = 0.0;
      }
      return;
   }
   const bool staged = L.staged != 0;
   double ek = 0.0, ep = 0.0;
   unsigned long long cnt = 0;
   int cmax = -1, cmin = 0x7fffffff;
   if (staged)
      mbar_wait(&L.mbar, 0);
   const int live = sph_live_count(P);
   // whole warps walk the targets so that the warp-wide votes below see every lane
   for (int tb = (int)(threadIdx.x & ~31u); tb < ntargets; tb += kFThreads)
   {
      const int tnum = tb + lane;
      const bool valid = tnum < ntargets;
      const int tn = valid ? tnum : ntargets - 1;
      int r = 0;
#pragma unroll
      for (int i = 1; i < FTROWS; i++)
         r += (tn >= L.tgt_off[i]) ? 1 : 0;
      const int k = L.tgt_k0[r] + (tn - L.tgt_off[r]);
      const int rz = r / FTY;
      const int hr0 = (rz + 1) * (FTY + 2) + (r - rz * FTY) + 1;
      bool active = valid;
      if (P.slab && active)
      {
         // ghost-layer particles only lend their density / velocity to owned neighbours;
         // the rank that owns them integrates them
         const int cz = Z0 + rz;
         active = cz >= 2 * P.ghost_lo && cz < P.fz - 2 * P.ghost_hi;
      }
      const int self_slot = k + L.row_delta[hr0];
      const unsigned info = active ? SPH_LD_ONCE(&hit_info[k]) : 0u;
      const float4 pi = staged ? sA[self_slot] : s_posA4[k];
      const float4 vi = staged ? sB[self_slot] : s_velB4[k];
      const float rho_i = SPH_LD_ONCE(&s_rho[k]);
      const bool scan = (info & 0xffu) == kNoStream;
      const int nw = scan ? 0 : (int)(info & 0xffu);
      const int nhits = scan ? 0 : (int)(info >> 8);
      const uint2* rec = hit_rec + stream_base(k);
      const ForceI I = make_force_i(P, pi, vi, rho_i);
      Vec3 pg = {0.0f, 0.0f, 0.0f}, vt = {0.0f, 0.0f, 0.0f};
      int count = 0;
      if (staged)
         force_walk<UNIT_SCALE, true>(P, I, L, hr0 - (FTY + 2) - 1, self_slot, sA, sB, rec, nw, nhits, pg, vt, count);
      else
         force_walk<UNIT_SCALE, false>(P, I, L, 0, k, s_posA4, s_velB4, rec, nw, nhits, pg, vt, count);
      if (scan && active)
      {
         int b[9], e[9];
         global_runs(P, keys_sorted[k], cell_start, b, e);
#pragma unroll 1
         for (int q = 0; q < 9; q++)
            for (int j = b[q]; j < e[q]; j++)
               count += force_candidate<UNIT_SCALE>(P, I, __ldg(&s_posA4[j]), __ldg(&s_velB4[j]), j != k, pg, vt);
      }
      float4 new_pos = make_float4(0.0f, 0.0f, 0.0f, 0.0f), new_vel = new_pos;
      if (active)
      {
         double e_k, e_p;
         force_store(P, k, I, vt, pg, count, s_pos4, idx_sorted, pos4, vel4, s_acc4, s_count, e_k, e_p, new_pos,
                     new_vel);
         ek += e_k;
         ep += e_p;
         cnt += (unsigned long long)count;
         cmax = max(cmax, count);
         cmin = min(cmin, count);
      }
      if (P.slab)
      {
         // see k_force_stream: the new position decides whether the particle goes into the
         // next halo message or migrates
         bool owned = active;
         uint32_t o = 0;
         if (valid && k < live)
         {
            o = idx_sorted[k];
            if (!active)
            {
               s_count[k] = 0;
               s_acc4[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            if (!active && P.slot_state[o] == SLOT_OWNED)
            {
               new_pos = pos4[o];
               new_vel = vel4[o];
               owned = true;
            }
         }
         const unsigned char st = sph_slab_emit(P, owned, new_pos, new_vel, o);
         if (owned && st != SLOT_OWNED)
            P.slot_state[o] = st;
      }
   }
   sph_block_reduce_scalars(ek, ep, cnt, cmax, cmin, block_partials, scal, bid);
}

size_t force_smem() { return 32 * (size_t)kFCap; }

// velocities into the force records after a density sweep that ran with defer_velocity
__global__ void __launch_bounds__(kFlatThreads)
   k_gather_vel(DevParams P, const uint32_t* __restrict__ idx_sorted, const float4* __restrict__ vel4,
                float4* __restrict__ s_velB4)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= sph_live_count(P))
      return;
   float4 v = __ldg(&vel4[idx_sorted[k]]);
#if SPH_FORCE_AOS
   float4* vb = &(reinterpret_cast<ForceRec*>(s_velB4)[k].b);   // the launcher passes the record base
#else
   float4* vb = &s_velB4[k];
#endif
   *vb = make_float4(v.x, v.y, v.z, vb->w);
}

// ---- on-demand outputs --------------------------------------------------------

// mNeighbors / mNeighborDistancesScaled of the last step (sph.h:173-174) in the
// order the force sweep visited them: ascending (fine cell, particle index).
__global__ void __launch_bounds__(kFlatThreads)
   k_build_lists(DevParams P, const float4* __restrict__ s_pos4, const uint32_t* __restrict__ keys_sorted,
                 const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ idx_sorted,
                 uint32_t* __restrict__ nbr_idx, float* __restrict__ nbr_dist, int* __restrict__ nbr_count,
                 StepScalars* scal)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= P.n)
      return;
   float4 pi = s_pos4[k];
   int b[9], e[9];
   global_runs(P, keys_sorted[k], cell_start, b, e);
   const int E = P.examine;
   uint32_t o = idx_sorted[k];
   int count = 0;
#pragma unroll
   for (int r = 0; r < 9; r++)
      for (int j = b[r]; j < e[r]; j++)
      {
         float4 pj = __ldg(&s_pos4[j]);
         float d2 = sph_dist2_exact(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z);
         if (d2 < P.h2 && j != k)
         {
            if (count < E)
            {
               nbr_idx[(size_t)o * E + count] = idx_sorted[j];
               nbr_dist[(size_t)o * E + count] = __fmul_rn(__fsqrt_rn(d2), P.scale);
            }
            count++;
         }
      }
   nbr_count[o] = count;
   if (count > E)
      atomicMax(&scal->overflow, count);
}

// The same lists from the hit-mask stream of the last step: every record is walked exactly as
// the force sweep walks it (MSB first, exact test on the survivors), so the result IS the
// sequence of neighbours the hot path visited.  Particles without a stream scan, as there.
__global__ void __launch_bounds__(kFlatThreads)
   k_lists_from_stream(DevParams P, const float4* __restrict__ s_pos4, const uint32_t* __restrict__ keys_sorted,
                       const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ idx_sorted,
                       const uint2* __restrict__ hit_rec, const unsigned* __restrict__ hit_info,
                       uint32_t* __restrict__ nbr_idx, float* __restrict__ nbr_dist, int* __restrict__ nbr_count,
                       StepScalars* scal)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= P.n)
      return;
   const float4 pi = s_pos4[k];
   const int E = P.examine;
   const uint32_t o = idx_sorted[k];
   int count = 0;
   auto visit = [&](int j) {
      float4 pj = __ldg(&s_pos4[j]);
      float d2 = sph_dist2_exact(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z);
      if (d2 < P.h2 && j != k)
      {
         if (count < E)
         {
            nbr_idx[(size_t)o * E + count] = idx_sorted[j];
            nbr_dist[(size_t)o * E + count] = __fmul_rn(__fsqrt_rn(d2), P.scale);
         }
         count++;
      }
   };
   const unsigned info = hit_info[k];
   if ((info & 0xffu) == kNoStream)
   {
      int b[9], e[9];
      global_runs(P, keys_sorted[k], cell_start, b, e);
#pragma unroll 1
      for (int r = 0; r < 9; r++)
         for (int j = b[r]; j < e[r]; j++)
            visit(j);
   }
   else
   {
      const uint2* rec = hit_rec + stream_base(k);
      const int nw = (int)(info & 0xffu);
      for (int w = 0; w < nw; w++)
      {
         uint2 r2 = rec[(size_t)w * 32];
         unsigned m = r2.x;
         const int base = (int)(r2.y & 0x0fffffffu);
         while (m)
         {
            int lead = __clz((int)m);
            m &= ~(0x80000000u >> lead);
            visit(base + lead);
         }
      }
   }
   nbr_count[o] = count;
   if (count > E)
      atomicMax(&scal->overflow, count);
}

// sorted per-particle outputs back to particle-index order (download only)
__global__ void __launch_bounds__(kFlatThreads)
   k_unsort(DevParams P, const uint32_t* __restrict__ idx_sorted, const float* __restrict__ s_rho,
            const float4* __restrict__ s_acc4, const int* __restrict__ s_count, float* __restrict__ rho,
            float4* __restrict__ acc4, int* __restrict__ nbr_count)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= sph_live_count(P))
      return;
   uint32_t o = idx_sorted[k];
   rho[o] = s_rho[k];
   acc4[o] = s_acc4[k];
   nbr_count[o] = s_count[k];
}

size_t density_smem() { return 12 * (size_t)(kCap + 4); }   // == 16 * (kCapMass + 4) rounded up
static_assert(16 * (kCapMass + 4) <= 12 * (kCap + 4), "both layouts share one shared-memory size");


}  // namespace

int sph_full_configure(sphb200_ctx* ctx)
{
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_density_tiled<true, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem()));
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_density_tiled<true, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem()));
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_density_tiled<false, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem()));
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_density_persist<true, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem()));
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_density_persist<true, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem()));
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_density_persist<false, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem()));
   {
      cudaDeviceProp prop;
      SPH_CUDA_CHECK(ctx, cudaGetDeviceProperties(&prop, ctx->device));
      ctx->sm_count = prop.multiProcessorCount;
      if (1023 * TBX < 2 * ctx->params.grid_x || 1023 * TBY < 2 * ctx->params.grid_y || 4095 * TBZ < 2 * ctx->params.grid_z)
         return sph_fail(ctx, SPHB200_E_INVALID, "FULL mode: grid too large for the tile list encoding");
      if (!ctx->tile_list)
      {
         SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->tile_list, sizeof(uint32_t) * (size_t)sph_full_tile_count(ctx)));
         SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->tile_ctl, 2 * sizeof(int)));
      }
   }
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_force_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)force_smem()));
   SPH_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_force_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)force_smem()));
   if (ctx->capacity >= (1 << 28))
      return sph_fail(ctx, SPHB200_E_INVALID, "FULL mode: at most 2^28 - 1 particle slots per GPU (hit-mask record format)");
   // hit-mask stream: WCAP records per particle, interleaved per 32 sorted particles
   size_t groups = ((size_t)ctx->capacity + 31) / 32;
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->hit_rec, sizeof(uint2) * (groups ? groups : 1) * WCAP * 32));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->hit_info, sizeof(unsigned) * (size_t)(ctx->capacity + 1)));
   if (SPH_FORCE_TEX && ctx->capacity > 0)
   {
      cudaResourceDesc rd = {};
      rd.resType = cudaResourceTypeLinear;
      rd.res.linear.desc = cudaCreateChannelDesc<float4>();
      rd.res.linear.sizeInBytes = sizeof(float4) * (size_t)ctx->capacity;
      cudaTextureDesc td = {};
      td.readMode = cudaReadModeElementType;
      rd.res.linear.devPtr = ctx->s_posA4;
      SPH_CUDA_CHECK(ctx, cudaCreateTextureObject(&ctx->tex_posA, &rd, &td, nullptr));
      rd.res.linear.devPtr = ctx->s_velB4;
      SPH_CUDA_CHECK(ctx, cudaCreateTextureObject(&ctx->tex_velB, &rd, &td, nullptr));
   }
   return SPHB200_OK;
}

// CTAs of the largest tiled launch (sizes the per-block partials of the energy reduction)
int sph_full_tile_count(const sphb200_ctx* ctx)
{
   int fx = 2 * ctx->params.grid_x, fy = 2 * ctx->params.grid_y, fz = 2 * ctx->params.grid_z;
   int dens = ((fx + TBX - 1) / TBX) * ((fy + TBY - 1) / TBY) * ((fz + TBZ - 1) / TBZ);
   int force = ((fx + FTX - 1) / FTX) * ((fy + FTY - 1) / FTY) * ((fz + FTZ - 1) / FTZ);
   return dens > force ? dens : force;
}

int sph_step_full(sphb200_ctx* ctx)
{
   // slab mode: the force sweep of this step builds the next halo messages; their number
   // (and with it the destination buffers in DevParams) is fixed first
   if (ctx->comm)
   {
      int rc0 = sph_comm_begin_step(ctx);
      if (rc0)
         return rc0;
   }
   DevParams P = sph_dev_params(ctx);
   const int n = ctx->n_local;
   cudaStream_t st = ctx->stream;
   const bool timed = ctx->params.enable_timers != 0;
   if (timed) cudaEventRecord(ctx->ev[0], st);
   int rc = sph_bin_and_sort(ctx, true);
   if (rc)
      return rc;
   if (timed) cudaEventRecord(ctx->ev[1], st);
   rc = sph_reset_scalars(ctx);
   if (rc)
      return rc;
   if (timed) cudaEventRecord(ctx->ev[2], st);   // neighbour search is fused into the two sweeps
   if (n == 0)
      return SPHB200_OK;
   const bool tiled = ctx->params.kernel_variant != 1;
   if (tiled)
   {
      dim3 tiles((P.fx + TBX - 1) / TBX, (P.fy + TBY - 1) / TBY, (P.fz + TBZ - 1) / TBZ);   // local grid
#if SPH_DENS_PERSIST
      const int ntiles = (int)(tiles.x * tiles.y * tiles.z);
      SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->tile_ctl, 0, 2 * sizeof(int), st));
      k_tile_list<<<(ntiles + 7) / 8, 256, 0, st>>>(P, (int)tiles.x, (int)tiles.y, (int)tiles.z, ctx->cell_start,
                                                        ctx->tile_list, ctx->tile_ctl);
      ctx->launches++;
      auto kd = k_density_persist<false, false>;
      if (P.scale == 1.0f)
         kd = ctx->uniform_mass ? k_density_persist<true, true> : k_density_persist<true, false>;
      const int ctas = min(ntiles, kTileCtas * ctx->sm_count);
      kd<<<ctas, kTileThreads, density_smem(), st>>>(P, ctx->s_pos4, ctx->cell_start, ctx->idx_order, ctx->vel4,
                                                    ctx->s_posA4, ctx->s_velB4, ctx->s_rho, ctx->hit_rec,
                                                    ctx->hit_info, ctx->tile_list, ctx->tile_ctl);
#else
      auto kd = k_density_tiled<false, false>;
      if (P.scale == 1.0f)
         kd = ctx->uniform_mass ? k_density_tiled<true, true> : k_density_tiled<true, false>;
      kd<<<tiles, kTileThreads, density_smem(), st>>>(P, ctx->s_pos4, ctx->cell_start, ctx->idx_order, ctx->vel4,
                                                     ctx->s_posA4, ctx->s_velB4, ctx->s_rho, ctx->hit_rec,
                                                     ctx->hit_info);
#endif
   }
   else
      k_density_flat<<<(n + kFlatThreads - 1) / kFlatThreads, kFlatThreads, 0, st>>>(
         P, ctx->s_pos4, ctx->keys_sorted, ctx->cell_start, ctx->idx_order, ctx->vel4, ctx->s_posA4, ctx->s_velB4,
         ctx->s_rho, ctx->hit_info);
   if (timed) cudaEventRecord(ctx->ev[3], st);
   if (ctx->deferred_vel_event)
   {
      // sphb200_step_host: the velocity upload ran on a second stream beside binning and the
      // density sweep; join it here
      SPH_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, ctx->deferred_vel_event, 0));
      k_gather_vel<<<(n + kFlatThreads - 1) / kFlatThreads, kFlatThreads, 0, st>>>(
         P, ctx->idx_order, ctx->vel4, SPH_FORCE_AOS ? ctx->s_posA4 : ctx->s_velB4);
      ctx->launches++;
   }
   if (timed) cudaEventRecord(ctx->ev[4], st);
   // force sweep: the flat L1-gather kernel (default), or -- kernel_variant 3 -- the tiled kernel
   // that stages the neighbour records in shared memory by TMA bulk copies (measured slower on
   // B200, profiles/r02_history.md: 5.09 vs 2.41 ms at 16.7M particles; kept for A/B)
   int blocks;
   if (ctx->params.kernel_variant == 3 && !SPH_FORCE_AOS)   // (the TMA-staged sweep copies rows of the two arrays)
   {
      dim3 ft((P.fx + FTX - 1) / FTX, (P.fy + FTY - 1) / FTY, (P.fz + FTZ - 1) / FTZ);
      blocks = (int)(ft.x * ft.y * ft.z);
      if (P.scale == 1.0f)
         k_force_tiled<true><<<ft, kFThreads, force_smem(), st>>>(
            P, ctx->s_pos4, ctx->s_posA4, ctx->s_velB4, ctx->s_rho, ctx->keys_sorted, ctx->cell_start, ctx->idx_order,
            ctx->hit_rec, ctx->hit_info, ctx->pos4, ctx->vel4, ctx->s_acc4, ctx->s_count, ctx->d_block_partials,
            ctx->d_scalars);
      else
         k_force_tiled<false><<<ft, kFThreads, force_smem(), st>>>(
            P, ctx->s_pos4, ctx->s_posA4, ctx->s_velB4, ctx->s_rho, ctx->keys_sorted, ctx->cell_start, ctx->idx_order,
            ctx->hit_rec, ctx->hit_info, ctx->pos4, ctx->vel4, ctx->s_acc4, ctx->s_count, ctx->d_block_partials,
            ctx->d_scalars);
   }
   else
   {
      blocks = (n + kForceThreads - 1) / kForceThreads;
      if (P.scale == 1.0f)
         k_force_stream<true><<<blocks, kForceThreads, 0, st>>>(
            P, ctx->s_pos4, ctx->s_posA4, ctx->s_velB4, ctx->s_rho, ctx->keys_sorted, ctx->cell_start, ctx->idx_order,
            ctx->hit_rec, ctx->hit_info, ctx->pos4, ctx->vel4, ctx->s_acc4, ctx->s_count, ctx->d_block_partials,
            ctx->d_scalars, ctx->tex_posA, ctx->tex_velB);
      else
         k_force_stream<false><<<blocks, kForceThreads, 0, st>>>(
            P, ctx->s_pos4, ctx->s_posA4, ctx->s_velB4, ctx->s_rho, ctx->keys_sorted, ctx->cell_start, ctx->idx_order,
            ctx->hit_rec, ctx->hit_info, ctx->pos4, ctx->vel4, ctx->s_acc4, ctx->s_count, ctx->d_block_partials,
            ctx->d_scalars, ctx->tex_posA, ctx->tex_velB);
   }
   ctx->launches += 2;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   if (timed) cudaEventRecord(ctx->ev[5], st);   // integrate is fused into the force sweep
   rc = sph_finish_scalars(ctx, blocks);
   if (rc)
      return rc;
   if (timed) cudaEventRecord(ctx->ev[6], st);
   if (ctx->comm)
   {
      rc = sph_comm_end_step(ctx);
      if (rc)
         return rc;
   }
   ctx->lists_valid = false;
   ctx->snapshot_valid = true;
   ctx->stream_valid = tiled;
   ctx->unsorted_valid = false;
   return SPHB200_OK;
}

int sph_full_unsort(sphb200_ctx* ctx)
{
   const int n = ctx->n_local;
   if (n == 0 || !ctx->snapshot_valid)
      return SPHB200_OK;
   k_unsort<<<(n + kFlatThreads - 1) / kFlatThreads, kFlatThreads, 0, ctx->stream>>>(
      sph_dev_params(ctx), ctx->idx_order, ctx->s_rho, ctx->s_acc4, ctx->s_count, ctx->rho, ctx->acc4, ctx->nbr_count);
   ctx->launches++;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   ctx->unsorted_valid = true;
   return SPHB200_OK;
}

int sph_full_build_lists(sphb200_ctx* ctx, bool from_stream)
{
   if (!ctx->snapshot_valid)
      return sph_fail(ctx, SPHB200_E_INVALID, "build_neighbor_lists: no FULL-mode step snapshot (step first)");
   if (from_stream && !ctx->stream_valid)
      return sph_fail(ctx, SPHB200_E_INVALID, "build_neighbor_lists_visited: the last step left no hit-mask stream");
   if (from_stream && ctx->comm)
      return sph_fail(ctx, SPHB200_E_INVALID, "build_neighbor_lists_visited: single-GPU contexts only");
   DevParams P = sph_dev_params(ctx);
   const int n = ctx->n_local;
   if (n > 0 && from_stream)
   {
      k_lists_from_stream<<<(n + kFlatThreads - 1) / kFlatThreads, kFlatThreads, 0, ctx->stream>>>(
         P, ctx->s_pos4, ctx->keys_sorted, ctx->cell_start, ctx->idx_order, ctx->hit_rec, ctx->hit_info, ctx->nbr_idx,
         ctx->nbr_dist, ctx->nbr_count, ctx->d_scalars);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   else if (n > 0)
   {
      k_build_lists<<<(n + kFlatThreads - 1) / kFlatThreads, kFlatThreads, 0, ctx->stream>>>(
         P, ctx->s_pos4, ctx->keys_sorted, ctx->cell_start, ctx->idx_order, ctx->nbr_idx, ctx->nbr_dist,
         ctx->nbr_count, ctx->d_scalars);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   StepScalars h;
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&h, ctx->d_scalars, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   if (h.overflow > ctx->params.examine_count)
      return sph_fail(ctx, SPHB200_E_CAPACITY,
                      "build_neighbor_lists: a particle has " + std::to_string(h.overflow) +
                         " neighbours, examine_count is " + std::to_string(ctx->params.examine_count));
   ctx->lists_valid = true;
   return SPHB200_OK;
}
