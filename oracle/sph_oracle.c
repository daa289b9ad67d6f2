/* oracle/sph_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the per-timestep SPH pipeline of
 * DanielaCourel/smoothed_particle_hydrodynamics (src/sph.cpp).  It is the
 * checker that travels to the GPU box (the compiled reference in oracle/_ref
 * cannot be rebuilt there).  Parity status: PINNED -- tests/test_oracle.py
 * checks every function below against oracle/_ref (the unmodified reference
 * compiled in place) when that is present and against the committed fixtures
 * in tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  The product (libsphb200.so) never does.
 *
 * Every function names the reference lines it restates.  Arithmetic is FP32
 * with one rounding per operation (build with -ffp-contract=off), in the
 * reference's operation order, so integer outputs are bit-exact and FP fields
 * agree with the IEEE build of the reference to the last bit in practice.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct OracleParams
{
   /* inputs (defaults = SPH::SPH(), sph.cpp:36-118) */
   int particle_count;
   int grid_x, grid_y, grid_z;      /* voxels per axis, edge 2h          */
   int examine_count;               /* E, sph.cpp:98                      */
   float h;                         /* sph.cpp:47                         */
   float simulation_scale;          /* sph.cpp:48                         */
   float time_step;                 /* sph.cpp:70                         */
   float rho0, stiffness, viscosity, damping, cfl_limit;   /* 74-78, 89 */
   float grav_constant, central_mass;                       /* 80-81    */
   float central_pos[3];                                    /* 83-85    */
   float softening;                                         /* 86       */
   float gravity[3];                /* inert in the reference (F7)        */
   /* derived by oracle_derive() with the ctor's expressions             */
   float h2, h_times2, h_times2_inv, hs, hs2, hs6, hs9;
   float kernel1, kernel2, kernel3;
   float max_x, max_y, max_z;
   float cfl_limit2;
} OracleParams;

/* sph.cpp:47-98.  pow() is the double overload there (float,int -> double). */
void oracle_default_params(OracleParams* p)
{
   memset(p, 0, sizeof(*p));
   p->particle_count = 32 * 1024;
   p->grid_x = p->grid_y = p->grid_z = 32;
   p->examine_count = 32;
   p->h = 0.1f;
   p->simulation_scale = 1.0f;
   p->time_step = 0.001f;
   p->rho0 = 0.1f;
   p->stiffness = 0.001f;
   p->viscosity = 0.01f;
   p->damping = 0.001f;
   p->cfl_limit = 10000.0f;
   p->grav_constant = 4.3009e-3f;
   p->central_mass = 1e+5f;
   p->softening = -1.0f;            /* <0: derive as h*scale             */
   p->central_pos[0] = p->central_pos[1] = p->central_pos[2] = -1.0f;  /* <0: box centre */
}

void oracle_derive(OracleParams* p)
{
   float h = p->h;
   float s = p->simulation_scale;
   p->h2 = (float)pow((double)h, 2.0);                    /* :51 */
   p->h_times2 = h * 2.0f;                                 /* :52 */
   p->h_times2_inv = 1.0f / p->h_times2;                   /* :53 */
   p->hs = h * s;                                          /* :54 */
   p->hs2 = (float)pow((double)(h * s), 2.0);              /* :55 */
   p->hs6 = (float)pow((double)(h * s), 6.0);              /* :56 */
   p->hs9 = (float)pow((double)(h * s), 9.0);              /* :57 */
   float cell = 2.0f * h;                                  /* :64 */
   p->max_x = cell * (float)p->grid_x;                     /* :65-67 */
   p->max_y = cell * (float)p->grid_y;
   p->max_z = cell * (float)p->grid_z;
   if (p->central_pos[0] < 0.0f && p->central_pos[1] < 0.0f && p->central_pos[2] < 0.0f)
   {
      p->central_pos[0] = p->max_x * 0.5f;                 /* :83-85 */
      p->central_pos[1] = p->max_y * 0.5f;
      p->central_pos[2] = p->max_z * 0.5f;
   }
   if (p->softening < 0.0f)
      p->softening = p->hs;                                /* :86 */
   p->cfl_limit2 = p->cfl_limit * p->cfl_limit;            /* :90 */
   float pi_f = (float)3.14159265358979323846;
   p->kernel1 = 315.0f / (64.0f * pi_f * p->hs9);          /* :93 */
   p->kernel2 = -45.0f / (pi_f * p->hs6);                  /* :94 */
   p->kernel3 = -p->kernel2;                               /* :95 */
}

/* ---- scene: initParticlePolitionsSphere, sph.cpp:361-425 -----------------
 * glibc rand() after srand(42): three draws per rejection try, one more for
 * v_y.  pow/sin/cos/atan2 are evaluated in double as there (float args
 * promoted by the `mHScaled*0.5` double literal / C++ overloads on float). */
void oracle_init_sphere(const OracleParams* p, float* pos, float* vel)
{
   srand(42);
   float cx = p->max_x * 0.5f, cy = p->max_y * 0.5f, cz = p->max_z * 0.5f;
   float radius = 2.0f;
   for (int i = 0; i < p->particle_count; i++)
   {
      float x, y, z, dist;
      do
      {
         x = (float)rand() / (float)RAND_MAX;
         y = (float)rand() / (float)RAND_MAX;
         z = (float)rand() / (float)RAND_MAX;
         x *= (float)p->grid_x * p->h_times2;
         y *= (float)p->grid_y * p->h_times2;
         z *= (float)p->grid_z * p->h_times2;
         if (x == (float)p->grid_x) x -= 0.00001f;
         if (y == (float)p->grid_y) y -= 0.00001f;
         if (z == (float)p->grid_z) z -= 0.00001f;
         dist = (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz);
         dist = sqrtf(dist);
      } while (dist > radius);
      pos[3 * i] = x;
      pos[3 * i + 1] = y;
      pos[3 * i + 2] = z;
      float phi = atan2f(z - p->max_z * 0.5f, x - p->max_x * 0.5f);
      double amp = pow((double)dist + (double)p->hs * 0.5, -0.5);
      float vx = (float)((double)20.0f * amp * (double)(-sinf(phi)));
      float vz = (float)((double)20.0f * amp * (double)cosf(phi));
      float vy = (((float)rand() / (float)RAND_MAX) * 0.5f) - 0.25f;
      vel[3 * i] = vx;
      vel[3 * i + 1] = vy;
      vel[3 * i + 2] = vz;
   }
}

/* ---- A.1 binning: voxelizeParticles pass 1 + computeVoxelId ---------------
 * sph.cpp:443-473, 1151-1154.  One f32 multiply, floor, clamp. */
static int clampi(int v, int hi)
{
   if (v < 0) v = 0;
   if (v >= hi) v = hi - 1;
   return v;
}

void oracle_voxelize(const OracleParams* p, const float* pos, int* voxel_ids, int* voxel_xyz)
{
   for (int i = 0; i < p->particle_count; i++)
   {
      int vx = clampi((int)floorf(pos[3 * i] * p->h_times2_inv), p->grid_x);
      int vy = clampi((int)floorf(pos[3 * i + 1] * p->h_times2_inv), p->grid_y);
      int vz = clampi((int)floorf(pos[3 * i + 2] * p->h_times2_inv), p->grid_z);
      voxel_xyz[3 * i] = vx;
      voxel_xyz[3 * i + 1] = vy;
      voxel_xyz[3 * i + 2] = vz;
      voxel_ids[i] = (vz * p->grid_y + vy) * p->grid_x + vx;
   }
}

/* ---- A.2 membership: clearGrid + voxelizeParticles pass 2 -----------------
 * sph.cpp:429-435, 476-480.  push_back in ascending i == stable counting
 * sort.  start[cells+1], members[n]. */
void oracle_build_lists(int n, int cells, const int* keys, int* start, uint32_t* members)
{
   memset(start, 0, sizeof(int) * ((size_t)cells + 1));
   for (int i = 0; i < n; i++)
      start[keys[i] + 1]++;
   for (int c = 0; c < cells; c++)
      start[c + 1] += start[c];
   int* fill = (int*)malloc(sizeof(int) * (size_t)(cells > 0 ? cells : 1));
   memcpy(fill, start, sizeof(int) * (size_t)cells);
   for (int i = 0; i < n; i++)
      members[fill[keys[i]]++] = (uint32_t)i;
   free(fill);
}

/* FULL mode only (A.8): inside every fine cell the members are ordered by ascending
 * (x, particle index) -- x compared through the usual order-preserving map of the float
 * bits to unsigned (-0 < +0); a NaN (which is binned into cell 0 of its row) comes first.  Together with the cell
 * order (x fastest) this makes every x-run of three cells ascending in x, which is what
 * lets the CUDA sweeps cut a run down to |dx| < h before testing candidates.  The
 * reference has no FULL mode; its sampled mode keeps the push_back order (ascending index). */
static uint32_t x_order_key(float x)
{
   uint32_t b;
   if (x != x)
      return 0u;
   memcpy(&b, &x, sizeof b);
   return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

void oracle_order_cells_by_x(int cells, const int* start, uint32_t* members, const float* pos)
{
   for (int c = 0; c < cells; c++)
      for (int a = start[c] + 1; a < start[c + 1]; a++)       /* stable insertion sort: ties keep ascending index */
      {
         uint32_t q = members[a];
         uint32_t kq = x_order_key(pos[3 * (size_t)q]);
         int b = a - 1;
         while (b >= start[c] && x_order_key(pos[3 * (size_t)members[b]]) > kq)
         {
            members[b + 1] = members[b];
            b--;
         }
         members[b + 1] = q;
      }
}

/* orientation inside the voxel and the octant direction, sph.cpp:504-515 */
static void octant_signs(const OracleParams* p, const float* pos_i, const int* v, int* s)
{
   for (int k = 0; k < 3; k++)
   {
      float o = pos_i[k] - ((float)v[k] * p->h_times2);
      s[k] = (o > p->h) ? 1 : -1;
   }
}

/* ---- A.3 REFERENCE_SAMPLED neighbour search: findNeighbors ---------------
 * sph.cpp:484-692.  Deterministic sub-sampler; quirks kept on purpose:
 *  - slot 3 is assigned twice (536-543) so (0,0,sz) is never visited and
 *    slot 4 is never assigned (treated as skipped);
 *  - bounds are strict on the low side (578-582);
 *  - windows of K=8 consecutive list entries from an LCG offset (590-604),
 *    whole window dropped if any entry is out of range (609-620);
 *  - only lanes 0..3 of a window are distance tested (651-663);
 *  - stop once more than E-8 neighbours are held (679-688). */
void oracle_find_sampled(const OracleParams* p, const float* pos, const int* voxel_xyz,
                         const int* start, const uint32_t* members,
                         uint32_t* nbr, float* dist, int* count)
{
   static const int slot_mask[8][3] = {
      {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {-1, -1, -1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
   const int E = p->examine_count;
   const int G[3] = {p->grid_x, p->grid_y, p->grid_z};
   for (int i = 0; i < p->particle_count; i++)
   {
      const float* pi = &pos[3 * (size_t)i];
      const int* v = &voxel_xyz[3 * (size_t)i];
      uint32_t* out_n = &nbr[(size_t)i * E];
      float* out_d = &dist[(size_t)i * E];
      int s[3];
      octant_signs(p, pi, v, s);
      int found = 0;
      int visited = 0;     /* `almost_a_random` */
      int done = 0;
      for (int slot = 0; slot < 8 && !done; slot++)
      {
         if (slot == 4)
            continue;
         int c[3], ok = 1;
         for (int k = 0; k < 3; k++)
         {
            c[k] = v[k] + slot_mask[slot][k] * s[k];
            ok = ok && (c[k] > 0) && (c[k] < G[k]);
         }
         if (!ok)
            continue;
         int cell = (c[2] * p->grid_y + c[1]) * p->grid_x + c[0];
         int len = start[cell + 1] - start[cell];
         if (len == 0)
            continue;
         const uint32_t* list = &members[start[cell]];
         /* int overflow wraps in every build of the reference that was probed */
         int32_t lcg = (int32_t)(1664525u * (uint32_t)(i + visited) + 1013904223u);
         int off = lcg % len;          /* C truncation: sign of lcg */
         visited++;
         int dir = (i % 2) ? -1 : 1;
         int base = 0;
         int windows = (len + 7) / 8;
         for (int w = 0; w < windows; w++)
         {
            int first = off + base * dir;
            if (first < 0 || first + 7 >= len)     /* any of the 8 out of range */
               break;
            base += 8;
            for (int j = 0; j < 4; j++)
            {
               uint32_t q = list[first + j];
               if ((int)q == i)
                  continue;
               const float* pq = &pos[3 * (size_t)q];
               float dx = pi[0] - pq[0];
               float dy = pi[1] - pq[1];
               float dz = pi[2] - pq[2];
               float d2 = dx * dx + dy * dy + dz * dz;
               if (d2 < p->h2)
               {
                  out_n[found] = q;
                  out_d[found] = sqrtf(d2) * p->simulation_scale;
                  found++;
               }
            }
            if (found > E - 8)
            {
               done = 1;
               break;
            }
         }
      }
      count[i] = found;
   }
}

/* ---- A.8 FULL neighbour mode (ours; the reference has none) ---------------
 * Fine cell (edge h) = 2*voxel + (orientation > h) per axis, with the
 * reference's voxel (A.1) and orientation (sph.cpp:504-515).  Neighbours of i
 * = every j != i in the 27 fine cells around i with d2 < h2 (test and stored
 * distance as sph.cpp:633-641, 653, 668).  Order: ascending (fine key, x_j, j) -- the
 * member order oracle_order_cells_by_x leaves in fmembers.
 * count[i] is the true count; only the first E are stored. */
void oracle_fine_keys(const OracleParams* p, const float* pos, const int* voxel_xyz,
                      int* fine_xyz, int* fine_keys)
{
   int fx = 2 * p->grid_x, fy = 2 * p->grid_y;
   for (int i = 0; i < p->particle_count; i++)
   {
      int s[3];
      octant_signs(p, &pos[3 * (size_t)i], &voxel_xyz[3 * (size_t)i], s);
      int c[3];
      for (int k = 0; k < 3; k++)
      {
         c[k] = 2 * voxel_xyz[3 * (size_t)i + k] + (s[k] > 0 ? 1 : 0);
         fine_xyz[3 * (size_t)i + k] = c[k];
      }
      fine_keys[i] = (c[2] * fy + c[1]) * fx + c[0];
   }
}

void oracle_find_full(const OracleParams* p, const float* pos, const int* fine_xyz,
                      const int* fstart, const uint32_t* fmembers,
                      uint32_t* nbr, float* dist, int* count)
{
   const int E = p->examine_count;
   const int fx = 2 * p->grid_x, fy = 2 * p->grid_y, fz = 2 * p->grid_z;
   for (int i = 0; i < p->particle_count; i++)
   {
      const float* pi = &pos[3 * (size_t)i];
      const int* c = &fine_xyz[3 * (size_t)i];
      int found = 0;
      for (int dz = -1; dz <= 1; dz++)
      {
         int z = c[2] + dz;
         if (z < 0 || z >= fz) continue;
         for (int dy = -1; dy <= 1; dy++)
         {
            int y = c[1] + dy;
            if (y < 0 || y >= fy) continue;
            int x0 = c[0] - 1 < 0 ? 0 : c[0] - 1;
            int x1 = c[0] + 1 >= fx ? fx - 1 : c[0] + 1;
            int row = (z * fy + y) * fx;
            for (int k = fstart[row + x0]; k < fstart[row + x1 + 1]; k++)
            {
               uint32_t q = fmembers[k];
               if ((int)q == i) continue;
               const float* pq = &pos[3 * (size_t)q];
               float dx = pi[0] - pq[0];
               float dy2 = pi[1] - pq[1];
               float dz2 = pi[2] - pq[2];
               float d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
               if (d2 < p->h2)
               {
                  if (found < E)
                  {
                     nbr[(size_t)i * E + found] = q;
                     dist[(size_t)i * E + found] = sqrtf(d2) * p->simulation_scale;
                  }
                  found++;
               }
            }
         }
      }
      count[i] = found;
   }
}

/* ---- A.4 density: computeDensity, sph.cpp:721-766 ------------------------ */
void oracle_density(const OracleParams* p, const float* mass, const uint32_t* nbr,
                    const float* dist, const int* count, float* rho)
{
   const int E = p->examine_count;
   for (int i = 0; i < p->particle_count; i++)
   {
      float sum = 0.0f;
      int cnt = count[i] < E ? count[i] : E;
      for (int k = 0; k < cnt; k++)
      {
         uint32_t q = nbr[(size_t)i * E + k];
         if (q >= (uint32_t)p->particle_count)      /* :734 */
            break;
         if ((int)q == i)
            continue;
         float d = dist[(size_t)i * E + k];
         if (d > p->hs)                             /* :744 */
            continue;
         float t = p->hs2 - (d * d);
         t = t * t * t;
         float w = p->kernel1 * t;
         sum += mass[q] * w;
      }
      rho[i] = sum;
   }
}

/* central point-mass term shared by computeAcceleration (893-915) and
 * integrate (973-989): returns (r - c) / (|r - c| + eps)^3 per axis and
 * writes the cubed softened distance. */
static void central_term(const OracleParams* p, const float* r, float* g, float* d3_out)
{
   float rel[3];
   for (int k = 0; k < 3; k++)
      rel[k] = (r[k] - p->central_pos[k]) * p->simulation_scale;
   float dot = (rel[0] * rel[0]) + (rel[1] * rel[1]) + (rel[2] * rel[2]);
   dot = sqrtf(dot);
   float sd = dot + p->softening;
   float d3 = sd * sd * sd;
   for (int k = 0; k < 3; k++)
      g[k] = rel[k] / d3;
   *d3_out = d3;
}

/* ---- A.5 acceleration: computeAcceleration, sph.cpp:778-934 --------------
 * Quirks kept: rhoiInv is 1/p_i (pressure) when p_i>0 (785-788); pressure
 * term is a product (860); grad W divides by (d + 0.01) in double (854-856);
 * viscous accumulator is scaled by mu*rhoiInv INSIDE the loop (880-882).
 * use_gravity adds the uniform field after the CFL clamp (harness switch). */
void oracle_acceleration(const OracleParams* p, const float* pos, const float* vel,
                         const float* mass, const float* rho, const uint32_t* nbr,
                         const float* dist, const int* count, int use_gravity, float* acc)
{
   const int E = p->examine_count;
   for (int i = 0; i < p->particle_count; i++)
   {
      float pi = (rho[i] - p->rho0) * p->stiffness;
      float rhoi_inv = (pi > 0.0f) ? (1.0f / pi) : 1.0f;
      float rhoi_inv2 = rhoi_inv * rhoi_inv;
      float pi_div = pi * rhoi_inv2;
      const float* r = &pos[3 * (size_t)i];
      const float* vi = &vel[3 * (size_t)i];
      float pg[3] = {0.0f, 0.0f, 0.0f};
      float vt[3] = {0.0f, 0.0f, 0.0f};
      int cnt = count[i] < E ? count[i] : E;
      for (int k = 0; k < cnt; k++)
      {
         uint32_t q = nbr[(size_t)i * E + k];
         float rhoj = rho[q];
         float pj = (rhoj - p->rho0) * p->stiffness;
         float rhoj_inv = (rhoj > 0.0f) ? (1.0f / rhoj) : 1.0f;
         float rhoj_inv2 = rhoj_inv * rhoj_inv;
         const float* rj = &pos[3 * (size_t)q];
         const float* vj = &vel[3 * (size_t)q];
         float mj = mass[q];
         float d = dist[(size_t)i * E + k];
         float grad[3];
         for (int a = 0; a < 3; a++)
         {
            float rel = (r[a] - rj[a]) * p->simulation_scale;
            grad[a] = (float)((double)(p->kernel2 * rel) / ((double)d + 0.01));
         }
         float c = p->hs - d;
         c *= c;
         c *= mj * pi_div * (pj * rhoj_inv2);
         for (int a = 0; a < 3; a++)
            pg[a] += grad[a] * c;
         float cv = p->hs - d;
         cv *= rhoj_inv * mj * p->kernel3;
         float s = p->viscosity * rhoi_inv;
         for (int a = 0; a < 3; a++)
         {
            vt[a] += (vj[a] - vi[a]) * cv;
            vt[a] *= s;
         }
      }
      float a3[3];
      for (int a = 0; a < 3; a++)
         a3[a] = vt[a] - pg[a];
      float g[3], d3;
      central_term(p, r, g, &d3);
      for (int a = 0; a < 3; a++)
         a3[a] += -p->grav_constant * p->central_mass * g[a];
      float dot = (a3[0] * a3[0]) + (a3[1] * a3[1]) + (a3[2] * a3[2]);
      if (dot > p->cfl_limit2)                       /* :921-929 */
      {
         float len = sqrtf(dot);
         float sc = p->cfl_limit / len;
         for (int a = 0; a < 3; a++)
            a3[a] *= sc;
      }
      if (use_gravity)
         for (int a = 0; a < 3; a++)
            a3[a] += p->gravity[a];
      for (int a = 0; a < 3; a++)
         acc[3 * (size_t)i + a] = a3[a];
   }
}

/* ---- A.7 wall collision: handleBoundaryConditions + applyBoundary --------
 * sph.cpp:1025-1148 (dead code in the reference).  Axis by axis, in x,y,z
 * order, using the PRE-step position and the current new velocity. */
static void wall_axis(const OracleParams* p, const float* old_pos, float dt, int axis,
                      float wall_max, float* new_pos, float* new_vel)
{
   float n[3] = {0.0f, 0.0f, 0.0f};
   float t;
   if (new_pos[axis] < 0.0f)
   {
      n[axis] = 1.0f;
      t = -old_pos[axis] / new_vel[axis];
   }
   else if (new_pos[axis] > wall_max)
   {
      n[axis] = -1.0f;
      t = (wall_max - old_pos[axis]) / new_vel[axis];
   }
   else
      return;
   float hit[3], refl[3];
   for (int k = 0; k < 3; k++)
      hit[k] = old_pos[k] + new_vel[k] * t;
   float dot = new_vel[0] * n[0] + new_vel[1] * n[1] + new_vel[2] * n[2];
   for (int k = 0; k < 3; k++)
      refl[k] = new_vel[k] - (n[k] * dot) * 2.0f;
   float remaining = dt - t;
   float f = remaining * p->damping;
   for (int k = 0; k < 3; k++)
   {
      new_vel[k] = refl[k];
      new_pos[k] = hit[k] + refl[k] * f;
   }
}

/* ---- A.6 integration: integrate, sph.cpp:937-1022 -------------------------
 * In place, in particle order; energy sums are serial f32 like the
 * reference's members.  use_gravity / use_walls are the harness switches
 * (second half kick of the uniform field; wall reflection after integrate). */
void oracle_integrate(const OracleParams* p, float* pos, float* vel, const float* acc,
                      const float* mass, int use_gravity, int use_walls,
                      float* ekin_out, float* epot_out)
{
   float ekin = 0.0f, epot = 0.0f;
   float dt = p->time_step;
   float pos_dt = dt * (1.0f / p->simulation_scale);     /* :956, :49 */
   for (int i = 0; i < p->particle_count; i++)
   {
      float* r = &pos[3 * (size_t)i];
      float* v = &vel[3 * (size_t)i];
      const float* a = &acc[3 * (size_t)i];
      float old_pos[3] = {r[0], r[1], r[2]};
      float vh[3], nr[3], nv[3];
      for (int k = 0; k < 3; k++)
         vh[k] = v[k] + (a[k] * dt * 0.5f);
      for (int k = 0; k < 3; k++)
         nr[k] = r[k] + (vh[k] * pos_dt);
      float g[3], d3;
      central_term(p, nr, g, &d3);
      for (int k = 0; k < 3; k++)
      {
         float a2 = -p->grav_constant * p->central_mass * g[k];
         nv[k] = vh[k] + (a2 * dt);
      }
      float dot = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      if (dot > 0)                                        /* :1001 */
      {
         ekin += 0.5f * mass[i] * dot;
         epot -= p->grav_constant * p->central_mass * mass[i] / d3;
      }
      if (use_gravity)
         for (int k = 0; k < 3; k++)
            nv[k] = nv[k] + (p->gravity[k] * dt * 0.5f);
      if (use_walls)
      {
         wall_axis(p, old_pos, dt, 0, p->max_x, nr, nv);
         wall_axis(p, old_pos, dt, 1, p->max_y, nr, nv);
         wall_axis(p, old_pos, dt, 2, p->max_z, nr, nv);
      }
      for (int k = 0; k < 3; k++)
      {
         r[k] = nr[k];
         v[k] = nv[k];
      }
   }
   *ekin_out = ekin;
   *epot_out = epot;
}
