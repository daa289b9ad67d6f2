// sph_scene.cpp -- host-side scene generators behind the C ABI.
//
//  * sphb200_scene_sphere : the reference constructor's scene,
//    initParticlePolitionsSphere (sph.cpp:361-425): srand(42), rejection-sampled
//    points in a radius-2 sphere at the box centre (3 rand() per try), tangential
//    velocity 20*(r + h/2)^-1/2 about the y axis, one more rand() for v_y.
//    glibc rand() is part of the contract, so this stays on the host.
//  * sphb200_scene_lattice: the jittered cubic lattice of the throughput
//    configs (SURVEY 8(d) scene rule).  Counter-based hash => any id range can
//    be generated independently (each slab rank makes its own particles).
// Built with -ffp-contract=off: one rounding per operation so the numpy
// restatement in oracle/scenes.py is bit-identical.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sphb200.h"

extern "C" {

int sphb200_scene_sphere(const SphParams* p, float* pos_xyz, float* vel_xyz)
{
   if (!p || !pos_xyz || !vel_xyz)
      return SPHB200_E_INVALID;
   SphDerived d;
   int rc = sphb200_derive(p, &d);
   if (rc)
      return rc;
   srand(42);
   const float cx = d.max_x * 0.5f, cy = d.max_y * 0.5f, cz = d.max_z * 0.5f;
   const float radius = 2.0f;
   for (int i = 0; i < p->particle_count; i++)
   {
      float x, y, z, dist;
      do
      {
         x = rand() / (float)RAND_MAX;
         y = rand() / (float)RAND_MAX;
         z = rand() / (float)RAND_MAX;
         x *= p->grid_x * d.h_times2;
         y *= p->grid_y * d.h_times2;
         z *= p->grid_z * d.h_times2;
         if (x == (float)p->grid_x) x -= 0.00001f;
         if (y == (float)p->grid_y) y -= 0.00001f;
         if (z == (float)p->grid_z) z -= 0.00001f;
         float ex = x - cx, ey = y - cy, ez = z - cz;
         dist = sqrtf(ex * ex + ey * ey + ez * ez);
      } while (dist > radius);
      pos_xyz[3 * (size_t)i] = x;
      pos_xyz[3 * (size_t)i + 1] = y;
      pos_xyz[3 * (size_t)i + 2] = z;
      float phi = atan2f(z - cz, x - cx);
      double amp = 20.0 * pow((double)dist + (double)d.h_scaled * 0.5, -0.5);
      vel_xyz[3 * (size_t)i] = (float)(amp * (double)(-sinf(phi)));
      vel_xyz[3 * (size_t)i + 2] = (float)(amp * (double)cosf(phi));
      vel_xyz[3 * (size_t)i + 1] = ((rand() / (float)RAND_MAX) * 0.5f) - 0.25f;
   }
   return SPHB200_OK;
}

static inline float hash01(uint32_t v, uint32_t seed)
{
   uint32_t x = v ^ (seed * 0x9E3779B9u);
   x ^= x >> 16;
   x *= 0x85EBCA6Bu;
   x ^= x >> 13;
   x *= 0xC2B2AE35u;
   x ^= x >> 16;
   return (float)(x >> 8) * (1.0f / 16777216.0f);
}

int sphb200_scene_lattice(int nx, int ny, int nz, float spacing, const float origin[3], uint32_t seed,
                          long long first_id, long long count, float* pos_xyz)
{
   if (nx < 1 || ny < 1 || nz < 1 || !origin || !pos_xyz || first_id < 0 || count < 0 ||
       first_id + count > (long long)nx * ny * nz)
      return SPHB200_E_INVALID;
   const float amp = 0.1f * spacing;
   for (long long k = 0; k < count; k++)
   {
      long long id = first_id + k;
      int site[3];
      site[0] = (int)(id % nx);
      site[1] = (int)((id / nx) % ny);
      site[2] = (int)(id / ((long long)nx * ny));
      for (int axis = 0; axis < 3; axis++)
      {
         float r = hash01((uint32_t)(id * 3 + axis), seed);
         float jit = (2.0f * r - 1.0f) * amp;
         float base = ((float)site[axis] + 0.5f) * spacing;
         pos_xyz[3 * (size_t)k + axis] = (origin[axis] + base) + jit;
      }
   }
   return SPHB200_OK;
}

namespace
{
struct SceneRow
{
   int sites[3], grid[3], origin_vox[3];
};
// SURVEY 8(d) configs 2-4 (+ two small cases for the parity tests), at nu = 40
const SceneRow kScenes[] = {
   {{32, 16, 32}, {20, 8, 8}, {0, 0, 0}},          // DAMBREAK_16K
   {{64, 32, 64}, {40, 16, 16}, {0, 0, 0}},        // DAMBREAK_128K
   {{128, 64, 128}, {80, 32, 32}, {0, 0, 0}},      // DAMBREAK_1M
   {{256, 128, 512}, {160, 64, 128}, {0, 0, 0}},   // DAMBREAK_16M
   {{256, 128, 512}, {160, 64, 128}, {49, 16, 3}}, // BOXDROP_16M
};
}  // namespace

int sphb200_scene_config(int scene, float nu, SphParams* p, SphSceneLattice* lat)
{
   if (scene < 0 || scene >= (int)(sizeof(kScenes) / sizeof(kScenes[0])) || !p || !lat || !(nu > 0.0f))
      return SPHB200_E_INVALID;
   const SceneRow& r = kScenes[scene];
   int rc = sphb200_default_params(p);
   if (rc)
      return rc;
   // d = h (4 pi / (3 nu))^(1/3) in double, rounded once; h is the nominal 0.1
   const float d = (float)(0.1 * pow(4.0 * M_PI / (3.0 * (double)nu), 1.0 / 3.0));
   memset(lat, 0, sizeof(*lat));
   lat->nx = r.sites[0];
   lat->ny = r.sites[1];
   lat->nz = r.sites[2];
   lat->spacing = d;
   lat->seed = 42u;
   const float voxel = 0.2f;
   int grid[3];
   for (int k = 0; k < 3; k++)
   {
      lat->origin[k] = (float)((double)r.origin_vox[k] * 0.2);
      // sparser lattices are taller than the nominal box: grow it (config 5)
      int need = (int)ceil((double)lat->origin[k] / 0.2 + (double)r.sites[k] * (double)d / 0.2) + 2;
      grid[k] = need > r.grid[k] ? need : r.grid[k];
   }
   (void)voxel;
   p->particle_count = r.sites[0] * r.sites[1] * r.sites[2];
   p->grid_x = grid[0];
   p->grid_y = grid[1];
   p->grid_z = grid[2];
   p->neighbor_mode = SPHB200_NEIGHBORS_FULL;
   p->use_uniform_gravity = 1;
   p->use_wall_collision = 1;
   p->central_mass = 0.0f;
   p->gravity[0] = 0.0f;
   p->gravity[1] = -9.8f;
   p->gravity[2] = 0.0f;
   p->rho0 = 1.0f / (d * d * d);          // rest density of the lattice (unit masses)
   p->examine_count = nu <= 40.0f ? 96 : (int)(nu * 1.6f) + 32;
   return SPHB200_OK;
}

int sphb200_scene_generate(const SphSceneLattice* lat, long long first_id, long long count, float* pos_xyz,
                           float* vel_xyz)
{
   if (!lat)
      return SPHB200_E_INVALID;
   int rc = sphb200_scene_lattice(lat->nx, lat->ny, lat->nz, lat->spacing, lat->origin, lat->seed, first_id, count,
                                  pos_xyz);
   if (rc == SPHB200_OK && vel_xyz)
      memset(vel_xyz, 0, sizeof(float) * 3 * (size_t)count);
   return rc;
}

}  // extern "C"
