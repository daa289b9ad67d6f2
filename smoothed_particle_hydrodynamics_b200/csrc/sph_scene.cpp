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

}  // extern "C"
