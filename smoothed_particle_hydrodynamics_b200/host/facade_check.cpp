// facade_check.cpp -- GPU checks of the C++ facade's readback path (the GL
// view's contract, /root/reference/src/visualization.cpp:137-213):
//   1. getGrid()[c].count() == SPHB200_F_CELL_COUNT of the same positions, for every voxel;
//      getParticles()->mPosition == SPHB200_F_POSITION after a step (interval 0);
//   2. step() with ReadbackPositions costs < 1.2x a step() with ReadbackNone at 1 M particles
//      (the snapshot leaves over a side stream; the 16 ms throttle drops the rest);
//   3. the GUI's gravity row is live: setGravity(non-zero) changes the trajectory,
//      setGravity(0) does not (sphconfig.cpp:76-95 -> sph.cpp:1225-1228).
// Prints one line per check and exits non-zero on a failure.
//
//   facade_check [scene id, default 2 = dam-break 1 M]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "particle.h"
#include "sph.h"

namespace
{
double seconds()
{
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int failures = 0;
void report(const char* what, bool ok, const char* detail)
{
   std::printf("%s %s %s\n", ok ? "PASS" : "FAIL", what, detail);
   if (!ok)
      failures++;
}
}  // namespace

int main(int argc, char** argv)
{
   const int scene = argc > 1 ? std::atoi(argv[1]) : SPHB200_SCENE_DAMBREAK_1M;
   SphParams p;
   SphSceneLattice lat;
   if (sphb200_scene_config(scene, 40.0f, &p, &lat) != SPHB200_OK)
      return 2;
   p.enable_timers = 1;   // what SPH::SPH() sets: the facade reports the six phase times every step
   const size_t n = (size_t)p.particle_count;
   std::vector<float> pos(3 * n), vel(3 * n);
   if (sphb200_scene_generate(&lat, 0, (long long)n, pos.data(), vel.data()) != SPHB200_OK)
      return 2;
   char line[256];
   try
   {
      SPH sph(p, false, 0);
      sph.uploadState(pos.data(), vel.data(), nullptr);

      // ---- 1. the mirror and the voxel counts after a step ---------------------------
      sph.setReadback(SPH::ReadbackPositions);
      sph.setReadbackIntervalMs(0);
      sph.step();
      sph.step();
      Particle* part = sph.getParticles();
      std::vector<float> ref(3 * n);
      int rc = sphb200_download(sph.context(), SPHB200_F_POSITION, ref.data(), sizeof(float) * 3 * n);
      bool same = rc == SPHB200_OK && std::memcmp(ref.data(), part->mPosition.data(), sizeof(float) * 3 * n) == 0;
      report("mirror_positions", same, "getParticles()->mPosition == device positions after step()");
      int gx, gy, gz;
      sph.getGridCellCounts(gx, gy, gz);
      const size_t cells = (size_t)gx * gy * gz;
      QList<uint32_t>* grid = sph.getGrid();
      std::vector<int> counts(cells);
      rc = sphb200_download(sph.context(), SPHB200_F_CELL_COUNT, counts.data(), sizeof(int) * cells);
      size_t bad = 0;
      long long total = 0;
      for (size_t c = 0; c < cells; c++)
      {
         bad += grid[c].count() != counts[c];
         total += grid[c].count();
      }
      std::snprintf(line, sizeof line, "%zu voxels, %zu mismatches, sum %lld of %zu particles", cells, bad, total, n);
      report("grid_counts", rc == SPHB200_OK && bad == 0 && total == (long long)n, line);
      sph.setGridMembership(true);
      grid = sph.getGrid();
      bad = 0;
      for (size_t c = 0; c < cells; c++)
      {
         bad += grid[c].count() != counts[c];
         for (int k = 1; k < grid[c].count(); k++)
            bad += grid[c][k - 1] >= grid[c][k];      // push_back order = ascending particle index (sph.cpp:476-480)
      }
      report("grid_members", bad == 0, "membership lists: counts agree, ascending particle index inside a voxel");
      sph.setGridMembership(false);

      // ---- 2. cost of the readback path --------------------------------------------
      const int steps = 40;
      double t[3];
      const int mode[3] = {SPH::ReadbackNone, SPH::ReadbackPositions, SPH::ReadbackPositions};
      const int interval[3] = {0, 16, 0};
      for (int m = 0; m < 3; m++)
      {
         sph.setReadback((SPH::Readback)mode[m]);
         sph.setReadbackIntervalMs(interval[m]);
         for (int i = 0; i < 5; i++)
            sph.step();
         sph.getParticles();
         const double t0 = seconds();
         for (int i = 0; i < steps; i++)
            sph.step();
         sphb200_synchronize(sph.context());
         t[m] = (seconds() - t0) / steps * 1e3;
         sph.getParticles();   // drains the last snapshot
      }
      std::snprintf(line, sizeof line, "ms per step(): none %.3f, positions every 16 ms %.3f (x%.2f), every step %.3f (x%.2f)",
                    t[0], t[1], t[1] / t[0], t[2], t[2] / t[0]);
      report("readback_cost", t[1] < 1.2 * t[0], line);

      // ---- 3. the gravity row ---------------------------------------------------------
      sph.setReadback(SPH::ReadbackPositions);
      sph.setReadbackIntervalMs(0);
      sph.uploadState(pos.data(), vel.data(), nullptr);
      sph.setUniformGravity(false);
      sph.setGravity(vec3(0.0f, 0.0f, 0.0f));
      sph.step();
      std::vector<float> a(sph.getParticles()->mPosition);
      sph.uploadState(pos.data(), vel.data(), nullptr);
      sph.setGravity(vec3(0.0f, -9.8f, 0.0f));
      sph.step();
      std::vector<float> b(sph.getParticles()->mPosition);
      double dy = 0.0, dx = 0.0;
      for (size_t i = 0; i < n; i++)
      {
         dy += (double)b[3 * i + 1] - (double)a[3 * i + 1];
         dx += std::fabs((double)b[3 * i] - (double)a[3 * i]);
      }
      dy /= (double)n;
      // one step from rest: v' = g dt / 2 + ... and x' = x + v_half dt: the mean shift is -g dt^2 / 2 up to the clamp
      std::snprintf(line, sizeof line, "mean dy %.3e (expected about %.3e), sum |dx| %.3e", dy,
                    -0.5 * 9.8 * p.time_step * p.time_step, dx);
      report("gravity_row_live", dy < 0.0 && dx == 0.0, line);
   }
   catch (const std::exception& e)
   {
      std::fprintf(stderr, "%s\n", e.what());
      return 1;
   }
   return failures ? 1 : 0;
}
