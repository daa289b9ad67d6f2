// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable probe around the UNMODIFIED reference class `SPH`
// (/root/reference/src/sph.h:15-215, sph.cpp).  The reference sources are
// compiled where they lie (see oracle/Makefile) against oracle/qt_shim; this
// file only subclasses SPH to reach its `protected` members (sph.h:87-215) and
// exports them through a flat C ABI that tests/ and bench.py's cpu_baseline /
// `--impl reference` legs load with ctypes.
//
// What is the reference's and what is the harness's:
//   * ref_step()                 -> SPH::step() verbatim (sph.cpp:190-304).
//   * ref_voxelize/find_sampled/density/accel -> the reference's own phase
//     functions, looped exactly like step() loops them (sph.cpp:210-277).
//   * ref_find_full()            -> HARNESS code: an all-within-h cell search
//     (the reference has no such mode).  Its output is fed to the
//     reference's own computeDensity / computeAcceleration.
//   * ref_integrate(g, walls)    -> SPH::integrate (sph.cpp:937-1022) per
//     particle, plus the two switches the reference lacks:
//     uniform gravity (added by the harness) and wall collision (the
//     reference's dead handleBoundaryConditions, sph.cpp:1025-1148, called by
//     the harness where the upstream design called it).
#include "sph.h"
#include "particle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// moc would generate these two signal bodies (sph.h:73-84).
void SPH::updateElapsed(int, int, int, int, int, int) {}
void SPH::stepFinished() {}

extern "C" {

struct RefParams
{
   int particle_count;
   int grid_x, grid_y, grid_z;
   int examine_count;
   float h;
   float simulation_scale;
   float time_step;
   float rho0;
   float stiffness;
   float viscosity;
   float damping;
   float cfl_limit;
   float grav_constant;
   float central_mass;
   float central_pos[3];
   float softening;
   float gravity[3];
   float kernel1, kernel2, kernel3;   // read-only (derived)
   float h2, h_times2, h_times2_inv;  // read-only (derived)
   float max_x, max_y, max_z;         // read-only (derived)
};

}  // extern "C"

namespace
{

struct Probe : public SPH
{
   // canonical FULL-mode fine cell (edge h): f = 2*voxel + (orientation > h),
   // orientation exactly as sph.cpp:504-515 computes it.
   std::vector<int> fineKey;
   double phaseNs[6];
   long long neighborTotal;
   int neighborMax, neighborMin;

   void resize(int n, int gx, int gy, int gz, int examine)
   {
      // same allocations as the ctor (sph.cpp:100-113), new sizes
      delete mSrcParticles;
      delete[] mVoxelIds;
      delete[] mVoxelCoords;
      delete[] mGrid;
      delete[] mNeighbors;
      delete[] mNeighborDistancesScaled;
      mParticleCount = n;
      mGridCellsX = gx;
      mGridCellsY = gy;
      mGridCellsZ = gz;
      mGridCellCount = gx * gy * gz;
      mExamineCount = examine;
      mMaxX = mCellSize * mGridCellsX;   // sph.cpp:65-67
      mMaxY = mCellSize * mGridCellsY;
      mMaxZ = mCellSize * mGridCellsZ;
      mSrcParticles = new Particle(n);
      mVoxelIds = new int[n];
      mVoxelCoords = new vec3i[n];
      mGrid = new QList<uint32_t>[mGridCellCount];
      size_t cap = (size_t)n * (size_t)examine;
      mNeighbors = new uint32_t[cap];
      mNeighborDistancesScaled = new float[cap];
      std::memset(mNeighbors, 0, cap * sizeof(uint32_t));
      std::memset(mNeighborDistancesScaled, 0, cap * sizeof(float));
      for (int i = 0; i < n; i++)
         mSrcParticles->mMass[i] = 1.0f;   // sph.cpp:88,105-108
   }

   void setH(float h)
   {
      // the ctor's own expressions (sph.cpp:47-57, 64, 86, 93-95)
      mH = h;
      mH2 = pow(h, 2);
      mHTimes2 = h * 2.0f;
      mHTimes2Inv = 1.0f / mHTimes2;
      mHScaled = h * mSimulationScale;
      mHScaled2 = pow(h * mSimulationScale, 2);
      mHScaled6 = pow(h * mSimulationScale, 6);
      mHScaled9 = pow(h * mSimulationScale, 9);
      mCellSize = 2.0f * h;
      mMaxX = mCellSize * mGridCellsX;
      mMaxY = mCellSize * mGridCellsY;
      mMaxZ = mCellSize * mGridCellsZ;
      mKernel1Scaled = 315.0f / (64.0f * (float)(M_PI) * mHScaled9);
      mKernel2Scaled = -45.0f / ((float)(M_PI) * mHScaled6);
      mKernel3Scaled = -mKernel2Scaled;
   }

   void getParams(RefParams* p) const
   {
      p->particle_count = mParticleCount;
      p->grid_x = mGridCellsX;
      p->grid_y = mGridCellsY;
      p->grid_z = mGridCellsZ;
      p->examine_count = mExamineCount;
      p->h = mH;
      p->simulation_scale = mSimulationScale;
      p->time_step = mTimeStep;
      p->rho0 = mRho0;
      p->stiffness = mStiffness;
      p->viscosity = mViscosityScalar;
      p->damping = mDamping;
      p->cfl_limit = mCflLimit;
      p->grav_constant = mGravConstant;
      p->central_mass = mCentralMass;
      p->central_pos[0] = mCentralPos[0];
      p->central_pos[1] = mCentralPos[1];
      p->central_pos[2] = mCentralPos[2];
      p->softening = mSoftening;
      p->gravity[0] = mGravity.x;
      p->gravity[1] = mGravity.y;
      p->gravity[2] = mGravity.z;
      p->kernel1 = mKernel1Scaled;
      p->kernel2 = mKernel2Scaled;
      p->kernel3 = mKernel3Scaled;
      p->h2 = mH2;
      p->h_times2 = mHTimes2;
      p->h_times2_inv = mHTimes2Inv;
      p->max_x = mMaxX;
      p->max_y = mMaxY;
      p->max_z = mMaxZ;
   }

   void setParams(const RefParams* p)
   {
      if (p->simulation_scale != mSimulationScale)
      {
         mSimulationScale = p->simulation_scale;
         mSimulationScaleInverse = 1.0f / mSimulationScale;   // sph.cpp:49
         setH(mH);
      }
      if (p->h != mH)
         setH(p->h);
      mTimeStep = p->time_step;
      mRho0 = p->rho0;
      // through the public setters where the reference has them (sph.cpp:1225-1289)
      setStiffness(p->stiffness);
      setViscosityScalar(p->viscosity);
      setDamping(p->damping);
      setCflLimit(p->cfl_limit);
      setGravity(vec3(p->gravity[0], p->gravity[1], p->gravity[2]));
      mGravConstant = p->grav_constant;
      mCentralMass = p->central_mass;
      mCentralPos[0] = p->central_pos[0];
      mCentralPos[1] = p->central_pos[1];
      mCentralPos[2] = p->central_pos[2];
      mSoftening = p->softening;
   }

   // ---- phase loops, ordered exactly as SPH::step() orders them ------------
   void phaseVoxelize() { voxelizeParticles(); }   // sph.cpp:210

   void phaseFindSampled()                          // sph.cpp:216-231
   {
      long long total = 0;
      int mx = -1, mn = 34;
      for (int i = 0; i < mParticleCount; i++)
      {
         const vec3i& v = mVoxelCoords[i];
         findNeighbors(i, &mNeighbors[(size_t)i * mExamineCount], v.x, v.y, v.z,
                       &mNeighborDistancesScaled[(size_t)i * mExamineCount]);
         int c = mSrcParticles->mNeighborCount[i];
         total += c;
         if (c > mx) mx = c;
         if (c < mn) mn = c;
      }
      neighborTotal = total;
      neighborMax = mx;
      neighborMin = mn;
   }

   // HARNESS: FULL neighbour mode.  Candidate set = the 27 fine cells (edge h)
   // around the particle's fine cell, which all lie inside the 2x2x2 voxel
   // octant that findNeighbors picks (sph.cpp:504-556).  Test and stored
   // distance are the reference's (sph.cpp:633-641, 653, 668): d2 < mH2,
   // sqrtf(d2) * scale.  Order: ascending (fine key, x, particle index).
   // Returns the largest count; counts above capacity are truncated (the
   // return value tells the caller to enlarge examine_count).
   int phaseFindFull()
   {
      const int n = mParticleCount;
      const int fx = 2 * mGridCellsX, fy = 2 * mGridCellsY, fz = 2 * mGridCellsZ;
      const size_t cells = (size_t)fx * fy * fz;
      fineKey.resize(n);
      std::vector<int> cx(n), cy(n), cz(n);
      std::vector<uint32_t> start(cells + 1, 0);
      for (int i = 0; i < n; i++)
      {
         const float* pos = &mSrcParticles->mPosition[(size_t)i * 3];
         const vec3i& v = mVoxelCoords[i];
         float ox = pos[0] - (v.x * mHTimes2);
         float oy = pos[1] - (v.y * mHTimes2);
         float oz = pos[2] - (v.z * mHTimes2);
         cx[i] = 2 * v.x + ((ox > mH) ? 1 : 0);
         cy[i] = 2 * v.y + ((oy > mH) ? 1 : 0);
         cz[i] = 2 * v.z + ((oz > mH) ? 1 : 0);
         fineKey[i] = (cz[i] * fy + cy[i]) * fx + cx[i];
         start[(size_t)fineKey[i] + 1]++;
      }
      for (size_t c = 0; c < cells; c++)
         start[c + 1] += start[c];
      std::vector<uint32_t> members(n);
      {
         std::vector<uint32_t> fill(start.begin(), start.end() - 1);
         for (int i = 0; i < n; i++)          // ascending i inside each cell
            members[fill[fineKey[i]]++] = (uint32_t)i;
      }
      // ... then ascending (x, i) inside each cell: x through the order-preserving map of the float bits
      // (FULL-mode canonical order; makes every x-run ascending in x)
      {
         const float* P = mSrcParticles->mPosition.data();
         auto key = [P](uint32_t q) -> uint32_t {
            uint32_t b;
            const float x = P[(size_t)q * 3];
            if (x != x)
               return 0u;                       // NaN first
            memcpy(&b, &x, sizeof b);
            return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
         };
         for (size_t c = 0; c < cells; c++)
            std::stable_sort(members.begin() + start[c], members.begin() + start[c + 1],
                             [&key](uint32_t a, uint32_t b) -> bool { return key(a) < key(b); });
      }
      long long total = 0;
      int mx = -1, mn = 0x7fffffff;
      for (int i = 0; i < n; i++)
      {
         const float* pos = &mSrcParticles->mPosition[(size_t)i * 3];
         uint32_t* nb = &mNeighbors[(size_t)i * mExamineCount];
         float* nd = &mNeighborDistancesScaled[(size_t)i * mExamineCount];
         int count = 0, stored = 0;
         for (int dz = -1; dz <= 1; dz++)
         {
            int z = cz[i] + dz;
            if (z < 0 || z >= fz) continue;
            for (int dy = -1; dy <= 1; dy++)
            {
               int y = cy[i] + dy;
               if (y < 0 || y >= fy) continue;
               int x0 = std::max(cx[i] - 1, 0), x1 = std::min(cx[i] + 1, fx - 1);
               size_t c0 = ((size_t)z * fy + y) * fx + x0;
               size_t c1 = ((size_t)z * fy + y) * fx + x1;
               for (uint32_t k = start[c0]; k < start[c1 + 1]; k++)
               {
                  uint32_t q = members[k];
                  if ((int)q == i) continue;                       // sph.cpp:614-615
                  const float* pq = &mSrcParticles->mPosition[(size_t)q * 3];
                  float ddx = pos[0] - pq[0];
                  float ddy = pos[1] - pq[1];
                  float ddz = pos[2] - pq[2];
                  float dot = ddx * ddx + ddy * ddy + ddz * ddz;   // sph.cpp:641
                  if (dot < mH2)                                   // sph.cpp:653
                  {
                     if (stored < mExamineCount)
                     {
                        nb[stored] = q;
                        nd[stored] = sqrtf(dot) * mSimulationScale;   // sph.cpp:668
                        stored++;
                     }
                     count++;
                  }
               }
            }
         }
         mSrcParticles->mNeighborCount[i] = stored;
         total += stored;
         if (count > mx) mx = count;
         if (stored < mn) mn = stored;
      }
      neighborTotal = total;
      neighborMax = mx;
      neighborMin = mn;
      return mx;
   }

   void phaseDensity()                              // sph.cpp:242-249
   {
      for (int i = 0; i < mParticleCount; i++)
         computeDensity(i, &mNeighbors[(size_t)i * mExamineCount],
                        &mNeighborDistancesScaled[(size_t)i * mExamineCount]);
   }

   void phaseAccel(int useGravity)                  // sph.cpp:270-277
   {
      for (int i = 0; i < mParticleCount; i++)
      {
         computeAcceleration(i, &mNeighbors[(size_t)i * mExamineCount],
                             &mNeighborDistancesScaled[(size_t)i * mExamineCount]);
         if (useGravity)   // HARNESS switch: mGravity is inert in the reference (F7)
         {
            mSrcParticles->mAcceleration[(size_t)i * 3] += mGravity.x;
            mSrcParticles->mAcceleration[(size_t)i * 3 + 1] += mGravity.y;
            mSrcParticles->mAcceleration[(size_t)i * 3 + 2] += mGravity.z;
         }
      }
   }

   void phaseIntegrate(int useGravity, int useWalls)   // sph.cpp:285-289
   {
      mKineticEnergyTotal = 0.0f;      // step() zeroes them at sph.cpp:199-200
      mPotentialEnergyTotal = 0.0f;
      for (int i = 0; i < mParticleCount; i++)
      {
         float* P = &mSrcParticles->mPosition[(size_t)i * 3];
         float* V = &mSrcParticles->mVelocity[(size_t)i * 3];
         vec3 oldPos(P[0], P[1], P[2]);
         integrate(i);
         if (useGravity)   // second half kick of the uniform field (HARNESS)
         {
            V[0] = V[0] + (mGravity.x * mTimeStep * 0.5f);
            V[1] = V[1] + (mGravity.y * mTimeStep * 0.5f);
            V[2] = V[2] + (mGravity.z * mTimeStep * 0.5f);
         }
         if (useWalls)     // the reference's dead code, sph.cpp:1025-1148
         {
            vec3 newPos(P[0], P[1], P[2]);
            vec3 newVel(V[0], V[1], V[2]);
            handleBoundaryConditions(oldPos, &newVel, mTimeStep, &newPos);
            P[0] = newPos.x; P[1] = newPos.y; P[2] = newPos.z;
            V[0] = newVel.x; V[1] = newVel.y; V[2] = newVel.z;
         }
      }
   }

   static double nowNs()
   {
      return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(
         std::chrono::steady_clock::now().time_since_epoch()).count();
   }

   // one step with ns timers around the same five loops step() times with
   // integer-ms QElapsedTimer reads (sph.cpp:209-290)
   int stepPhased(int fullMode, int useGravity, int useWalls)
   {
      int maxCount = 0;
      double t0 = nowNs();
      phaseVoxelize();
      double t1 = nowNs();
      if (fullMode) maxCount = phaseFindFull(); else phaseFindSampled();
      double t2 = nowNs();
      phaseDensity();
      double t3 = nowNs();
      double t4 = nowNs();   // pressure loop is commented out (sph.cpp:253-263)
      phaseAccel(useGravity);
      double t5 = nowNs();
      phaseIntegrate(useGravity, useWalls);
      double t6 = nowNs();
      phaseNs[0] = t1 - t0; phaseNs[1] = t2 - t1; phaseNs[2] = t3 - t2;
      phaseNs[3] = t4 - t3; phaseNs[4] = t5 - t4; phaseNs[5] = t6 - t5;
      return maxCount;
   }

   Particle* particles() { return mSrcParticles; }
   int* voxelIds() { return mVoxelIds; }
   vec3i* voxelCoords() { return mVoxelCoords; }
   QList<uint32_t>* grid() { return mGrid; }
   int cells() const { return mGridCellCount; }
   int n() const { return mParticleCount; }
   int examine() const { return mExamineCount; }
   uint32_t* neighbors() { return mNeighbors; }
   float* distances() { return mNeighborDistancesScaled; }
   float ekin() const { return mKineticEnergyTotal; }
   float epot() const { return mPotentialEnergyTotal; }
   void timers(int* t) const
   {
      t[0] = timeVoxelize; t[1] = timeFindNeighbors; t[2] = timeComputeDensity;
      t[3] = timeComputePressure; t[4] = timeComputeAcceleration; t[5] = timeIntegrate;
   }
};

inline Probe* H(void* h) { return static_cast<Probe*>(h); }

}  // namespace


extern "C" {

void* ref_create() { return new Probe(); }   // SPH::SPH(): seeded sphere scene
void ref_destroy(void* h) { delete H(h); }

void ref_resize(void* h, int n, int gx, int gy, int gz, int examine) { H(h)->resize(n, gx, gy, gz, examine); }
void ref_get_params(void* h, RefParams* p) { H(h)->getParams(p); }
void ref_set_params(void* h, const RefParams* p) { H(h)->setParams(p); }

void ref_set_state(void* h, const float* pos, const float* vel, const float* mass)
{
   Particle* s = H(h)->particles();
   size_t n = (size_t)H(h)->n();
   if (pos) std::memcpy(s->mPosition.data(), pos, n * 3 * sizeof(float));
   if (vel) std::memcpy(s->mVelocity.data(), vel, n * 3 * sizeof(float));
   if (mass) std::memcpy(s->mMass.data(), mass, n * sizeof(float));
}

void ref_get_state(void* h, float* pos, float* vel, float* mass)
{
   Particle* s = H(h)->particles();
   size_t n = (size_t)H(h)->n();
   if (pos) std::memcpy(pos, s->mPosition.data(), n * 3 * sizeof(float));
   if (vel) std::memcpy(vel, s->mVelocity.data(), n * 3 * sizeof(float));
   if (mass) std::memcpy(mass, s->mMass.data(), n * sizeof(float));
}

void ref_get_density(void* h, float* dst)
{
   Particle* s = H(h)->particles();
   std::memcpy(dst, s->mDensity.data(), (size_t)H(h)->n() * sizeof(float));
}

void ref_set_density(void* h, const float* src)
{
   Particle* s = H(h)->particles();
   std::memcpy(s->mDensity.data(), src, (size_t)H(h)->n() * sizeof(float));
}

void ref_get_acceleration(void* h, float* dst)
{
   Particle* s = H(h)->particles();
   std::memcpy(dst, s->mAcceleration.data(), (size_t)H(h)->n() * 3 * sizeof(float));
}

void ref_get_neighbor_counts(void* h, int* dst)
{
   Particle* s = H(h)->particles();
   std::memcpy(dst, s->mNeighborCount.data(), (size_t)H(h)->n() * sizeof(int));
}

void ref_get_neighbors(void* h, uint32_t* idx, float* dist)
{
   size_t cap = (size_t)H(h)->n() * (size_t)H(h)->examine();
   if (idx) std::memcpy(idx, H(h)->neighbors(), cap * sizeof(uint32_t));
   if (dist) std::memcpy(dist, H(h)->distances(), cap * sizeof(float));
}

// overwrite the neighbour table (lets a test feed computeDensity /
// computeAcceleration a list produced elsewhere)
void ref_set_neighbors(void* h, const uint32_t* idx, const float* dist, const int* counts)
{
   size_t n = (size_t)H(h)->n();
   size_t cap = n * (size_t)H(h)->examine();
   std::memcpy(H(h)->neighbors(), idx, cap * sizeof(uint32_t));
   std::memcpy(H(h)->distances(), dist, cap * sizeof(float));
   std::memcpy(H(h)->particles()->mNeighborCount.data(), counts, n * sizeof(int));
}

void ref_get_voxels(void* h, int* ids, int* coords_xyz)
{
   size_t n = (size_t)H(h)->n();
   if (ids) std::memcpy(ids, H(h)->voxelIds(), n * sizeof(int));
   if (coords_xyz) std::memcpy(coords_xyz, H(h)->voxelCoords(), n * 3 * sizeof(int));
}

// per-voxel membership as CSR: start[cells+1], members[n] (mGrid[c] in order)
void ref_get_grid(void* h, int* start, uint32_t* members)
{
   Probe* p = H(h);
   QList<uint32_t>* g = p->grid();
   int cells = p->cells();
   int k = 0;
   for (int c = 0; c < cells; c++)
   {
      start[c] = k;
      for (int j = 0; j < g[c].length(); j++)
         members[k++] = g[c][j];
   }
   start[cells] = k;
}

void ref_get_fine_keys(void* h, int* dst)
{
   Probe* p = H(h);
   std::memcpy(dst, p->fineKey.data(), p->fineKey.size() * sizeof(int));
}

void ref_get_energies(void* h, float* ekin, float* epot)
{
   *ekin = H(h)->ekin();
   *epot = H(h)->epot();
}

void ref_get_neighbor_stats(void* h, long long* total, int* mx, int* mn)
{
   *total = H(h)->neighborTotal;
   *mx = H(h)->neighborMax;
   *mn = H(h)->neighborMin;
}

void ref_get_timers_ms(void* h, int* t6) { H(h)->timers(t6); }
void ref_get_phase_ns(void* h, double* t6) { std::memcpy(t6, H(h)->phaseNs, sizeof(double) * 6); }

// the reference's own step (writes out/neighbors.txt when ./out exists, sph.cpp:203,232)
void ref_step(void* h) { H(h)->step(); }
// the reference's own run loop (sph.cpp:149-187): totalSteps+1 steps + log files
void ref_run(void* h) { H(h)->run(); }

void ref_voxelize(void* h) { H(h)->phaseVoxelize(); }
void ref_find_sampled(void* h) { H(h)->phaseFindSampled(); }
int ref_find_full(void* h) { return H(h)->phaseFindFull(); }
void ref_density(void* h) { H(h)->phaseDensity(); }
void ref_accel(void* h, int use_gravity) { H(h)->phaseAccel(use_gravity); }
void ref_integrate(void* h, int use_gravity, int use_walls) { H(h)->phaseIntegrate(use_gravity, use_walls); }
int ref_step_phased(void* h, int full_mode, int use_gravity, int use_walls)
{
   return H(h)->stepPhased(full_mode, use_gravity, use_walls);
}

}  // extern "C"
