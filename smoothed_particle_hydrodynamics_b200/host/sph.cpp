// sph.cpp -- the SPH facade: reference interface (src/sph.h:20-84) over the C ABI.
#include "sph.h"

#include <sys/stat.h>
#include <sys/types.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include "particle.h"

namespace
{
SphParams referenceDefaults()
{
   SphParams p;
   sphb200_default_params(&p);
   p.enable_timers = 1;   // the reference times every phase of every step (sph.cpp:209-290)
   return p;
}

long long nowNs()
{
   return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}
}  // namespace

SPH::SPH() { init(referenceDefaults(), true, -1); }

SPH::SPH(const SphParams& params, bool initSphereScene, int device) { init(params, initSphereScene, device); }

void SPH::check(int rc, const char* what) const
{
   if (rc != SPHB200_OK)
   {
      // the reference has no error path at all; a GPU failure cannot be ignored and
      // there is no CPU implementation to fall back to
      std::string msg = std::string("SPH: ") + what + " failed: " + sphb200_last_error(mCtx);
      std::cerr << msg << std::endl;
      throw std::runtime_error(msg);
   }
}

void SPH::init(const SphParams& params, bool sphereScene, int device)
{
   mCtx = nullptr;
   mParams = params;
   mSrcParticles = nullptr;
   mReadback = ReadbackPositions;
   mReadbackIntervalMs = 0;
   mGridMembership = false;
   mDefaultDamping = params.damping;
   mSnapshotHave = mGridHave = -1;
   mStepIndex = 0;
   mLastRequestNs = 0;
   mKineticEnergyTotal = mPotentialEnergyTotal = 0.0f;
   timeVoxelize = timeFindNeighbors = timeComputeDensity = timeComputePressure = timeComputeAcceleration =
      timeIntegrate = 0;
   mNeighborTotal = 0;
   mNeighborMax = -1;
   mNeighborMin = 34;
   mStopped = mPaused = false;
   mOutDir = "out";
   check(sphb200_derive(&mParams, &mDerived), "sphb200_derive");
   mTotalSteps = mDerived.total_steps;
   check(sphb200_create(&mParams, device, &mCtx), "sphb200_create");
   const size_t n = (size_t)mParams.particle_count;
   mSrcParticles = new Particle(n);
   for (size_t i = 0; i < n; i++)
      mSrcParticles->mMass[i] = 1.0f;   // sph.cpp:88, 105-108
   if (sphereScene)
   {
      check(sphb200_scene_sphere(&mParams, mSrcParticles->mPosition.data(), mSrcParticles->mVelocity.data()),
            "sphb200_scene_sphere");
      check(sphb200_upload_state(mCtx, mSrcParticles->mPosition.data(), mSrcParticles->mVelocity.data(),
                                 mSrcParticles->mMass.data()),
            "sphb200_upload_state");
   }
}

SPH::~SPH()
{
   stopSimulation();
   quit();
   wait();
   if (mCtx)
      sphb200_destroy(mCtx);
   delete mSrcParticles;
}

bool SPH::isStopped() const
{
   std::lock_guard<std::mutex> lock(mMutex);
   return mStopped;
}

bool SPH::isPaused() const
{
   std::lock_guard<std::mutex> lock(mMutex);
   return mPaused;
}

void SPH::pauseResume()
{
   std::lock_guard<std::mutex> lock(mMutex);
   mPaused = !mPaused;
}

void SPH::stopSimulation()
{
   std::lock_guard<std::mutex> lock(mMutex);
   mStopped = true;
}

void SPH::uploadState(const float* posXyz, const float* velXyz, const float* mass)
{
   std::lock_guard<std::mutex> ctxLock(mCtxMutex);
   check(sphb200_upload_state(mCtx, posXyz, velXyz, mass), "sphb200_upload_state");
   const size_t n = (size_t)mParams.particle_count;
   std::copy(posXyz, posXyz + 3 * n, mSrcParticles->mPosition.begin());
   std::copy(velXyz, velXyz + 3 * n, mSrcParticles->mVelocity.begin());
   if (mass)
      std::copy(mass, mass + n, mSrcParticles->mMass.begin());
}

void SPH::refreshParticles(Readback what)
{
   if (what == ReadbackNone)
      return;
   std::lock_guard<std::mutex> ctxLock(mCtxMutex);
   Particle* p = mSrcParticles;
   const size_t n = (size_t)mParams.particle_count;
   {
      // a snapshot older than this download must not overwrite it later
      std::lock_guard<std::mutex> lock(mMirrorMutex);
      check(sphb200_download(mCtx, SPHB200_F_POSITION, p->mPosition.data(), sizeof(float) * 3 * n), "download position");
      mSnapshotHave = mStepIndex;
   }
   if (what != ReadbackAll)
      return;
   check(sphb200_download(mCtx, SPHB200_F_VELOCITY, p->mVelocity.data(), sizeof(float) * 3 * n), "download velocity");
   check(sphb200_download(mCtx, SPHB200_F_MASS, p->mMass.data(), sizeof(float) * n), "download mass");
   check(sphb200_download(mCtx, SPHB200_F_DENSITY, p->mDensity.data(), sizeof(float) * n), "download density");
   check(sphb200_download(mCtx, SPHB200_F_ACCELERATION, p->mAcceleration.data(), sizeof(float) * 3 * n),
         "download acceleration");
   check(sphb200_download(mCtx, SPHB200_F_NEIGHBOR_COUNT, p->mNeighborCount.data(), sizeof(int) * n),
         "download neighbour count");
}

// One time step (reference: sph.cpp:190-304).  All five phases run on the GPU; afterwards ONE
// call and one synchronisation fetch the six phase times, the energies and the neighbour
// statistics (sphb200_get_step_report).  Positions and per-voxel counts for the GL view leave
// the device as an asynchronous snapshot that the next steps do not wait for.
void SPH::step()
{
   {
      std::lock_guard<std::mutex> ctxLock(mCtxMutex);
      check(sphb200_step(mCtx, 1), "sphb200_step");
      SphStepReport rep;
      check(sphb200_get_step_report(mCtx, &rep), "sphb200_get_step_report");
      // the reference truncates nanoseconds to whole milliseconds (sph.cpp:211, 233, ...)
      timeVoxelize = (int)rep.phase_ms[0];
      timeFindNeighbors = (int)rep.phase_ms[1];
      timeComputeDensity = (int)rep.phase_ms[2];
      timeComputePressure = (int)rep.phase_ms[3];
      timeComputeAcceleration = (int)rep.phase_ms[4];
      timeIntegrate = (int)rep.phase_ms[5];
      mKineticEnergyTotal = rep.e_kin;
      mPotentialEnergyTotal = rep.e_pot;
      mNeighborTotal = rep.nbr_total;
      mNeighborMax = rep.nbr_max;
      mNeighborMin = rep.nbr_min;
      mStepIndex = rep.step_index;
      if (mReadback == ReadbackPositions)
         requestSnapshot();
   }
   if (mReadback == ReadbackAll)
      refreshParticles(ReadbackAll);
#ifdef SPHB200_WITH_QT
   emit updateElapsed(timeVoxelize, timeFindNeighbors, timeComputeDensity, timeComputePressure,
                      timeComputeAcceleration, timeIntegrate);
   emit stepFinished();
#else
   if (updateElapsed)
      updateElapsed(timeVoxelize, timeFindNeighbors, timeComputeDensity, timeComputePressure,
                    timeComputeAcceleration, timeIntegrate);
   if (stepFinished)
      stepFinished();
#endif
}

// Throttle of the readback path: at most one snapshot per readback interval (the GL view
// repaints every 16 ms, visualization.cpp:24-33; a 16.7 M-particle snapshot is 200 MB of PCIe).
void SPH::requestSnapshot()
{
   const long long now = nowNs();
   if (mReadbackIntervalMs > 0 && mLastRequestNs != 0 && now - mLastRequestNs < 1000000LL * mReadbackIntervalMs)
      return;
   mLastRequestNs = now;
   check(sphb200_snapshot_request(mCtx, SPHB200_SNAP_POSITIONS | SPHB200_SNAP_CELL_COUNTS), "sphb200_snapshot_request");
}

// newest completed snapshot -> host mirror.  Does not touch the context's stream, so the GUI
// thread may call it while the worker is inside step().
bool SPH::pullSnapshot(bool positions, bool counts)
{
   std::lock_guard<std::mutex> lock(mMirrorMutex);
   const int wait = mReadbackIntervalMs == 0 ? 1 : 0;
   bool fresh = false;
   if (positions)
   {
      long long have = mSnapshotHave;
      std::vector<float>& pos = mSrcParticles->mPosition;
      check(sphb200_snapshot_read(mCtx, wait, pos.data(), sizeof(float) * pos.size(), nullptr, 0, &have),
            "sphb200_snapshot_read");
      fresh = have != mSnapshotHave;
      mSnapshotHave = have;
   }
   if (counts)
   {
      long long have = mGridHave;
      mGridCounts.resize((size_t)mDerived.grid_cell_count);
      check(sphb200_snapshot_read(mCtx, wait, nullptr, 0, mGridCounts.data(), sizeof(int) * mGridCounts.size(), &have),
            "sphb200_snapshot_read");
      fresh = fresh || have != mGridHave;
      mGridHave = have;
   }
   return fresh;
}

// The worker loop (reference: sph.cpp:149-187): totalSteps + 1 steps, and the four
// text logs in the reference's formats -- headers at sph.cpp:163-167, rows at
// 176-178 and 232 (neighbors.txt: integer average, max, min with min starting at 34).
void SPH::run()
{
   int made = mkdir(mOutDir.c_str(), 0777);
   std::cout << (made == 0 ? "Directory created" : "Directory already exists") << std::endl;
   std::ofstream energy((mOutDir + "/energy.txt").c_str());
   energy << "Step, Kinetic Energy, Potential Energy, Total Energy" << std::endl;
   std::ofstream momentum((mOutDir + "/angularmomentum.txt").c_str());
   momentum << "Step, Angular Momentum" << std::endl;
   std::ofstream timing((mOutDir + "/timing.txt").c_str());
   timing << "Step, Voxelize, Find Neighbors, Compute Density, Compute Pressure, Compute Acceleration, Integrate"
          << std::endl;
   std::ofstream neighbors((mOutDir + "/neighbors.txt").c_str());
   int stepCount = 0;
   while (!isStopped() && stepCount <= mTotalSteps)
   {
      if (isPaused())
         continue;
      step();
      int minSeen = mNeighborMin < 34 ? mNeighborMin : 34;
      neighbors << mNeighborTotal / (long long)mParams.particle_count << ", " << mNeighborMax << ", " << minSeen
                << std::endl;
      energy << stepCount << ", " << mKineticEnergyTotal << ", " << mPotentialEnergyTotal << ", "
             << mKineticEnergyTotal + mPotentialEnergyTotal << std::endl;
      momentum << stepCount << ", " << mAngularMomentumTotal.length() << std::endl;
      timing << stepCount << ", " << timeVoxelize << ", " << timeFindNeighbors << ", " << timeComputeDensity << ", "
             << timeComputePressure << ", " << timeComputeAcceleration << ", " << timeIntegrate << std::endl;
      stepCount++;
   }
   // snapshots are requested, not waited for: leave the mirror at the final state
   if (mReadback == ReadbackPositions)
      refreshParticles(ReadbackPositions);
}

Particle* SPH::getParticles()
{
   if (mReadback == ReadbackPositions)
      pullSnapshot(true, false);
   return mSrcParticles;
}
int SPH::getParticleCount() const { return mParams.particle_count; }

void SPH::getGridCellCounts(int& x, int& y, int& z)
{
   x = mParams.grid_x;
   y = mParams.grid_y;
   z = mParams.grid_z;
}

void SPH::getParticleBounds(float& x, float& y, float& z)
{
   x = mDerived.max_x;
   y = mDerived.max_y;
   z = mDerived.max_z;
}

float SPH::getInteractionRadius2() const { return mDerived.h_scaled2; }
float SPH::getCellSize() const { return mDerived.cell_size; }

// mGrid of the reference (QList<uint32_t> per voxel, sph.h:172).  Its one reader, drawVoxels,
// calls it once per frame and only asks count() of every voxel (visualization.cpp:178-193):
// the counts come from the newest snapshot (one int per voxel, histogrammed on the device).
// The member indices are filled only in a Qt build (a QList has no count without members) or
// after setGridMembership(true): that is a download of every particle index per call.
SphCellList* SPH::getGrid()
{
   const int cells = mDerived.grid_cell_count;
   mGrid.resize((size_t)cells);
#ifdef SPHB200_WITH_QT
   const bool members = true;
#else
   const bool members = mGridMembership;
#endif
   if (members || mReadback != ReadbackPositions)
   {
      std::lock_guard<std::mutex> ctxLock(mCtxMutex);
      const size_t n = (size_t)mParams.particle_count;
      if (members)
      {
         mGridStart.resize((size_t)cells + 1);
         mGridMembers.resize(n ? n : 1);
         check(sphb200_download(mCtx, SPHB200_F_GRID_START, mGridStart.data(), sizeof(int) * ((size_t)cells + 1)),
               "download grid start");
         check(sphb200_download(mCtx, SPHB200_F_GRID_MEMBERS, mGridMembers.data(), sizeof(uint32_t) * n),
               "download grid members");
      }
      else
      {
         mGridCounts.resize((size_t)cells);
         check(sphb200_download(mCtx, SPHB200_F_CELL_COUNT, mGridCounts.data(), sizeof(int) * (size_t)cells),
               "download cell counts");
      }
   }
   else if (!pullSnapshot(false, true) && mGridHave >= 0)
      return mGrid.data();          // nothing newer than what the lists already show
   for (int c = 0; c < cells; c++)
   {
#ifdef SPHB200_WITH_QT
      mGrid[c].clear();
      for (int k = mGridStart[c]; k < mGridStart[c + 1]; k++)
         mGrid[c].push_back(mGridMembers[k]);
#else
      if (members)
         mGrid[c].assign(mGridMembers.data() + mGridStart[c], mGridStart[c + 1] - mGridStart[c]);
      else
         mGrid[c].assign(nullptr, (size_t)c < mGridCounts.size() ? mGridCounts[c] : 0);
#endif
   }
   return mGrid.data();
}

void SPH::pushParams()
{
   std::lock_guard<std::mutex> ctxLock(mCtxMutex);
   check(sphb200_set_params(mCtx, &mParams), "sphb200_set_params");
   check(sphb200_get_derived(mCtx, &mDerived), "sphb200_get_derived");
}

vec3 SPH::getGravity() const { return vec3(mParams.gravity[0], mParams.gravity[1], mParams.gravity[2]); }

void SPH::setGravity(const vec3& g)
{
   mParams.gravity[0] = g.x;
   mParams.gravity[1] = g.y;
   mParams.gravity[2] = g.z;
   // the reference stores mGravity and never reads it (sph.cpp:1225-1228); here a
   // non-zero value makes the row live: a += g after the CFL clamp (zero = the reference)
   mParams.use_uniform_gravity = (g.x != 0.0f || g.y != 0.0f || g.z != 0.0f) ? 1 : 0;
   pushParams();
}

void SPH::setUniformGravity(bool on) { mParams.use_uniform_gravity = on ? 1 : 0; pushParams(); }
void SPH::setWallCollision(bool on) { mParams.use_wall_collision = on ? 1 : 0; pushParams(); }

float SPH::getStiffness() const { return mParams.stiffness; }
void SPH::setStiffness(float v) { mParams.stiffness = v; pushParams(); }
float SPH::getViscosityScalar() const { return mParams.viscosity; }
void SPH::setViscosityScalar(float v) { mParams.viscosity = v; pushParams(); }
float SPH::getTimeStep() const { return mParams.time_step; }
void SPH::setTimeStep(float v) { mParams.time_step = v; pushParams(); }
float SPH::getDamping() const { return mParams.damping; }
// mDamping only feeds handleBoundaryConditions, which the reference never calls (sph.cpp:
// 1025-1148).  SphConfig::writeValuesToSimulation passes the unchanged default on
// every Apply, so only a value the user actually edited switches the wall code on.
void SPH::setDamping(float v)
{
   mParams.damping = v;
   if (v != mDefaultDamping)
      mParams.use_wall_collision = 1;
   pushParams();
}
float SPH::getCflLimit() const { return mParams.cfl_limit; }
void SPH::setCflLimit(float v) { mParams.cfl_limit = v; pushParams(); }   // cfl^2 is re-derived (sph.cpp:1237-1241)
