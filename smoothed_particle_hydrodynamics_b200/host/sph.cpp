// sph.cpp -- the SPH facade: reference interface (src/sph.h:20-84) over the C ABI.
#include "sph.h"

#include <sys/stat.h>
#include <sys/types.h>

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
   Particle* p = mSrcParticles;
   const size_t n = (size_t)mParams.particle_count;
   check(sphb200_download(mCtx, SPHB200_F_POSITION, p->mPosition.data(), sizeof(float) * 3 * n), "download position");
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

// One time step (reference: sph.cpp:190-304).  All five phases run on the GPU;
// afterwards the six phase times, the energies and the neighbour statistics are
// fetched (a few scalars) and the host mirror is refreshed per the readback policy.
void SPH::step()
{
   check(sphb200_step(mCtx, 1), "sphb200_step");
   float ms[6];
   check(sphb200_get_timings(mCtx, ms), "sphb200_get_timings");
   // the reference truncates nanoseconds to whole milliseconds (sph.cpp:211, 233, ...)
   timeVoxelize = (int)ms[0];
   timeFindNeighbors = (int)ms[1];
   timeComputeDensity = (int)ms[2];
   timeComputePressure = (int)ms[3];
   timeComputeAcceleration = (int)ms[4];
   timeIntegrate = (int)ms[5];
   check(sphb200_get_energies(mCtx, &mKineticEnergyTotal, &mPotentialEnergyTotal), "sphb200_get_energies");
   check(sphb200_get_neighbor_stats(mCtx, &mNeighborTotal, &mNeighborMax, &mNeighborMin),
         "sphb200_get_neighbor_stats");
   refreshParticles(mReadback);
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
}

Particle* SPH::getParticles() { return mSrcParticles; }
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

// Per-voxel membership lists of the last binning (reference: QList<uint32_t> mGrid,
// sph.h:172).  Rebuilt from the device on every call: the GL view calls it once per
// frame (visualization.cpp:180) and only needs count().
SphCellList* SPH::getGrid()
{
   const int cells = mDerived.grid_cell_count;
   const size_t n = (size_t)mParams.particle_count;
   mGridStart.resize((size_t)cells + 1);
   mGridMembers.resize(n ? n : 1);
   check(sphb200_download(mCtx, SPHB200_F_GRID_START, mGridStart.data(), sizeof(int) * ((size_t)cells + 1)),
         "download grid start");
   check(sphb200_download(mCtx, SPHB200_F_GRID_MEMBERS, mGridMembers.data(), sizeof(uint32_t) * n),
         "download grid members");
   mGrid.resize((size_t)cells);
   for (int c = 0; c < cells; c++)
   {
#ifdef SPHB200_WITH_QT
      mGrid[c].clear();
      for (int k = mGridStart[c]; k < mGridStart[c + 1]; k++)
         mGrid[c].push_back(mGridMembers[k]);
#else
      mGrid[c].assign(mGridMembers.data() + mGridStart[c], mGridStart[c + 1] - mGridStart[c]);
#endif
   }
   return mGrid.data();
}

void SPH::pushParams()
{
   check(sphb200_set_params(mCtx, &mParams), "sphb200_set_params");
   check(sphb200_get_derived(mCtx, &mDerived), "sphb200_get_derived");
}

vec3 SPH::getGravity() const { return vec3(mParams.gravity[0], mParams.gravity[1], mParams.gravity[2]); }

void SPH::setGravity(const vec3& g)
{
   mParams.gravity[0] = g.x;
   mParams.gravity[1] = g.y;
   mParams.gravity[2] = g.z;
   pushParams();
}

float SPH::getStiffness() const { return mParams.stiffness; }
void SPH::setStiffness(float v) { mParams.stiffness = v; pushParams(); }
float SPH::getViscosityScalar() const { return mParams.viscosity; }
void SPH::setViscosityScalar(float v) { mParams.viscosity = v; pushParams(); }
float SPH::getTimeStep() const { return mParams.time_step; }
void SPH::setTimeStep(float v) { mParams.time_step = v; pushParams(); }
float SPH::getDamping() const { return mParams.damping; }
void SPH::setDamping(float v) { mParams.damping = v; pushParams(); }
float SPH::getCflLimit() const { return mParams.cfl_limit; }
void SPH::setCflLimit(float v) { mParams.cfl_limit = v; pushParams(); }   // cfl^2 is re-derived (sph.cpp:1237-1241)
