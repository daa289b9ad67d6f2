// reference_callsites.cpp -- compile-only conformance unit: every statement with
// which the reference's GUI and main() touch `class SPH`, restated against the facade
// host/sph.h in its non-Qt build.  Qt and GL are absent from this image, so the widgets are
// reduced to the few members the call sites use; the SPH expressions themselves are the
// reference's (file:line cited per block).  tests/test_abi_cpu.py compiles this file with
// `g++ -fsyntax-only`; it is never linked or run.
#include <string>

#include "particle.h"
#include "sph.h"

// ---- what the call sites need from GL / Qt ------------------------------------------------
static void glVertex3f(float, float, float) {}
static void glColor4f(float, float, float, float) {}
static void drawVoxel(float, float, float, float, float, float) {}
struct TextItem
{
   float value;
   void setNumber(float v) { value = v; }
   float toFloat() const { return value; }
};

// ---- visualization.cpp:144-157 (drawParticles) --------------------------------------------
void drawParticles(SPH* mSph)
{
   int count = mSph->getParticleCount();
   Particle* particles = mSph->getParticles();
   for (int i = 0; i < count; i++)
   {
      const float& pos_x = particles->mPosition[i * 3];
      const float& pos_y = particles->mPosition[i * 3 + 1];
      const float& pos_z = particles->mPosition[i * 3 + 2];
      glVertex3f(pos_x, pos_y, pos_z);
   }
}

// ---- visualization.cpp:175-193 (drawVoxels) -----------------------------------------------
void drawVoxels(SPH* mSph)
{
   int x = 0;
   int y = 0;
   int z = 0;
   int index = 0;
   int count = 0;
   float cellSize = mSph->getCellSize();
   mSph->getGridCellCounts(x, y, z);
   QList<uint32_t>* grid = mSph->getGrid();
   for (int zi = 0; zi < z; zi++)
      for (int yi = 0; yi < y; yi++)
         for (int xi = 0; xi < x; xi++)
         {
            index = (xi) + (yi * x) + (zi * x * y);
            count = grid[index].count();
            if (count > 0)
            {
               glColor4f(count * 0.02f, 0.0f, 0.0f, 1.0f);
               drawVoxel(xi * cellSize, (xi + 1) * cellSize, yi * cellSize, (yi + 1) * cellSize, zi * cellSize,
                         (zi + 1) * cellSize);
            }
         }
}

// ---- visualization.cpp:327-335 (paintGL) --------------------------------------------------
void paintGL(SPH* mSph)
{
   float maxX;
   float maxY;
   float maxZ;
   mSph->getParticleBounds(maxX, maxY, maxZ);
   float invScaleX = 1.0f / maxX;
   float invScaleY = 1.0f / maxY;
   float invScaleZ = 1.0f / maxZ;
   (void)invScaleX; (void)invScaleY; (void)invScaleZ;
}

// ---- sphconfig.cpp:56-95 (SphConfig round trip of the eight values) -------------------------
struct SphConfigRows
{
   SPH* mSph;
   TextItem mGravityX, mGravityY, mGravityZ, mStiffness, mViscosity, mDamping, mTimeStep, mCflLimit;

   void readValuesFromSimulation()
   {
      vec3 gravity = mSph->getGravity();
      float stiffness = mSph->getStiffness();
      float viscosity = mSph->getViscosityScalar();
      float damping = mSph->getDamping();
      float timeStep = mSph->getTimeStep();
      float cfl = mSph->getCflLimit();
      mGravityX.setNumber(gravity.x);
      mGravityY.setNumber(gravity.y);
      mGravityZ.setNumber(gravity.z);
      mStiffness.setNumber(stiffness);
      mViscosity.setNumber(viscosity);
      mDamping.setNumber(damping);
      mTimeStep.setNumber(timeStep);
      mCflLimit.setNumber(cfl);
   }

   void writeValuesToSimulation()
   {
      vec3 gravity;
      float gravityX = mGravityX.toFloat();
      float gravityY = mGravityY.toFloat();
      float gravityZ = mGravityZ.toFloat();
      gravity.set(gravityX, gravityY, gravityZ);
      float stiffness = mStiffness.toFloat();
      float viscosity = mViscosity.toFloat();
      float damping = mDamping.toFloat();
      float timestep = mTimeStep.toFloat();
      float cfl = mCflLimit.toFloat();
      mSph->setGravity(gravity);
      mSph->setStiffness(stiffness);
      mSph->setViscosityScalar(viscosity);
      mSph->setDamping(damping);
      mSph->setTimeStep(timestep);
      mSph->setCflLimit(cfl);
   }
};

// ---- widget.cpp:108-125 (the slot updateElapsed(int x6) is connected to) -------------------
struct WidgetSlots
{
   int mElapsedVoxelize, mElapsedFindNeighbors, mElapsedComputeDensity, mElapsedComputePressure,
      mElapsedComputeAcceleration, mElapsedIntegrate;
   void updateElapsedSph(int timeVoxelize, int timeFindNeighbors, int timeComputeDensity, int timeComputePressure,
                         int timeComputeAcceleration, int integrate)
   {
      mElapsedVoxelize = timeVoxelize;
      mElapsedFindNeighbors = timeFindNeighbors;
      mElapsedComputeDensity = timeComputeDensity;
      mElapsedComputePressure = timeComputePressure;
      mElapsedComputeAcceleration = timeComputeAcceleration;
      mElapsedIntegrate = integrate;
   }
};

// ---- main.cpp:20-69 -----------------------------------------------------------------------
int referenceMain(int argc, char* argv[])
{
   SPH sph;

   // run simulation whitout visualization
   if (argc > 1 && std::string(argv[1]) == "r")
   {
      sph.start();
      sph.wait();
      return 0;
   }

   SphConfigRows config;
   config.mSph = &sph;                      // w.Config()->setSph(&sph)
   config.readValuesFromSimulation();
   WidgetSlots w;
   // a.connect(&sph, SIGNAL(updateElapsed(int x6)), &w, SLOT(updateElapsedSph(int x6)), Qt::QueuedConnection):
   // without Qt the signal is a std::function hook with the same six-int signature
   sph.updateElapsed = [&w](int a, int b, int c, int d, int e, int f) { w.updateElapsedSph(a, b, c, d, e, f); };
   sph.start();
   // a.connect(&w, SIGNAL(startClicked()), &sph, SLOT(pauseResume()));
   sph.pauseResume();
   // a.connect(&w, SIGNAL(shutDownClicked()), &sph, SLOT(stopSimulation()));
   sph.stopSimulation();
   bool done = sph.isStopped() && !sph.isPaused();
   float r2 = sph.getInteractionRadius2();
   (void)r2;
   drawParticles(&sph);
   drawVoxels(&sph);
   paintGL(&sph);
   config.writeValuesToSimulation();
   return done ? 0 : 1;
}
