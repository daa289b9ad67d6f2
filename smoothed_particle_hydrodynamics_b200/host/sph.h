// sph.h -- C++ facade with the public interface of the reference's `class SPH`
// (/root/reference/src/sph.h:15-84) over the C ABI of include/sphb200.h.
//
// widget.cpp / visualization.cpp / sphconfig.cpp / main.cpp of the reference use:
// the default constructor, the getters/setters, the slots run/step/pauseResume/
// stopSimulation, the signals updateElapsed(int x6)/stepFinished(), getParticles()
// ->mPosition and getGrid()[c].count().  All of those exist here with the same
// names and meaning.  Built with -DSPHB200_WITH_QT the class derives from QThread
// and the signals are real Qt signals (the GUI sources then compile unmodified);
// without Qt (this repo's headless driver) the base is a minimal stand-in and the
// signals are std::function hooks.
#ifndef SPHB200_HOST_SPH_H
#define SPHB200_HOST_SPH_H

#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "sphb200.h"
#include "vec3.h"

#ifdef SPHB200_WITH_QT
#include <QList>
#include <QMutex>
#include <QThread>
typedef QThread SphThreadBase;
typedef QList<uint32_t> SphCellList;
#else
// what the facade needs from QThread when there is no Qt: start() runs run()
class SphThreadBase
{
public:
   virtual ~SphThreadBase() {}
   virtual void run() {}
   void start() { run(); }
   void quit() {}
   void wait() {}
};

// stand-in for QList<uint32_t> so that `QList<uint32_t>* grid = mSph->getGrid();
// grid[index].count()` (visualization.cpp:178, 193) compiles verbatim without Qt.  The
// GL view only calls count(); the members are present only after
// setGridMembership(true) (they cost a download of every particle index per call).
template <typename T> class QList;
template <> class QList<uint32_t>
{
public:
   QList() : mBegin(nullptr), mCount(0) {}
   int count() const { return mCount; }
   int length() const { return mCount; }
   int size() const { return mCount; }
   bool isEmpty() const { return mCount == 0; }
   bool hasMembers() const { return mBegin != nullptr || mCount == 0; }
   uint32_t operator[](int i) const { return mBegin[i]; }
   void assign(const uint32_t* begin, int n) { mBegin = begin; mCount = n; }
private:
   const uint32_t* mBegin;
   int mCount;
};
typedef QList<uint32_t> SphCellList;
#endif

class Particle;

class SPH : public SphThreadBase
{
#ifdef SPHB200_WITH_QT
   Q_OBJECT
#endif

public:
   // the reference constructor (sph.cpp:36-118): its literals + the seeded sphere scene
   SPH();
   // any other configuration (the reference hard-codes everything).
   // initSphereScene = true reproduces initParticlePolitionsSphere for this size.
   explicit SPH(const SphParams& params, bool initSphereScene = false, int device = -1);
   virtual ~SPH();

   // control
   bool isStopped() const;
   bool isPaused() const;

   // getters and setters (sph.h:32-61)
   Particle* getParticles();          // host mirror; picks up the newest position snapshot (see Readback)
   int getParticleCount() const;
   void getGridCellCounts(int& x, int& y, int& z);
   void getParticleBounds(float& x, float& y, float& z);
   float getInteractionRadius2() const;
   SphCellList* getGrid();            // per-voxel lists, [x + y*gx + z*gx*gy]; count() from the newest snapshot
   float getCellSize() const;
   vec3 getGravity() const;
   void setGravity(const vec3& gravity);
   float getStiffness() const;
   void setStiffness(float stiffness);
   float getViscosityScalar() const;
   void setViscosityScalar(float viscosityScalar);
   float getTimeStep() const;
   void setTimeStep(float timeStep);
   float getDamping() const;
   void setDamping(float damping);
   float getCflLimit() const;
   void setCflLimit(float cflLimit);

   // ---- extensions (no reference counterpart) -----------------------------------
   // What step() brings back from HBM.  The reference updates every array in place; its
   // only per-frame readers are the GL view's drawParticles (mPosition, visualization.cpp:
   // 137-163) and drawVoxels (getGrid()[c].count(), 166-213), on a 16 ms timer, from the
   // GUI thread.
   //   ReadbackPositions (default): step() REQUESTS an asynchronous snapshot (positions +
   //     per-voxel counts -> pinned double buffer, copied on a side stream while the next
   //     steps run; sphb200_snapshot_request) at most once per readback interval;
   //     getParticles() / getGrid() pick up the newest completed snapshot.  Interval 0 =
   //     a request every step, and the getters wait for it: the mirror is then current
   //     after every step, as in the reference.
   //   ReadbackAll: every Particle array is downloaded synchronously inside step().
   //   ReadbackNone: nothing; call refreshParticles() explicitly.
   enum Readback { ReadbackNone = 0, ReadbackPositions = 1, ReadbackAll = 2 };
   void setReadback(Readback r) { mReadback = r; }
   void setReadbackIntervalMs(int ms) { mReadbackIntervalMs = ms; }   // default 0; the GL timer is 16 (visualization.cpp:24-33)
   void setGridMembership(bool on) { mGridMembership = on; }          // getGrid() also fills the member indices (always in a Qt build)
   // the two switches the GUI rows `gravity` / `damping` were meant to drive (sphconfig.cpp:
   // 76-95; both inert in the reference).  setGravity(non-zero) and
   // setDamping(value != the constructor's) turn them on implicitly.
   void setUniformGravity(bool on);
   void setWallCollision(bool on);
   void uploadState(const float* posXyz, const float* velXyz, const float* mass);
   void refreshParticles(Readback what);           // explicit device -> host mirror copy
   void setTotalSteps(int n) { mTotalSteps = n; }  // run() performs n + 1 steps like the reference
   void setOutputDirectory(const std::string& dir) { mOutDir = dir; }
   float kineticEnergy() const { return mKineticEnergyTotal; }
   float potentialEnergy() const { return mPotentialEnergyTotal; }
   sphb200_ctx* context() { return mCtx; }

#ifdef SPHB200_WITH_QT
public slots:
#endif
   void run();             // sph.cpp:149-187: totalSteps + 1 steps and the four log files
   void step();            // sph.cpp:190-304
   void pauseResume();
   void stopSimulation();

#ifdef SPHB200_WITH_QT
signals:
   void updateElapsed(int, int, int, int, int, int);
   void stepFinished();
#else
public:
   std::function<void(int, int, int, int, int, int)> updateElapsed;   // ms per phase
   std::function<void()> stepFinished;
#endif

protected:
   void init(const SphParams& params, bool sphereScene, int device);
   void check(int rc, const char* what) const;
   void pushParams();

   void requestSnapshot();                       // stepping thread, under mCtxMutex
   bool pullSnapshot(bool positions, bool counts);   // any thread

   sphb200_ctx* mCtx;
   mutable std::mutex mCtxMutex;   // every call on mCtx except sphb200_snapshot_read (the GUI thread's setters
                                   // and getGrid() run beside the worker's step(); sphb200.h: calls must not overlap)
   SphParams mParams;
   SphDerived mDerived;
   Particle* mSrcParticles;
   Readback mReadback;
   int mReadbackIntervalMs;
   bool mGridMembership;
   float mDefaultDamping;
   long long mStepIndex;           // steps run so far (from the step report)
   long long mSnapshotHave;        // step index of the snapshot in the mirror (-1: none)
   long long mGridHave;
   long long mLastRequestNs;
   std::mutex mMirrorMutex;        // the two getters may be called from different threads

   std::vector<SphCellList> mGrid;
   std::vector<int> mGridCounts;
   std::vector<int> mGridStart;
   std::vector<uint32_t> mGridMembers;

   int mTotalSteps;
   std::string mOutDir;
   float mKineticEnergyTotal;
   float mPotentialEnergyTotal;
   vec3 mAngularMomentumTotal;
   int timeVoxelize, timeFindNeighbors, timeComputeDensity, timeComputePressure, timeComputeAcceleration,
      timeIntegrate;
   long long mNeighborTotal;
   int mNeighborMax, mNeighborMin;

   mutable std::mutex mMutex;
   bool mStopped;
   bool mPaused;
};

#endif
