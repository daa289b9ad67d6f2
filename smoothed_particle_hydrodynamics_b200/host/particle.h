// particle.h -- host mirror of the particle arrays with the reference's member
// names and layouts (src/particle.h:7-20): one object, structure of arrays,
// xyz interleaved with stride 3.  The GL view reads mPosition[i*3+k] directly
// (src/visualization.cpp:144-157), so these stay std::vector<float>.  The
// authoritative state lives in HBM; SPH refreshes this mirror (see sph.h).
#ifndef SPHB200_HOST_PARTICLE_H
#define SPHB200_HOST_PARTICLE_H

#include <cstddef>
#include <vector>

class Particle
{
public:
   explicit Particle(size_t numParticles)
    : mMass(numParticles, 0.0f),
      mDensity(numParticles, 0.0f),
      mPosition(numParticles * 3, 0.0f),
      mVelocity(numParticles * 3, 0.0f),
      mAcceleration(numParticles * 3, 0.0f),
      mNeighborCount(numParticles, 0)
   {
   }

   std::vector<float> mMass;
   std::vector<float> mDensity;
   std::vector<float> mPosition;
   std::vector<float> mVelocity;
   std::vector<float> mAcceleration;
   std::vector<int> mNeighborCount;
};

#endif
