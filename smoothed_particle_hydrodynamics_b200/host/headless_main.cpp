// headless_main.cpp -- the reference's headless branch (`./sph r`,
// src/main.cpp:23-28: sph.start(); sph.wait();) without QApplication: constructs
// the reference scene, runs totalSteps + 1 steps on the GPU and writes
// out/energy.txt, out/angularmomentum.txt, out/timing.txt, out/neighbors.txt in the
// reference's formats.
//
//   sph_headless [steps] [outdir]        steps: run steps+1 steps (default 1000 like the reference)
#include <cstdlib>
#include <iostream>
#include <string>

#include "particle.h"
#include "sph.h"

int main(int argc, char** argv)
{
   try
   {
      SPH sph;   // reference defaults + seeded sphere (sph.cpp:36-118)
      if (argc > 1)
         sph.setTotalSteps(std::atoi(argv[1]));
      if (argc > 2)
         sph.setOutputDirectory(argv[2]);
      sph.setReadback(SPH::ReadbackPositions);   // what the GL view would read every frame
      sph.start();
      sph.wait();
      Particle* p = sph.getParticles();
      std::cout << "particles " << sph.getParticleCount() << " first position " << p->mPosition[0] << " "
                << p->mPosition[1] << " " << p->mPosition[2] << " E_kin " << sph.kineticEnergy() << " E_pot "
                << sph.potentialEnergy() << std::endl;
   }
   catch (const std::exception& e)
   {
      std::cerr << e.what() << std::endl;
      return 1;
   }
   return 0;
}
