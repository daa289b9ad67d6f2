import sys
sys.path.insert(0, '.')
import smoothed_particle_hydrodynamics_b200 as S
sph = S.SPH(); sph.step_n(20); sph.synchronize()
