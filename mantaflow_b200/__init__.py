"""mantaflow_b200 -- B200-native (sm_100a) pressure projection of mantaflow behind the reference's plugin API.

    from mantaflow_b200 import *
    s = Solver(gridSize=(64, 64, 64), dim=3)
    flags = s.create(FlagGrid); vel = s.create(MACGrid); pressure = s.create(RealGrid)
    flags.initDomain(); flags.fillGrid()
    solvePressure(flags=flags, vel=vel, pressure=pressure, cgAccuracy=1e-4, preconditioner=PcMIC)

Importing the package does not need a GPU; creating a Solver does (there is no CPU fallback)."""
from ._lib import (MantaError, PcMGDynamic, PcMGStatic, PcMIC, PcNone, PressureParams, SolveInfo, declared_symbols,
                   device_count, load)
from .grid import (FlagEmpty, FlagFluid, FlagGrid, FlagInflow, FlagObstacle, FlagOpen, FlagOutflow, FlagStick,
                   LevelsetGrid, MACGrid, RealGrid, Solver, VecGrid)
from .pressure import (computePressureRhs, correctVelocity, lastSolveInfo, releaseMG, solvePressure, solvePressureHost,
                       solvePressureSystem)
from .cg import GridCg, GridMg, cgSolveDiffusion, cgSolveWE, vicPoisson
from .step import (PD_fluid_guiding, addBuoyancy, addGravity, addGravityNoScale, advectSemiLagrange, extrapolateLsSimple, extrapolateMACFromWeight, extrapolateMACSimple,
                   extrapolateVec3Simple, getCurvature, getLaplacian, lastGuidingIterations, releaseBlurPrecomp, setObstacleFlags, setWallBcs, updateFractions)
from .particles import (PDELETE, PNEW, IntEuler, IntRK2, IntRK4, BasicParticleSystem, IntGrid, ParticleIndexSystem, PdataInt, PdataReal, PdataVec3, flipVelocityUpdate, gridParticleIndex, mapMACToParts,
                        mapPartsToMAC, markFluidCells, pushOutofObs, unionParticleLevelset, addForcePvel, updateVelocityFromDeltaPos, eulerStep, setPartType,
                        markIsolatedFluidCell)

__all__ = [n for n in dir() if not n.startswith("_")]
