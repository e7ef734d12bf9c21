#!/bin/bash
# Builds the reference's own `manta` executable (Python enabled, OpenMP, no GUI, float) from the sources where they lie under $REF, with ONE
# change: in the PREPROCESSED copy of source/plugin/pressure.cpp the bodies of solvePressure() and releaseMG() call libmantapress through
# the C-ABI (applied by tools/bind_pressure_plugin.py).  Everything else -- the Python wrapper, the registry, every
# other plugin, Grid / FlagGrid / MACGrid -- is the unmodified reference, so scenes/*.py run as they are and solvePressure(...) runs on
# the GPU.  The reference's cmake build system is not run; this follows CMakeLists.txt:131-700 by hand (prep generate / link / register,
# then g++).  All outputs go to oracle/_ref/bound/ (git-ignored build output that travels to the GPU box; no reference source enters the
# repository).  usage: tools/build_bound_manta.sh            then, on a GPU box:  oracle/_ref/bound/manta oracle/_ref/bound/scenes/simpleplume.py
set -e
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/oracle/_ref/bound
SRC=$REF/source
JOBS=${JOBS:-8}
if [ ! -f $SRC/plugin/pressure.cpp ]; then echo "reference tree $REF not present: keeping the prebuilt $OUT (if any)"; exit 0; fi
mkdir -p $OUT/pp/source/plugin $OUT/pp/source/util $OUT/pp/source/fileio $OUT/pp/source/python $OUT/obj $OUT/scenes
g++ -O2 -w -o $OUT/prep $SRC/preprocessor/*.cpp

PP_SOURCES="general.cpp fluidsolver.cpp conjugategrad.cpp multigrid.cpp grid.cpp grid4d.cpp levelset.cpp fastmarch.cpp shapes.cpp mesh.cpp particle.cpp movingobs.cpp
 fileio/ioutil.cpp fileio/iogrids.cpp fileio/iomeshes.cpp fileio/ioparticles.cpp fileio/iovdb.cpp fileio/mantaio.cpp noisefield.cpp kernel.cpp vortexsheet.cpp vortexpart.cpp
 turbulencepart.cpp timing.cpp edgecollapse.cpp plugin/advection.cpp plugin/extforces.cpp plugin/apic.cpp plugin/flip.cpp plugin/fire.cpp plugin/fluidguiding.cpp plugin/kepsilon.cpp
 plugin/implicitdensityprojection.cpp plugin/initplugins.cpp plugin/meshplugins.cpp plugin/pressure.cpp plugin/ptsplugins.cpp plugin/secondaryparticles.cpp plugin/surfaceturbulence.cpp
 plugin/vortexplugins.cpp plugin/waveletturbulence.cpp plugin/waves.cpp python/defines.py test.cpp plugin/numpyconvert.cpp plugin/tfplugins.cpp"
PP_HEADERS="general.h commonkernels.h conjugategrad.h multigrid.h fastmarch.h fluidsolver.h grid.h grid4d.h mesh.h particle.h levelset.h shapes.h noisefield.h vortexsheet.h kernel.h
 timing.h movingobs.h fileio/mantaio.h edgecollapse.h vortexpart.h turbulencepart.h"
NOPP_SOURCES="pwrapper/pymain.cpp pwrapper/pclass.cpp pwrapper/pvec3.cpp pwrapper/pconvert.cpp pwrapper/registry.cpp pwrapper/numpyWrap.cpp util/vectorbase.cpp util/vector4d.cpp util/simpleimage.cpp"

cd $OUT
REGS=""; GEN=""
for f in $PP_SOURCES $PP_HEADERS; do
	./prep generate 0 OPENMP $SRC/ $f $OUT/pp/source/$f > /dev/null
	case $f in *.h|*.py) REGS="$REGS $OUT/pp/source/$f.reg";; esac
	case $f in *.cpp) GEN="$GEN $OUT/pp/source/$f";; esac
done
echo "// git info not determined (bound build)" > $OUT/pp/source/gitinfo.h
./prep link $REGS > /dev/null
REGCPP=""; for r in $REGS; do REGCPP="$REGCPP $r.cpp"; done
# the one change: solvePressure / releaseMG call the C-ABI
python3 $ROOT/tools/bind_pressure_plugin.py $OUT/pp/source/plugin/pressure.cpp
ALL="$GEN $REGCPP"; for f in $NOPP_SOURCES; do ALL="$ALL $SRC/$f"; done
./prep register $ALL $REF/dependencies/cnpy/cnpy.cpp $OUT/pp/source/registration.cpp > /dev/null
ALL="$ALL $REF/dependencies/cnpy/cnpy.cpp $OUT/pp/source/registration.cpp"

PYINC=$(python3 -c "import sysconfig; print(sysconfig.get_config_var('INCLUDEPY'))")
PYLIB=$(python3 -c "import sysconfig; print(sysconfig.get_config_var('LIBDIR') + '/' + sysconfig.get_config_var('LDLIBRARY'))")
NPINC=$(python3 -c "import numpy; print(numpy.get_include())")
FLAGS="-O3 -DNDEBUG -std=c++14 -fopenmp -pthread -w -DOPENMP=1 -DMANTA_MT=1 -DNUMPY=1 -DCUDA_PRESSURE=1 -DMANTAVERSION=\"bound\"
 -I$OUT/pp/source -I$OUT/pp/source/util -I$OUT/pp/source/fileio -I$SRC/pwrapper -I$SRC/util -I$SRC/fileio -I$SRC -I$REF/dependencies/cnpy -I$PYINC -I$NPINC -I$ROOT/include"
i=0; OBJS=""
for f in $ALL; do
	o=$OUT/obj/$(echo ${f#$OUT/} | tr '/.' '__').o; OBJS="$OBJS $o"
	if [ ! -f $o ] || [ $f -nt $o ]; then
		g++ $FLAGS -c $f -o $o &
		i=$((i+1)); if [ $((i % JOBS)) -eq 0 ]; then wait; fi
	fi
done
wait
g++ -fopenmp -o $OUT/manta $OBJS $PYLIB -lz -L$ROOT/mantaflow_b200 -lmantapress -Wl,-rpath,'$ORIGIN/../../../mantaflow_b200' -ldl -lutil
# the scenes the bound binary is run on (build output, not committed)
cp $REF/scenes/simpleplume.py $OUT/scenes/
ls -la $OUT/manta
