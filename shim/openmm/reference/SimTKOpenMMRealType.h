#ifndef SIMTK_OPENMM_REAL_TYPE_H_
#define SIMTK_OPENMM_REAL_TYPE_H_
/* Coulomb constant in kJ mol^-1 nm e^-2 as defined by the OpenMM 7.3-7.5 line the plugin targets
 * (simtk.openmm python import, python/openmmcoul.i:4). OpenMM >= 7.6 uses 138.93545764...; this
 * repository fixes the older value everywhere (oracle, reference build, CUDA kernels). */
#include <cmath>
#define ONE_4PI_EPS0 138.935456
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#endif
