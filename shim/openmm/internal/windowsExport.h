#ifndef OPENMM_WINDOWSEXPORT_H_
#define OPENMM_WINDOWSEXPORT_H_
/* Stand-in for OpenMM's export macro (OpenMM is not installed in this image). */
#define OPENMM_EXPORT __attribute__((visibility("default")))
#endif
