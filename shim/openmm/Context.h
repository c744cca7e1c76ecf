#ifndef OPENMM_CONTEXT_H_
#define OPENMM_CONTEXT_H_
/* The plugin's API header includes openmm/Context.h only for its transitive includes. */
#include "Vec3.h"
#include "System.h"
#include "Platform.h"
#include <map>
#include <string>
#include <vector>
#endif
