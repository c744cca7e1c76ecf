#ifndef OPENMM_ASSERTIONUTILITIES_H_
#define OPENMM_ASSERTIONUTILITIES_H_
#include "openmm/OpenMMException.h"
#include <sstream>
#define ASSERT(cond) { if (!(cond)) throw OpenMM::OpenMMException(std::string("Assertion failure: ") + #cond); }
#define ASSERT_VALID_INDEX(index, vector) { if ((index) < 0 || (index) >= (int) (vector).size()) throw OpenMM::OpenMMException("Index out of range"); }
#endif
