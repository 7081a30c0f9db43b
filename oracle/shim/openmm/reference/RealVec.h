/* TEST INFRASTRUCTURE ONLY (oracle build shim). OpenMM <= 7.x: RealVec is a double Vec3, RealOpenMM is double. */
#ifndef ORACLE_SHIM_REALVEC_H
#define ORACLE_SHIM_REALVEC_H
#include "openmm/Vec3.h"
typedef double RealOpenMM;
namespace OpenMM { typedef Vec3 RealVec; }
#endif
