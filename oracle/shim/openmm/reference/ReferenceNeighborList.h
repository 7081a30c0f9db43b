/* TEST INFRASTRUCTURE ONLY (oracle build shim). Only the typedefs the plugin's sources use. */
#ifndef ORACLE_SHIM_REFNL_H
#define ORACLE_SHIM_REFNL_H
#include <utility>
#include <vector>
#include "openmm/reference/RealVec.h"
namespace OpenMM {
typedef unsigned int AtomIndex;
typedef std::pair<AtomIndex, AtomIndex> AtomPair;
typedef std::vector<AtomPair> NeighborList;
typedef std::vector<RealVec> AtomLocationList;
}
#endif
