#pragma once
#include <vector>
#include "CudaContext.h"
#include "openmm/Platform.h"
namespace OpenMM {
class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        std::vector<CudaContext*> contexts;
    };
};
}  // namespace OpenMM
