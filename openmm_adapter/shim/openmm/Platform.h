#pragma once
#include <map>
#include <string>
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
namespace OpenMM {
class Platform {
public:
    explicit Platform(const std::string& name = "CUDA") : name(name) {}
    const std::string& getName() const { return name; }
    void registerKernelFactory(const std::string& kernelName, KernelFactory* factory) { factories[kernelName] = factory; }
    KernelImpl* createKernel(const std::string& kernelName, ContextImpl& context) const {
        std::map<std::string, KernelFactory*>::const_iterator it = factories.find(kernelName);
        if (it == factories.end()) throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
        return it->second->createKernelImpl(kernelName, *this, context);
    }
    static Platform& getPlatformByName(const std::string& name) {
        static Platform cuda("CUDA");
        if (name != "CUDA") throw OpenMMException("There is no registered Platform called \"" + name + "\"");
        return cuda;
    }
private:
    std::string name;
    std::map<std::string, KernelFactory*> factories;
};
}  // namespace OpenMM
