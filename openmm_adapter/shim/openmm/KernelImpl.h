#pragma once
#include <string>
namespace OpenMM {
class Platform;
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    std::string name;
    const Platform* platform;
};
}  // namespace OpenMM
