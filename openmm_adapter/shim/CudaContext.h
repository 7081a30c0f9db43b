#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "CudaArray.h"
#include "CudaForceInfo.h"
namespace OpenMM {
// The slice of OpenMM::CudaContext the adapter touches, with OpenMM's buffer layouts:
//   posq          real4 (float4; double4 when getUseDoublePrecision()) per padded atom, REORDERED atom order
//   atomIndex     int per padded atom: slot -> original particle index
//   force         long long[3 * paddedNumAtoms], fixed point (value * 2^32), SoA x | y | z, slot order
//   energyBuffer  double (float in single precision) partial sums; kernels add into their own entry
class CudaContext {
public:
    CudaContext(int numAtoms, bool useDoublePrecision, int deviceIndex = 0)
        : numAtoms(numAtoms), paddedNumAtoms((numAtoms + 31) / 32 * 32), useDouble(useDoublePrecision), deviceIndex(deviceIndex),
          computeForceCount(0), stream(0) {
        cudaSetDevice(deviceIndex);
        posq.initialize(paddedNumAtoms, useDouble ? sizeof(double4) : sizeof(float4), "posq");
        atomIndexDevice.initialize(paddedNumAtoms, sizeof(int), "atomIndex");
        force.initialize(3 * paddedNumAtoms, sizeof(long long), "force");
        energyBuffer.initialize(1024, useDouble ? sizeof(double) : sizeof(float), "energyBuffer");
        atomIndex.resize(paddedNumAtoms);
        for (int i = 0; i < paddedNumAtoms; ++i) atomIndex[i] = i;
        atomIndexDevice.upload(atomIndex.data());
        box = make_double4(0, 0, 0, 0);
    }
    ~CudaContext() { for (size_t i = 0; i < forces.size(); ++i) delete forces[i]; }
    void setAsCurrent() { cudaSetDevice(deviceIndex); }
    int getDeviceIndex() const { return deviceIndex; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return false; }
    CudaArray& getPosq() { return posq; }
    CudaArray& getForce() { return force; }
    CudaArray& getEnergyBuffer() { return energyBuffer; }
    CudaArray& getAtomIndexArray() { return atomIndexDevice; }
    const std::vector<int>& getAtomIndex() const { return atomIndex; }
    cudaStream_t getCurrentStream() { return stream; }
    double4 getPeriodicBoxSize() const { return box; }
    void setPeriodicBoxSize(double x, double y, double z) { box = make_double4(x, y, z, 0); }
    void addForce(CudaForceInfo* info) { forces.push_back(info); }
    std::vector<CudaForceInfo*>& getForceInfos() { return forces; }
    int getComputeForceCount() const { return computeForceCount; }
    // test-side helpers (in OpenMM: CudaContext::reorderAtoms / CudaCalcForcesAndEnergyKernel::beginComputation)
    void setAtomOrder(const std::vector<int>& order) {
        for (int i = 0; i < numAtoms; ++i) atomIndex[i] = order[i];
        atomIndexDevice.upload(atomIndex.data());
    }
    void beginComputation() { ++computeForceCount; }
private:
    int numAtoms, paddedNumAtoms;
    bool useDouble;
    int deviceIndex, computeForceCount;
    cudaStream_t stream;
    CudaArray posq, atomIndexDevice, force, energyBuffer;
    std::vector<int> atomIndex;
    std::vector<CudaForceInfo*> forces;
    double4 box;
};
}  // namespace OpenMM
