#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <string>
#include "openmm/OpenMMException.h"
namespace OpenMM {
// the subset of OpenMM::CudaArray the adapter uses: a typed device buffer with getDevicePointer() / upload / download
class CudaArray {
public:
    CudaArray() : ptr(0), size(0), elementSize(0) {}
    CudaArray(int size, int elementSize, const std::string& name) : ptr(0), size(0), elementSize(0) { initialize(size, elementSize, name); }
    ~CudaArray() { if (ptr) cudaFree(ptr); }
    void initialize(int n, int elemSize, const std::string& name) {
        size = n; elementSize = elemSize;
        if (cudaMalloc(&ptr, (size_t)n * elemSize) != cudaSuccess) throw OpenMMException("Error creating array " + name);
        cudaMemset(ptr, 0, (size_t)n * elemSize);
    }
    void* getDevicePointer() { return ptr; }           // (real OpenMM returns CUdeviceptr&: same bits)
    int getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    void upload(const void* data) { cudaMemcpy(ptr, data, (size_t)size * elementSize, cudaMemcpyHostToDevice); }
    void download(void* data) const { cudaMemcpy(data, ptr, (size_t)size * elementSize, cudaMemcpyDeviceToHost); }
private:
    CudaArray(const CudaArray&);
    void* ptr;
    int size, elementSize;
};
}  // namespace OpenMM
