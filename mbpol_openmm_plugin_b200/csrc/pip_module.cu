// Polynomial stage: the generated sm_100a kernels pip2b_eval / pip3b_eval.
//
// The two PIP evaluators (replacing poly_2b_v6x_eval, platforms/reference/src/poly-2b-v6x.cpp:3-13841, and
// poly_3b_v2x::eval, poly-3b-v2x.cpp:3-28378) are emitted as PTX by tools/gen_interp.py (2B) and tools/gen_msym.py (3B), assembled to cubins by ptxas
// at build time and embedded here as byte arrays (gen/pip_cubins.inc).  They are loaded with the driver API, obtained
// through cudaGetDriverEntryPoint so that the library has no link-time dependency on libcuda (it must be loadable, for
// symbol checks, on a machine without a GPU driver).
#include <cuda.h>

#include <mutex>

#include "internal.h"

namespace mbpol {

thread_local int64_t g_launches = 0;

namespace {

#include "gen/pip_cubins.inc"   // pip{2b,3b}_cubin[], PIP{2B,3B}_NINSTR / _SMEM (dynamic shared bytes) / _BLOCK

struct DriverApi {
    CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**) = nullptr;
    CUresult (*getErrorString)(CUresult, const char**) = nullptr;
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    bool ok = false;
};

constexpr int MAX_KERN = 4;     // variable-class kernels per polynomial (partial values / gradients accumulate)
struct DeviceModule {
    CUmodule mod2 = nullptr, mod3 = nullptr;
    CUfunction fn2[MAX_KERN] = {}, fn3[MAX_KERN] = {};
    bool ready = false;
};

std::mutex g_mu;
DriverApi g_api;
DeviceModule g_dev[64];

template <class F>
bool entry(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || p == nullptr) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

std::string cu_msg(CUresult r) {
    const char* s = nullptr;
    if (g_api.getErrorString && g_api.getErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string((int)r);
}

int ensure_loaded(int dev, std::string* err) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_api.ok) {
        if (!entry("cuModuleLoadData", g_api.moduleLoadData) || !entry("cuModuleGetFunction", g_api.moduleGetFunction) ||
            !entry("cuLaunchKernel", g_api.launchKernel) || !entry("cuGetErrorString", g_api.getErrorString) ||
            !entry("cuFuncSetAttribute", g_api.funcSetAttribute)) {
            if (err) *err = "CUDA driver entry points unavailable (no GPU driver?)";
            return MBPOL_ECUDA;
        }
        g_api.ok = true;
    }
    DeviceModule& m = g_dev[dev];
    if (m.ready) return MBPOL_OK;
    static_assert(PIP2B_NKERN <= MAX_KERN && PIP3B_NKERN <= MAX_KERN, "too many class kernels");
    cudaFree(0);   // make the primary context of the current device current for the driver calls below
    CUresult r;
    if ((r = g_api.moduleLoadData(&m.mod2, pip2b_cubin)) == CUDA_SUCCESS) r = g_api.moduleLoadData(&m.mod3, pip3b_cubin);
    for (int k = 0; r == CUDA_SUCCESS && k < PIP2B_NKERN; ++k)
        if ((r = g_api.moduleGetFunction(&m.fn2[k], m.mod2, PIP2B_ENTRIES[k])) == CUDA_SUCCESS)
            r = g_api.funcSetAttribute(m.fn2[k], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, PIP2B_SMEM);
    for (int k = 0; r == CUDA_SUCCESS && k < PIP3B_NKERN; ++k)
        if ((r = g_api.moduleGetFunction(&m.fn3[k], m.mod3, PIP3B_ENTRIES[k])) == CUDA_SUCCESS)
            r = g_api.funcSetAttribute(m.fn3[k], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, PIP3B_SMEM);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "loading the sm_100a PIP cubins failed: " + cu_msg(r);
        m = DeviceModule();
        return MBPOL_ECUDA;
    }
    m.ready = true;
    return MBPOL_OK;
}

}  // namespace

int pip_instruction_count(int order) { return order == 2 ? PIP2B_NINSTR : order == 3 ? PIP3B_NINSTR : -1; }

int pip_launch(int order, const double* x, double* out, int64_t n, int64_t ld, cudaStream_t s, std::string* err,
               bool scale_by_x) {
    if (order != 2 && order != 3) {
        if (err) *err = "pip order must be 2 or 3";
        return MBPOL_EINVAL;
    }
    if (n <= 0) return MBPOL_OK;
    if (n > 0xFFFFFFFFLL || ld > 0xFFFFFFFFLL || ld < n) {
        if (err) *err = "pip_launch: n/ld out of range";
        return MBPOL_EINVAL;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        if (err) *err = "cudaGetDevice failed";
        return MBPOL_ECUDA;
    }
    int rc = ensure_loaded(dev, err);
    if (rc != MBPOL_OK) return rc;
    unsigned un = (unsigned)n, uld = (unsigned)ld, acc = 0;
    void* args[] = {(void*)&x, (void*)&out, (void*)&un, (void*)&uld, (void*)&acc};
    const unsigned block = order == 2 ? PIP2B_BLOCK : PIP3B_BLOCK, smem = order == 2 ? PIP2B_SMEM : PIP3B_SMEM;
    const unsigned grid = (unsigned)((n + block - 1) / block);
    const int nk = order == 2 ? PIP2B_NKERN : PIP3B_NKERN;
    for (int k = 0; k < nk; ++k) {      // mode bit 0: accumulate into `out` (unused: one kernel per polynomial),
        acc = (k > 0 ? 1u : 0u) | (scale_by_x ? 2u : 0u);      // bit 1: rows 1.. = x_i dE/dx_i
        CUresult r = g_api.launchKernel(order == 2 ? g_dev[dev].fn2[k] : g_dev[dev].fn3[k], grid, 1, 1, block, 1, 1, smem,
                                        (CUstream)s, args, nullptr);
        if (r != CUDA_SUCCESS) {
            if (err) *err = "launch of the PIP kernel failed: " + cu_msg(r);
            return MBPOL_ECUDA;
        }
        MBPOL_COUNT_LAUNCH(1);
    }
    return MBPOL_OK;
}

}  // namespace mbpol
