// Drives openmm_adapter/B200MBPolKernels.{h,cpp} the way OpenMM does -- plugin registration, KernelFactory, initialize,
// ContextImpl::calcForcesAndEnergy -> execute of both kernels -- through the stand-in CudaContext of openmm_adapter/shim,
// which owns real device buffers in OpenMM's layouts.  Mirrors platforms/cuda/tests/TestCudaMBPolThreeBodyForce.cpp:66-150
// (4 particles per molecule, M site must get zero force) on the water256_bulk system, plus what the reference test
// cannot see: a SHUFFLED atom order (cu.getAtomIndex()), real4 / double4 posq, the padded int64 force buffer pre-filled
// by "other forces", and the count of tuple-list builds per evaluation.
//
//   TestB200OpenMMAdapter <in.bin> <out.bin> <double|single>
//   in.bin : int32 n_mol, int32 periodic, double box[3], double cut2, double cut3, double pos[n_mol][3][3]   (nm)
//   out.bin: double energy, double forces[4 n_mol][3] in ORIGINAL particle order, int32 list builds, int32 evaluations
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "B200MBPolKernels.h"
#include "CudaPlatform.h"
#include "openmm/Platform.h"

using namespace OpenMM;
using namespace MBPolPlugin;

static void fail(const char* what) {
    std::printf("FAIL: %s\n", what);
    std::exit(1);
}

int main(int argc, char** argv) {
    if (argc < 4) fail("usage");
    const bool useDouble = std::strcmp(argv[3], "double") == 0;
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) fail("cannot open input");
    int n_mol = 0, periodic = 0;
    double box[3], cut2, cut3;
    if (std::fread(&n_mol, 4, 1, f) != 1 || std::fread(&periodic, 4, 1, f) != 1 || std::fread(box, 8, 3, f) != 3 ||
        std::fread(&cut2, 8, 1, f) != 1 || std::fread(&cut3, 8, 1, f) != 1)
        fail("short input");
    std::vector<double> pos(9 * (size_t)n_mol);
    if (std::fread(pos.data(), 8, pos.size(), f) != pos.size()) fail("short input");
    std::fclose(f);
    try {
        // ---- System + forces: O, H1, H2, M per molecule (python/mbpol.xml:9-14); forces name ORIGINAL particle indices
        const int numAtoms = 4 * n_mol;
        System system;
        MBPolTwoBodyForce two;
        MBPolThreeBodyForce three;
        two.setCutoff(cut2);
        three.setCutoff(cut3);
        two.setNonbondedMethod(periodic ? MBPolTwoBodyForce::CutoffPeriodic : MBPolTwoBodyForce::CutoffNonPeriodic);
        three.setNonbondedMethod(periodic ? MBPolThreeBodyForce::CutoffPeriodic : MBPolThreeBodyForce::CutoffNonPeriodic);
        three.setForceGroup(1);
        for (int m = 0; m < n_mol; ++m) {
            for (int k = 0; k < 4; ++k) system.addParticle(k == 0 ? 15.99943 : (k < 3 ? 1.007947 : 0.0));
            std::vector<int> p(3);
            for (int k = 0; k < 3; ++k) p[k] = 4 * m + k;
            two.addParticle(p);
            three.addParticle(p);
        }
        // ---- platform data with one CudaContext; OpenMM has REORDERED the atoms (molecules stay whole in OpenMM; a
        // fully shuffled order is the stronger test of the index handling)
        CudaContext cu(numAtoms, useDouble);
        std::vector<int> order(numAtoms);
        for (int i = 0; i < numAtoms; ++i) order[i] = i;
        unsigned long long seed = 88172645463325252ULL;
        for (int i = numAtoms - 1; i > 0; --i) {
            seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
            std::swap(order[i], order[(int)(seed % (unsigned long long)(i + 1))]);
        }
        cu.setAtomOrder(order);
        if (periodic) cu.setPeriodicBoxSize(box[0], box[1], box[2]);
        const int padded = cu.getPaddedNumAtoms();
        // posq in slot order; M site = average3 of mbpol.xml:14; charge in .w
        std::vector<double> posq_d(4 * (size_t)padded, 0.0);
        std::vector<float> posq_f(4 * (size_t)padded, 0.0f);
        for (int slot = 0; slot < numAtoms; ++slot) {
            const int atom = order[slot], m = atom / 4, k = atom % 4;
            double xyz[3];
            for (int c = 0; c < 3; ++c)
                xyz[c] = k < 3 ? pos[9 * m + 3 * k + c]
                               : 0.573293118 * pos[9 * m + c] + 0.213353441 * (pos[9 * m + 3 + c] + pos[9 * m + 6 + c]);
            for (int c = 0; c < 3; ++c) {
                posq_d[4 * slot + c] = xyz[c];
                posq_f[4 * slot + c] = (float)xyz[c];
            }
        }
        if (useDouble) cu.getPosq().upload(posq_d.data()); else cu.getPosq().upload(posq_f.data());
        // "other forces" already sit in the buffers: the kernels must ADD
        std::vector<long long> pre(3 * (size_t)padded);
        for (size_t i = 0; i < pre.size(); ++i) pre[i] = (long long)(i % 1000) * 4294967296LL;
        cu.getForce().upload(pre.data());
        std::vector<double> e_pre_d(1024, 0.0);
        std::vector<float> e_pre_f(1024, 0.0f);
        e_pre_d[0] = 7.0; e_pre_f[0] = 7.0f;
        if (useDouble) cu.getEnergyBuffer().upload(e_pre_d.data()); else cu.getEnergyBuffer().upload(e_pre_f.data());

        CudaPlatform::PlatformData data;
        data.contexts.push_back(&cu);
        ContextImpl context(system, &data);
        // ---- plugin registration + kernel creation exactly as OpenMM's plugin loader / ForceImpl::initialize do
        registerKernelFactories();
        Platform& platform = Platform::getPlatformByName("CUDA");
        CalcMBPolTwoBodyForceKernel* k2 = dynamic_cast<CalcMBPolTwoBodyForceKernel*>(platform.createKernel(CalcMBPolTwoBodyForceKernel::Name(), context));
        CalcMBPolThreeBodyForceKernel* k3 = dynamic_cast<CalcMBPolThreeBodyForceKernel*>(platform.createKernel(CalcMBPolThreeBodyForceKernel::Name(), context));
        if (!k2 || !k3) fail("kernel creation");
        k2->initialize(system, two);
        k3->initialize(system, three);
        if (cu.getForceInfos().size() != 2 || cu.getForceInfos()[0]->getNumParticleGroups() != n_mol) fail("CudaForceInfo");
        std::vector<int> grp;
        cu.getForceInfos()[1]->getParticlesInGroup(3, grp);
        if (grp.size() != 3 || grp[0] != 12 || grp[2] != 14) fail("CudaForceInfo groups");
        try {
            platform.createKernel("CalcSomethingElse", context);
            fail("illegal kernel name accepted");
        } catch (const OpenMMException&) {
        }
        // ---- ContextImpl::calcForcesAndEnergy(includeForces, includeEnergy, groups = all): both kernels execute
        cu.beginComputation();
        context.setLastForceGroups(0xFFFFFFFF);
        double ret = k2->execute(context, true, true) + k3->execute(context, true, true);
        if (ret != 0.0) fail("execute must return 0.0 on the CUDA platform");
        B200SharedEngine& eng = B200SharedEngine::get(cu);
        const int evaluations1 = eng.evaluations();
        const long long builds1 = eng.lastStats().n_row_builds;
        // ---- a second evaluation with only the 3B group active adds E3 / F3 once more; undone below
        std::vector<long long> after1(pre.size()), after2(pre.size());
        cudaDeviceSynchronize();
        cu.getForce().download(after1.data());
        cu.beginComputation();
        context.setLastForceGroups(1 << 1);
        k3->execute(context, true, true);
        cudaDeviceSynchronize();
        cu.getForce().download(after2.data());
        const int evaluations2 = eng.evaluations();
        double e_total;
        if (useDouble) { std::vector<double> e(1024); cu.getEnergyBuffer().download(e.data()); e_total = e[0] - 7.0; }
        else { std::vector<float> e(1024); cu.getEnergyBuffer().download(e.data()); e_total = (double)e[0] - 7.0; }
        // forces of evaluation 1 (2B + 3B) back in ORIGINAL particle order; padding slots must be untouched
        std::vector<double> force(3 * (size_t)numAtoms);
        for (int slot = 0; slot < padded; ++slot)
            for (int c = 0; c < 3; ++c) {
                const size_t i = (size_t)c * padded + slot;
                if (slot >= numAtoms) {
                    if (after2[i] != pre[i]) fail("padding entries of the force buffer were modified");
                    continue;
                }
                force[3 * (size_t)order[slot] + c] = (double)(after1[i] - pre[i]) / 4294967296.0;
            }
        // evaluation 2 must have added exactly the 3B part: F(2) - F(1) = F3 -> report its energy share via e_total
        std::FILE* o = std::fopen(argv[2], "wb");
        if (!o) fail("cannot open output");
        std::fwrite(&e_total, 8, 1, o);
        std::fwrite(force.data(), 8, force.size(), o);
        std::vector<double> f3(3 * (size_t)numAtoms);
        for (int slot = 0; slot < numAtoms; ++slot)
            for (int c = 0; c < 3; ++c)
                f3[3 * (size_t)order[slot] + c] = (double)(after2[(size_t)c * padded + slot] - after1[(size_t)c * padded + slot]) / 4294967296.0;
        std::fwrite(f3.data(), 8, f3.size(), o);
        const int meta[3] = {(int)builds1, evaluations1, evaluations2};
        std::fwrite(meta, 4, 3, o);
        std::fclose(o);
        // ---- copyParametersToContext: wrong molecule count is the reference's error (MBPolReferenceKernels.cpp:402-403)
        MBPolTwoBodyForce fewer;
        fewer.addParticle(std::vector<int>(3, 0));
        try {
            k2->copyParametersToContext(context, fewer);
            fail("particle-count change accepted");
        } catch (const OpenMMException& e) {
            if (std::string(e.what()).find("number of particles has changed") == std::string::npos) fail("wrong message");
        }
        delete k2;
        delete k3;                                  // the shared engine goes with its last kernel
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 1;
    }
    std::printf("Done\n");
    return 0;
}
