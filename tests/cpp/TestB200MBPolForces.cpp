// Known-answer tests of the 2B and 3B forces through the C++ host layer, written the way the reference tests its
// platforms (platforms/reference/tests/TestReferenceMBPolTwoBodyForce.cpp, TestReferenceMBPolThreeBodyForce.cpp,
// platforms/cuda/tests/TestCudaMBPolThreeBodyForce.cpp): same water geometries (Angstrom -> nm), same expected energies
// and gradients (legacy mbpol code, kcal/mol and kcal/mol/A), same three scenarios -- cluster, periodic box of 50 nm,
// periodic box with the second molecule displaced by one box length -- plus the 4-particles-per-molecule layout with an
// M site that must receive exactly zero force.  Tolerance tightened from the reference's 1e-3 to 1e-7.
#include <cmath>
#include <cstdio>
#include <iostream>

#include "mbpol_host.h"

using namespace OpenMM;
using namespace MBPolPlugin;

#define ASSERT_EQUAL_TOL(expected, found, tol) \
    do { double e_ = (expected), f_ = (found); if (!(std::fabs(e_ - f_) <= (tol) * std::fmax(1.0, std::fabs(e_)))) { \
        char b_[256]; std::snprintf(b_, sizeof b_, "line %d: expected %.10g found %.10g", __LINE__, e_, f_); throw OpenMMException(b_); } } while (0)
#define ASSERT_EQUAL_VEC(expected, found, tol) \
    do { for (int c_ = 0; c_ < 3; ++c_) ASSERT_EQUAL_TOL((expected)[c_], (found)[c_], tol); } while (0)

static const double CalToJoule = 4.184;
static const double waterA[9][3] = {
    {-1.516074336e+00, -2.023167650e-01, 1.454672917e+00}, {-6.218989773e-01, -6.009430735e-01, 1.572437625e+00},
    {-2.017613812e+00, -4.190350349e-01, 2.239642849e+00}, {-1.763651687e+00, -3.816594649e-01, -1.300353949e+00},
    {-1.903851736e+00, -4.935677617e-01, -3.457810126e-01}, {-2.527904158e+00, -7.613550077e-01, -1.733803676e+00},
    {-5.588472140e-01, 2.006699172e+00, -1.392786582e-01}, {-9.411558180e-01, 1.541226676e+00, 6.163293071e-01},
    {-9.858551734e-01, 1.567124294e+00, -8.830970941e-01}};

template <class Force>
static void addMolecules(Force& force, int numMolecules, int particlesPerMolecule) {
    for (int m = 0; m < numMolecules; ++m) {
        std::vector<int> particleIndices(3);
        for (int k = 0; k < 3; ++k) particleIndices[k] = m * particlesPerMolecule + k;
        force.addParticle(particleIndices);
    }
}

static std::vector<Vec3> makePositions(int numMolecules, int particlesPerMolecule, double boxDimension, bool addPositionOffset,
                                       double globalShift) {
    std::vector<Vec3> positions(numMolecules * particlesPerMolecule);
    for (int m = 0; m < numMolecules; ++m)
        for (int k = 0; k < 3; ++k) {
            const double* p = waterA[3 * m + k];
            positions[m * particlesPerMolecule + k] = Vec3(p[0] * 1e-1 + globalShift, p[1] * 1e-1 + globalShift, p[2] * 1e-1 + globalShift);
        }
    if (addPositionOffset)              // move second molecule 1 box dimension in Y direction
        for (int k = 0; k < 3; ++k) positions[particlesPerMolecule + k][1] += boxDimension;
    return positions;
}

void testTwoBody(double boxDimension, bool addPositionOffset) {
    MBPolTwoBodyForce* mbpolTwoBodyForce = new MBPolTwoBodyForce();
    mbpolTwoBodyForce->setCutoff(10);
    Vec3 box[3] = {Vec3(boxDimension, 0, 0), Vec3(0, boxDimension, 0), Vec3(0, 0, boxDimension)};
    mbpolTwoBodyForce->setNonbondedMethod(boxDimension > 0 ? MBPolTwoBodyForce::CutoffPeriodic : MBPolTwoBodyForce::CutoffNonPeriodic);
    addMolecules(*mbpolTwoBodyForce, 2, 3);
    const double expectedGradients[6][3] = {{-4.85337479, -4.47836379, -20.08989563}, {-0.31239868, 0.52518586, -1.88893830},
                                            {0.00886712, 0.73323536, -1.81715325},    {-0.65181727, -0.72947395, 5.88973293},
                                            {4.82340981, 3.20090213, 16.49522051},    {0.98531382, 0.74851439, 1.41103374}};
    const double expectedEnergy = 6.14207815;
    B200Context context(6, mbpolTwoBodyForce, nullptr, boxDimension > 0 ? box : nullptr);
    context.setPositions(makePositions(2, 3, boxDimension, addPositionOffset, 0.0));
    B200State state = context.getState();
    ASSERT_EQUAL_TOL(expectedEnergy, state.getPotentialEnergy() / CalToJoule, 1e-8);
    for (int i = 0; i < 6; ++i) {
        Vec3 f = state.getForces()[i];
        f /= CalToJoule * 10;
        Vec3 expected(-expectedGradients[i][0], -expectedGradients[i][1], -expectedGradients[i][2]);   // gradient -> force
        ASSERT_EQUAL_VEC(expected, f, 1e-7);
    }
    delete mbpolTwoBodyForce;
}

void testThreeBody(double boxDimension, bool addPositionOffset, int particlesPerMolecule, double globalShift) {
    MBPolThreeBodyForce* mbpolThreeBodyForce = new MBPolThreeBodyForce();
    mbpolThreeBodyForce->setCutoff(10);
    Vec3 box[3] = {Vec3(boxDimension, 0, 0), Vec3(0, boxDimension, 0), Vec3(0, 0, boxDimension)};
    mbpolThreeBodyForce->setNonbondedMethod(boxDimension > 0 ? MBPolThreeBodyForce::CutoffPeriodic : MBPolThreeBodyForce::CutoffNonPeriodic);
    addMolecules(*mbpolThreeBodyForce, 3, particlesPerMolecule);
    const double expectedGradients[9][3] = {{0.29919011, -0.34960381, -0.16238472}, {0.34138467, -0.01255068, -0.00998383},
                                            {-0.44376649, 0.03687577, 0.54604510},  {-0.01094164, -0.36171476, -0.05130395},
                                            {0.24939202, 1.29382952, 0.22930712},   {-0.13250943, -0.19313418, -0.34123592},
                                            {0.56722869, 0.46036139, -0.39999973},  {-0.75669111, -0.76132457, -0.29799486},
                                            {-0.11328682, -0.11273867, 0.48755080}};
    const double expectedEnergy = 0.15586446;
    B200Context context(3 * particlesPerMolecule, nullptr, mbpolThreeBodyForce, boxDimension > 0 ? box : nullptr);
    context.setPositions(makePositions(3, particlesPerMolecule, boxDimension, addPositionOffset, globalShift));
    B200State state = context.getState();
    ASSERT_EQUAL_TOL(expectedEnergy, state.getPotentialEnergy() / CalToJoule, 1e-8);
    for (int m = 0; m < 3; ++m)
        for (int k = 0; k < particlesPerMolecule; ++k) {
            Vec3 f = state.getForces()[m * particlesPerMolecule + k];
            f /= CalToJoule * 10;
            if (k == 3) { ASSERT_EQUAL_VEC(Vec3(0, 0, 0), f, 0.0); continue; }        // virtual site M: exactly zero
            const double* g = expectedGradients[3 * m + k];
            ASSERT_EQUAL_VEC(Vec3(-g[0], -g[1], -g[2]), f, 2e-7);
        }
    delete mbpolThreeBodyForce;
}

// platforms/reference/tests/TestReferenceMBPolOneBodyForce.cpp:52-140 (testOneOneBody): same molecule, same expected
// gradients (kcal/mol/A) and energy; the reference's tolerance is 1e-4
void testOneBody(double boxDimension) {
    MBPolOneBodyForce force;
    std::vector<int> ids = {0, 1, 2};
    force.addOneBody(ids);
    std::vector<Vec3> positions = makePositions(1, 3, boxDimension, false, 0.0);
    Vec3 boxv[3] = {Vec3(boxDimension, 0, 0), Vec3(0, boxDimension, 0), Vec3(0, 0, boxDimension)};
    if (boxDimension > 0) {
        force.setNonbondedMethod(MBPolOneBodyForce::Periodic);
        positions[1] = positions[1] + Vec3(boxDimension, 0, 0);        // a hydrogen one box edge away from its oxygen
    }
    B200Context context(3, nullptr, nullptr, boxDimension > 0 ? boxv : nullptr, &force);
    context.setPositions(positions);
    B200State state = context.getState();
    const double expectedGrad[3][3] = {{-27.48162433, 8.92495995, 2.80995323}, {30.78909844, -11.48714187, -0.27204770},
                                       {-3.30747410, 2.56218193, -2.53790553}};
    ASSERT_EQUAL_TOL(0.55975882, state.getPotentialEnergy() / CalToJoule, 1e-8);
    for (int i = 0; i < 3; ++i) {
        Vec3 f = state.getForces()[i] * (1.0 / (CalToJoule * 10));
        ASSERT_EQUAL_VEC(Vec3(-expectedGrad[i][0], -expectedGrad[i][1], -expectedGrad[i][2]), f, 1e-7);
    }
}

void testErrors() {
    MBPolTwoBodyForce force;
    force.setCutoff(0.65);
    force.setNonbondedMethod(MBPolTwoBodyForce::CutoffPeriodic);
    addMolecules(force, 2, 3);
    Vec3 small[3] = {Vec3(1.2, 0, 0), Vec3(0, 1.2, 0), Vec3(0, 0, 1.2)};
    bool threw = false;
    try { B200Context context(6, &force, nullptr, small); } catch (const OpenMMException&) { threw = true; }
    if (!threw) throw OpenMMException("cutoff > box/2 must be rejected (MBPolTwoBodyForceImpl.cpp:60-66)");
    Vec3 ok[3] = {Vec3(2, 0, 0), Vec3(0, 2, 0), Vec3(0, 0, 2)};
    B200Context context(6, &force, nullptr, ok);
    context.setPositions(makePositions(2, 3, 0, false, 0.0));
    context.setPeriodicBoxVectors(Vec3(1.25, 0, 0), Vec3(0, 1.25, 0), Vec3(0, 0, 1.25));   // box shrinks below 2 x cutoff
    threw = false;
    try { context.getState(); } catch (const OpenMMException& e) {
        threw = std::string(e.what()).find("less than twice the cutoff") != std::string::npos;   // MBPolReferenceKernels.cpp:387-390
    }
    if (!threw) throw OpenMMException("box < 2 x cutoff must throw at execute");
}

int main() {
    try {
        std::cout << "TestB200MBPolForces running test..." << std::endl;
        testTwoBody(0, false);
        testTwoBody(50, false);
        testTwoBody(50, true);
        testThreeBody(0, false, 3, 0.0);
        testThreeBody(50, false, 3, 0.0);
        testThreeBody(50, true, 3, 0.0);
        testThreeBody(50, false, 4, 25.0);      // TestCudaMBPolThreeBodyForce.cpp: M sites, positions offset by +25 nm
        testOneBody(0);
        testOneBody(50);
        testErrors();
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        std::cout << "FAIL - ERROR.  Test failed." << std::endl;
        return 1;
    }
    std::cout << "Done" << std::endl;
    return 0;
}
