/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product.
 *
 * Minimal stand-in for OpenMM's Vec3 so that the reference's hot-path sources
 * (platforms/reference/src/MBPolReference{Two,Three}BodyForce.cpp, ReferenceThreeNeighborList.cpp,
 * mbpol_2body_constants.cpp) compile unmodified, where they lie, without an OpenMM install.
 * Only the operations those files use are provided. */
#ifndef ORACLE_SHIM_VEC3_H
#define ORACLE_SHIM_VEC3_H
#include <cassert>
#include <cmath>   /* the real header pulls in <cmath>; the reference sources rely on that */
namespace OpenMM {
class Vec3 {
public:
    Vec3() { v[0] = v[1] = v[2] = 0.0; }
    Vec3(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    Vec3 operator*(double s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
    Vec3 operator/(double s) const { double r = 1.0 / s; return Vec3(v[0] * r, v[1] * r, v[2] * r); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    Vec3& operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    double dot(const Vec3& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
private:
    double v[3];
};
static inline Vec3 operator*(double s, const Vec3& a) { return a * s; }
}  // namespace OpenMM
#endif
