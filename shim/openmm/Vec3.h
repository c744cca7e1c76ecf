#ifndef OPENMM_VEC3_H_
#define OPENMM_VEC3_H_
/* Stand-in for OpenMM::Vec3 -- same public surface the plugin sources use
 * (operator[], +, -, unary -, scalar *, scalar /, +=, -=, dot, cross). */
#include "internal/windowsExport.h"
#include <cassert>
#include <cmath>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    bool operator==(const Vec3& r) const { return data[0]==r[0] && data[1]==r[1] && data[2]==r[2]; }
    bool operator!=(const Vec3& r) const { return !(*this == r); }
    Vec3 operator+() const { return *this; }
    Vec3 operator+(const Vec3& r) const { return Vec3(data[0]+r[0], data[1]+r[1], data[2]+r[2]); }
    Vec3& operator+=(const Vec3& r) { data[0]+=r[0]; data[1]+=r[1]; data[2]+=r[2]; return *this; }
    Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }
    Vec3 operator-(const Vec3& r) const { return Vec3(data[0]-r[0], data[1]-r[1], data[2]-r[2]); }
    Vec3& operator-=(const Vec3& r) { data[0]-=r[0]; data[1]-=r[1]; data[2]-=r[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(data[0]*s, data[1]*s, data[2]*s); }
    Vec3& operator*=(double s) { data[0]*=s; data[1]*=s; data[2]*=s; return *this; }
    Vec3 operator/(double s) const { double inv = 1.0/s; return Vec3(data[0]*inv, data[1]*inv, data[2]*inv); }
    Vec3& operator/=(double s) { double inv = 1.0/s; data[0]*=inv; data[1]*=inv; data[2]*=inv; return *this; }
    double dot(const Vec3& r) const { return data[0]*r[0] + data[1]*r[1] + data[2]*r[2]; }
    Vec3 cross(const Vec3& r) const {
        return Vec3(data[1]*r[2]-data[2]*r[1], data[2]*r[0]-data[0]*r[2], data[0]*r[1]-data[1]*r[0]);
    }
private:
    double data[3];
};
static inline Vec3 operator*(double s, const Vec3& v) { return Vec3(v[0]*s, v[1]*s, v[2]*s); }
} // namespace OpenMM
#endif
