// vec3.h -- float3 / int3 value types of the SPH facade's API (the reference
// declares the same public members in src/vec3.h:6-125 and src/vec3i.h:4-7; only
// getGravity / setGravity and a few internals use them).
#ifndef SPHB200_HOST_VEC3_H
#define SPHB200_HOST_VEC3_H

#include <cmath>

class vec3
{
public:
   float x, y, z;

   vec3() : x(0.0f), y(0.0f), z(0.0f) {}
   vec3(float ax, float ay, float az) : x(ax), y(ay), z(az) {}

   void set(float ax, float ay, float az) { x = ax; y = ay; z = az; }

   vec3 cross(const vec3& o) const { return vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
   float operator*(const vec3& o) const { return x * o.x + y * o.y + z * o.z; }   // dot product
   vec3 operator+(const vec3& o) const { return vec3(x + o.x, y + o.y, z + o.z); }
   vec3 operator-(const vec3& o) const { return vec3(x - o.x, y - o.y, z - o.z); }
   vec3 operator*(float s) const { return vec3(x * s, y * s, z * s); }
   friend vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
   vec3 operator/(float s) const { float r = 1.0f / s; return vec3(x * r, y * r, z * r); }
   void operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; }
   void operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; }
   void operator*=(float s) { x *= s; y *= s; z *= s; }
   void operator/=(float s) { float r = 1.0f / s; x *= r; y *= r; z *= r; }
   bool operator==(const vec3& o) const { return x == o.x && y == o.y && z == o.z; }
   bool operator!=(const vec3& o) const { return !(*this == o); }
   float length2() const { return x * x + y * y + z * z; }
   float length() const { return static_cast<float>(std::sqrt(length2())); }
};

struct vec3i
{
   int x, y, z;
};

#endif
