// compat_preamble.cuh -- BASELINE TOOLING, NOT PRODUCT.
//
// The plugin's existing CUDA kernels (/root/reference/platforms/cuda/src/kernels/*.cu) are source strings
// that OpenMM's CudaContext::createModule compiles at run time after prepending its own preamble (types,
// math macros, periodic-box macros, the vectorOps operators) and the #defines listed in
// platforms/cuda/src/CudaCoulKernels.cpp:377-389,466-506. OpenMM is not available here, so this header
// supplies an equivalent preamble for OpenMM's mixed-precision mode (real = float, mixed = double) and the
// kernels are compiled AOT for sm_100a, read in place from the reference tree (oracle/refcuda/Makefile).
// Nothing from the reference or from OpenMM is copied: these are the standard meanings of the names.
#pragma once
typedef float real;
typedef float2 real2;
typedef float3 real3;
typedef float4 real4;
typedef double mixed;
typedef double2 mixed2;
typedef double3 mixed3;
typedef double4 mixed4;
typedef unsigned int tileflags;
#define make_real2 make_float2
#define make_real3 make_float3
#define make_real4 make_float4
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#define SQRT sqrtf
#define RSQRT rsqrtf
#define RECIP(x) (1.0f/(x))
#define EXP expf
#define LOG logf
#define POW powf
#define COS cosf
#define SIN sinf
#define TAN tanf
#define ACOS acosf
#define ASIN asinf
#define ATAN atanf
#define ERF erff
#define ERFC erfcf
#define SHFL(var, srcLane) __shfl_sync(0xffffffff, var, srcLane)
#define BALLOT(var) __ballot_sync(0xffffffff, var)
// rectangular periodic box (the benchmark boxes are cubic)
#define APPLY_PERIODIC_TO_DELTA(delta) { \
    delta.x -= floorf(delta.x*invPeriodicBoxSize.x+0.5f)*periodicBoxSize.x; \
    delta.y -= floorf(delta.y*invPeriodicBoxSize.y+0.5f)*periodicBoxSize.y; \
    delta.z -= floorf(delta.z*invPeriodicBoxSize.z+0.5f)*periodicBoxSize.z; }
#define APPLY_PERIODIC_TO_POS(pos) { \
    pos.x -= floorf(pos.x*invPeriodicBoxSize.x)*periodicBoxSize.x; \
    pos.y -= floorf(pos.y*invPeriodicBoxSize.y)*periodicBoxSize.y; \
    pos.z -= floorf(pos.z*invPeriodicBoxSize.z)*periodicBoxSize.z; }
#define APPLY_PERIODIC_TO_POS_WITH_CENTER(pos, center) { \
    pos.x -= floorf((pos.x-center.x)*invPeriodicBoxSize.x+0.5f)*periodicBoxSize.x; \
    pos.y -= floorf((pos.y-center.y)*invPeriodicBoxSize.y+0.5f)*periodicBoxSize.y; \
    pos.z -= floorf((pos.z-center.z)*invPeriodicBoxSize.z+0.5f)*periodicBoxSize.z; }

// component-wise vector operators (what OpenMM's "vectorOps" source block provides)
#define CFX_VEC_OPS(T3, T4, S, mk3, mk4) \
__device__ inline T3 operator+(T3 a, T3 b) { return mk3(a.x+b.x, a.y+b.y, a.z+b.z); } \
__device__ inline T3 operator-(T3 a, T3 b) { return mk3(a.x-b.x, a.y-b.y, a.z-b.z); } \
__device__ inline T3 operator-(T3 a) { return mk3(-a.x, -a.y, -a.z); } \
__device__ inline T3 operator*(T3 a, S s) { return mk3(a.x*s, a.y*s, a.z*s); } \
__device__ inline T3 operator*(S s, T3 a) { return mk3(a.x*s, a.y*s, a.z*s); } \
__device__ inline T3 operator*(T3 a, T3 b) { return mk3(a.x*b.x, a.y*b.y, a.z*b.z); } \
__device__ inline T3 operator/(T3 a, S s) { S i = 1/s; return mk3(a.x*i, a.y*i, a.z*i); } \
__device__ inline void operator+=(T3& a, T3 b) { a.x += b.x; a.y += b.y; a.z += b.z; } \
__device__ inline void operator-=(T3& a, T3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; } \
__device__ inline void operator*=(T3& a, S s) { a.x *= s; a.y *= s; a.z *= s; } \
__device__ inline void operator/=(T3& a, S s) { S i = 1/s; a.x *= i; a.y *= i; a.z *= i; } \
__device__ inline S dot(T3 a, T3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; } \
__device__ inline T3 cross(T3 a, T3 b) { return mk3(a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x); } \
__device__ inline T3 trimTo3(T4 v) { return mk3(v.x, v.y, v.z); } \
__device__ inline T4 operator+(T4 a, T4 b) { return mk4(a.x+b.x, a.y+b.y, a.z+b.z, a.w+b.w); } \
__device__ inline T4 operator-(T4 a, T4 b) { return mk4(a.x-b.x, a.y-b.y, a.z-b.z, a.w-b.w); } \
__device__ inline T4 operator*(T4 a, S s) { return mk4(a.x*s, a.y*s, a.z*s, a.w*s); } \
__device__ inline T4 operator*(S s, T4 a) { return mk4(a.x*s, a.y*s, a.z*s, a.w*s); } \
__device__ inline void operator+=(T4& a, T4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; } \
__device__ inline void operator*=(T4& a, S s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }
CFX_VEC_OPS(float3, float4, float, make_float3, make_float4)
CFX_VEC_OPS(double3, double4, double, make_double3, make_double4)
__device__ inline float2 operator*(float2 a, float s) { return make_float2(a.x*s, a.y*s); }
__device__ inline float2 operator*(float s, float2 a) { return make_float2(a.x*s, a.y*s); }
__device__ inline float2 operator+(float2 a, float2 b) { return make_float2(a.x+b.x, a.y+b.y); }
__device__ inline float2 operator-(float2 a, float2 b) { return make_float2(a.x-b.x, a.y-b.y); }
__device__ inline void operator+=(float2& a, float2 b) { a.x += b.x; a.y += b.y; }
// single-argument constructors (vectorOps provides them)
__device__ inline float2 make_float2(float a) { return make_float2(a, a); }
__device__ inline float3 make_float3(float a) { return make_float3(a, a, a); }
__device__ inline float4 make_float4(float a) { return make_float4(a, a, a, a); }
__device__ inline double2 make_double2(double a) { return make_double2(a, a); }
__device__ inline double3 make_double3(double a) { return make_double3(a, a, a); }
__device__ inline double4 make_double4(double a) { return make_double4(a, a, a, a); }
