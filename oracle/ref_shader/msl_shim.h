// msl_shim.h — the part of <metal_stdlib> that the reference's src/shaders.metal uses, written for g++ so that the
// reference's OWN device source can be compiled and run on the CPU, unmodified, where it lies under /root/reference
// (see ref_shader.cpp and oracle/Makefile: output goes to oracle/_ref/ only; no reference source is copied).
//
// TEST INFRASTRUCTURE (like everything under oracle/): it pins oracle/mm_oracle.cpp — the restatement the CUDA kernel is
// checked against — to the control flow, expression order and constants of the reference's shader text itself.
//
// What is the reference's and what is this shim's: every statement of compute_shader and its helpers is executed from the
// reference file.  The shim supplies what Metal's library and hardware would: vector types and their component-wise
// operators, dot/cross/length/normalize/reflect/sign/min/max/sqrt/pow, texture read/write/sample, the float -> uint
// conversion, threadgroup memory and barriers.  Those follow the canonical arithmetic of SURVEY section 8 a-0 / DESIGN.md
// section 1: fp32, every operation one IEEE round-to-nearest (the build uses -ffp-contract=off, no fast-math),
// dot = (x*x + y*y) + z*z, normalize = v / length(v) component-wise, reflect = I - (2*dot(N,I))*N, min/max return the
// non-NaN operand, float -> uint truncates and saturates (NaN -> 0), unsuffixed literals are single precision (Metal has
// no double: the build uses -fsingle-precision-constant), textures hold fp32 RGBA, sampling with normalised coordinates +
// address::repeat + filter::nearest.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace metal {

// ---- uint: a 32-bit unsigned whose construction from float saturates (Metal / CUDA semantics; the C++ cast is undefined
// out of range).  All arithmetic happens on the built-in unsigned int through the conversion operator.
struct U32 {
    unsigned int v;
    U32() = default;
    constexpr U32(unsigned int x) : v(x) {}
    constexpr U32(int x) : v((unsigned int)x) {}
    U32(float f) {
        if (!(f > 0.0f)) v = 0u;                                  // negative, zero, NaN
        else if (f >= 4294967296.0f) v = 0xFFFFFFFFu;
        else v = (unsigned int)f;                                 // in range: truncation toward zero
    }
    constexpr operator unsigned int() const { return v; }
    U32 &operator++() { ++v; return *this; }
    U32 operator++(int) { U32 o = *this; ++v; return o; }
    U32 &operator--() { --v; return *this; }
    U32 operator--(int) { U32 o = *this; --v; return o; }
};
static_assert(sizeof(U32) == 4, "uint is 4 bytes");

struct float2; struct float3; struct float4; struct uint2;

// swizzles that alias the leading components of a vector (members of anonymous unions, hence trivial types)
struct swz2 { float v[2]; inline operator float2() const; };
struct swz3 { float v[3]; inline operator float3() const; inline float3 operator-() const; };

struct packed_float2 { float x, y; };
struct packed_float3 { float x, y, z; inline operator float3() const; };
struct packed_float4 { float x, y, z, w; inline operator float4() const; };
static_assert(sizeof(packed_float3) == 12 && sizeof(packed_float4) == 16 && sizeof(packed_float2) == 8, "packed layouts");

struct uint2 {
    U32 x, y;
    uint2() = default;
    uint2(U32 a, U32 b) : x(a), y(b) {}
    uint2(int a, int b) : x(a), y(b) {}
    explicit inline uint2(const float2 &f);
    explicit inline uint2(const swz2 &f);
};
static_assert(sizeof(uint2) == 8, "uint2 is 8 bytes");
inline uint2 operator+(uint2 a, uint2 b) { return uint2(U32(a.x + b.x), U32(a.y + b.y)); }

struct float2 {
    union { struct { float x, y; }; swz2 xy; };
    float2() : x(0), y(0) {}
    float2(float a, float b) : x(a), y(b) {}
    explicit float2(const uint2 &u) : x((float)(unsigned int)u.x), y((float)(unsigned int)u.y) {}
};

struct float3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; swz3 xyz; swz3 rgb; swz2 xy; };
    float3() : x(0), y(0), z(0) {}
    float3(float s) : x(s), y(s), z(s) {}
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    float3 &operator+=(const float3 &o) { x = x + o.x; y = y + o.y; z = z + o.z; return *this; }
    float3 &operator*=(const float3 &o) { x = x * o.x; y = y * o.y; z = z * o.z; return *this; }
};

struct float4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; swz3 xyz; swz3 rgb; swz2 xy; };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(const float3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float4 &operator+=(const float4 &o) { x = x + o.x; y = y + o.y; z = z + o.z; w = w + o.w; return *this; }
    float4 &operator/=(float s) { x = x / s; y = y / s; z = z / s; w = w / s; return *this; }
};
static_assert(sizeof(float4) == 16, "float4 is 16 bytes (emissions buffer)");

inline swz2::operator float2() const { return float2(v[0], v[1]); }
inline swz3::operator float3() const { return float3(v[0], v[1], v[2]); }
inline float3 swz3::operator-() const { return float3(-v[0], -v[1], -v[2]); }
inline packed_float3::operator float3() const { return float3(x, y, z); }
inline packed_float4::operator float4() const { return float4(x, y, z, w); }
inline uint2::uint2(const float2 &f) : x(f.x), y(f.y) {}
inline uint2::uint2(const swz2 &f) : x(f.v[0]), y(f.v[1]) {}

// component-wise arithmetic, one IEEE operation per component
inline float3 operator+(const float3 &a, const float3 &b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3 &a, const float3 &b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(const float3 &a, const float3 &b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(const float3 &a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, const float3 &a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(const float3 &a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator-(const float3 &a) { return float3(-a.x, -a.y, -a.z); }
inline float4 operator+(const float4 &a, const float4 &b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator/(const float4 &a, float s) { return float4(a.x / s, a.y / s, a.z / s, a.w / s); }

// library functions (canonical arithmetic, see the header comment)
inline float dot(const float3 &a, const float3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float3 cross(const float3 &a, const float3 &b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float sqrt(float x) { return ::sqrtf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float length(const float3 &a) { return ::sqrtf(dot(a, a)); }
inline float3 normalize(const float3 &a) { const float l = length(a); return float3(a.x / l, a.y / l, a.z / l); }
inline float3 reflect(const float3 &i, const float3 &n) { const float k = 2.0f * dot(n, i); return i - k * n; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }

// textures: fp32 RGBA, row-major
enum class access { read, write, read_write, sample };
enum class address { clamp_to_edge, repeat };
enum class filter { nearest, linear };
enum class mem_flags { mem_none, mem_device, mem_threadgroup };
struct sampler {
    address a; filter f;
    constexpr sampler(address a_, filter f_) : a(a_), f(f_) {}
};
template <typename T, access A>
struct texture2d {
    float *rgba;            // width * height * 4
    unsigned int width, height;
    float4 read(uint2 p) const {
        const unsigned int px = p.x, py = p.y;
        if (px >= width || py >= height) return float4();                     // out-of-bounds reads return 0
        const float *t = rgba + 4 * ((size_t)py * width + px);
        return float4(t[0], t[1], t[2], t[3]);
    }
    void write(float4 c, uint2 p) const {
        const unsigned int px = p.x, py = p.y;
        if (px >= width || py >= height) return;                              // out-of-bounds writes are dropped
        float *t = rgba + 4 * ((size_t)py * width + px);
        t[0] = c.x; t[1] = c.y; t[2] = c.z; t[3] = c.w;
    }
    // normalised coordinates, address::repeat, filter::nearest (the only sampler the shader builds)
    float4 sample(sampler, float2 uv) const {
        const float fu = uv.x - ::floorf(uv.x), fv = uv.y - ::floorf(uv.y);
        int ix = (int)::floorf(fu * (float)width), iy = (int)::floorf(fv * (float)height);
        ix = ix < 0 ? 0 : (ix > (int)width - 1 ? (int)width - 1 : ix);
        iy = iy < 0 ? 0 : (iy > (int)height - 1 ? (int)height - 1 : iy);
        const float *t = rgba + 4 * ((size_t)iy * width + (size_t)ix);
        return float4(t[0], t[1], t[2], t[3]);
    }
};

// threadgroup barrier: provided by the fiber scheduler in ref_shader.cpp
void threadgroup_barrier(mem_flags);

}  // namespace metal
