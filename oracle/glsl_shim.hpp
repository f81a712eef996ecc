// TEST INFRASTRUCTURE — a small GLSL-in-C++ shim of our own, so that pure functions of the reference's compute shaders can
// be compiled BY g++ FROM THE SOURCES WHERE THEY LIE (oracle/ref_extract.py) and run on the CPU as the pin of the oracle.
//
// What is OURS here (and therefore not "reference-run"): the vector types and the builtins below.  GLSL leaves the precision
// of several builtins to the implementation; this shim fixes them as follows (the policy the parity tests state):
//   * + - * / sqrt           IEEE fp32, one rounding per operation (the library is built with -ffp-contract=off)
//   * literals               `1.0` is a FLOAT in GLSL: ref_extract.py appends `f` to every floating literal it extracts
//   * min(x,y) / max(x,y)    y < x ? y : x  /  x < y ? y : x            (GLSL 4.60 8.3)
//   * sign(x)                1, 0 or -1
//   * dot(a,b)               (a.x*b.x + a.y*b.y) + a.z*b.z              (left to right)
//   * length(v), normalize   sqrt(dot(v,v)),  v / length(v)
//   * mat4 * vec4            (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)   — the order of glm, the only CPU-side matrix
//                            code the reference links (camera.cpp)
//   * tan, log, pow, exp2    libm tanf / logf / powf / exp2f of the build container
//   * inverse(mat4)          generic cofactor expansion, 1/det multiplied in (what a driver's generic lowering does).
//                            Tests may inject a matrix instead (glsl::g_inverse_override) to compare everything downstream
//                            of this one implementation-defined builtin bit for bit.
//   * uint(float)            C++ conversion (undefined for negative / NaN in both languages: the fixtures avoid them)
#pragma once
#include <cmath>
#include <cstdint>

namespace glsl {

typedef uint32_t uint;
constexpr uint UINT32_MAX_ = 0xFFFFFFFFu;
constexpr float INF = 1e35f;     // common.glsl:5
constexpr float EPS = 1e-5f;     // bounce_point_lights.comp:8

struct vec2; struct vec3; struct vec4; struct uvec2; struct uvec3; struct ivec2;

// swizzles: members of a union with the components, convertible to / assignable from the vector they name
template <typename T, int N, typename V, int A, int B>
struct swz2
{
    T d[N];
    operator V() const { return V(d[A], d[B]); }
    swz2& operator=(const V& v) { const T a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
};
template <typename T, int N, typename V, int A, int B, int C>
struct swz3
{
    T d[N];
    operator V() const { return V(d[A], d[B], d[C]); }
    swz3& operator=(const V& v) { const T a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }
};

struct vec2
{
    union { struct { float x, y; }; float d[2]; };
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
    vec2(const uvec2& u);
    float& operator[](uint i) { return d[i]; }
    float operator[](uint i) const { return d[i]; }
};
struct uvec2
{
    union { struct { uint x, y; }; uint d[2]; };
    uvec2() : x(0), y(0) {}
    explicit uvec2(uint s) : x(s), y(s) {}
    explicit uvec2(int s) : x((uint) s), y((uint) s) {}
    explicit uvec2(float s) : x((uint) s), y((uint) s) {}
    uvec2(uint a, uint b) : x(a), y(b) {}
    explicit uvec2(const vec2& v) : x((uint) v.x), y((uint) v.y) {}
};
inline vec2::vec2(const uvec2& u) : x((float) u.x), y((float) u.y) {}
struct ivec2
{
    union { struct { int x, y; }; int d[2]; swz2<int, 2, ivec2, 0, 1> xy; };
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const uvec2& u) : x((int) u.x), y((int) u.y) {}
};
struct vec3
{
    union { struct { float x, y, z; }; struct { float r, g, b; }; float d[3]; swz2<float, 3, vec2, 0, 1> xy; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    explicit vec3(int s) : x((float) s), y((float) s), z((float) s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](uint i) { return d[i]; }
    float operator[](uint i) const { return d[i]; }
};
struct uvec3
{
    union { struct { uint x, y, z; }; uint d[3]; swz2<uint, 3, uvec2, 0, 1> xy; swz2<uint, 3, uvec2, 1, 2> yz; };
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    explicit uvec3(const vec3& v) : x((uint) v.x), y((uint) v.y), z((uint) v.z) {}
};
struct vec4
{
    union
    {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        float d[4];
        swz2<float, 4, vec2, 0, 1> xy;
        swz3<float, 4, vec3, 0, 1, 2> xyz;
        swz3<float, 4, vec3, 0, 1, 2> rgb;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float e) : x(a), y(b), z(c), w(e) {}
    vec4(const vec2& v, float c, float e) : x(v.x), y(v.y), z(c), w(e) {}
    vec4(const vec3& v, float e) : x(v.x), y(v.y), z(v.z), w(e) {}
    float& operator[](uint i) { return d[i]; }
    float operator[](uint i) const { return d[i]; }
};

// ---- component-wise arithmetic (one IEEE fp32 rounding per operation) -----------------------------------------------------------
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator+(const vec2& a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(const vec2& a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator/(const vec2& a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator/(const uvec2& a, const vec2& b) { return vec2((float) a.x / b.x, (float) a.y / b.y); }   // uvec2 -> vec2 implicitly
inline uvec2 operator+(const uvec2& a, const uvec2& b) { return uvec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator*(const ivec2& a, int s) { return ivec2(a.x * s, a.y * s); }

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, const vec3& b) { a = a * b; return a; }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec4& operator/=(vec4& a, float s) { a = a / s; return a; }

// ---- builtins ----------------------------------------------------------------------------------------------------------------
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float floor(float x) { return std::floor(x); }
inline vec2 floor(const vec2& v) { return vec2(std::floor(v.x), std::floor(v.y)); }
inline vec3 floor(const vec3& v) { return vec3(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float tan(float x) { return ::tanf(x); }
inline float log(float x) { return ::logf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float exp2(float x) { return ::exp2f(x); }
inline float acos(float x) { return ::acosf(x); }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v / length(v); }
inline int findLSB(uint v) { return v == 0 ? -1 : __builtin_ctz(v); }
inline uint bitCount(uint v) { return (uint) __builtin_popcount(v); }

struct mat4
{
    vec4 c[4];   // columns
    vec4& operator[](uint i) { return c[i]; }
    const vec4& operator[](uint i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    return (m.c[0] * v.x + m.c[1] * v.y) + (m.c[2] * v.z + m.c[3] * v.w);
}

// tests can pin inverse() (implementation-defined in GLSL) to a given matrix to compare the rest of a function exactly
inline const mat4*& inverse_override() { static const mat4* p = nullptr; return p; }

inline mat4 inverse(const mat4& m)
{
    if (inverse_override()) return *inverse_override();
    // cofactor expansion: inv = adj(m) / det(m), a[i][j] = m.c[j][i] (row i, column j)
    float a[4][4], cof[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) a[i][j] = m.c[j][i];
    auto minor3 = [&](int r, int c) {
        int rr[3], cc[3];
        for (int i = 0, k = 0; i < 4; i++) if (i != r) rr[k++] = i;
        for (int j = 0, k = 0; j < 4; j++) if (j != c) cc[k++] = j;
        const float t0 = a[rr[0]][cc[0]] * (a[rr[1]][cc[1]] * a[rr[2]][cc[2]] - a[rr[1]][cc[2]] * a[rr[2]][cc[1]]);
        const float t1 = a[rr[0]][cc[1]] * (a[rr[1]][cc[0]] * a[rr[2]][cc[2]] - a[rr[1]][cc[2]] * a[rr[2]][cc[0]]);
        const float t2 = a[rr[0]][cc[2]] * (a[rr[1]][cc[0]] * a[rr[2]][cc[1]] - a[rr[1]][cc[1]] * a[rr[2]][cc[0]]);
        return (t0 - t1) + t2;
    };
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) cof[i][j] = (((i + j) & 1) ? -1.0f : 1.0f) * minor3(i, j);
    const float det = ((a[0][0] * cof[0][0] + a[0][1] * cof[0][1]) + a[0][2] * cof[0][2]) + a[0][3] * cof[0][3];
    const float inv_det = 1.0f / det;
    mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.c[j][i] = cof[j][i] * inv_det;   // inverse(row i, col j) = cof(j, i) / det
    return r;
}

// ---- the little bit of the shader environment the extracted bodies touch -----------------------------------------------------
template <typename T>
struct buffer_array   // `T name[];` inside a buffer block
{
    T* data = nullptr;
    uint count = 0;
    T& operator[](uint i) { return data[i]; }
    const T& operator[](uint i) const { return data[i]; }
    uint length() const { return count; }
};
struct image2D   // r32f storage image
{
    float* data = nullptr;
    int width = 0, height = 0;
};
inline ivec2 imageSize(const image2D& im) { return ivec2(im.width, im.height); }
inline vec4 imageLoad(const image2D& im, const ivec2& p) { return vec4(im.data[(size_t) p.y * im.width + p.x], 0.0f, 0.0f, 1.0f); }
inline void imageStore(image2D& im, const ivec2& p, const vec4& v) { im.data[(size_t) p.y * im.width + p.x] = v.x; }

}   // namespace glsl
