// f32x2.cuh -- a pair of floats that the sm_100a FP32 pipe processes with ONE instruction
// (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  Measured on B200
// (profiles/r1_microbench.txt): packed and scalar FMA have the same 128 lanes/clk/SM peak, so the
// gain is in ISSUE SLOTS -- one instruction carries two slices' worth of arithmetic.
// The pow2 "pk" kernels keep the same component of two consecutive time slices in one f2.
// Host build (CPU emulation used by the tests): plain two-float struct.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define VHD __host__ __device__ __forceinline__
#else
#define VHD inline
#endif

namespace veils {
namespace p2 {

#if defined(__CUDA_ARCH__)
struct f2 {
    unsigned long long v;
};
VHD f2 mk2(float a, float b) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
VHD float lo(f2 x) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
    (void)b;
    return a;
}
VHD float hi(f2 x) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
    (void)a;
    return b;
}
VHD f2 operator+(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
VHD f2 operator-(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
VHD f2 operator*(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
VHD f2 vfma(f2 a, f2 b, f2 c) {      // a * b + c
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
#else
struct f2 {
    float a, b;
};
VHD f2 mk2(float a, float b) { f2 r; r.a = a; r.b = b; return r; }
VHD float lo(f2 x) { return x.a; }
VHD float hi(f2 x) { return x.b; }
VHD f2 operator+(f2 x, f2 y) { return mk2(x.a + y.a, x.b + y.b); }
VHD f2 operator-(f2 x, f2 y) { return mk2(x.a - y.a, x.b - y.b); }
VHD f2 operator*(f2 x, f2 y) { return mk2(x.a * y.a, x.b * y.b); }
VHD f2 vfma(f2 x, f2 y, f2 z) { return mk2(std::fmaf(x.a, y.a, z.a), std::fmaf(x.b, y.b, z.b)); }
#endif

VHD f2 bc(float c) { return mk2(c, c); }              // broadcast (SASS: R.F32 operand form)
VHD f2 operator-(f2 a) { return bc(0.0f) - a; }

// scalar overloads so the same butterfly templates serve float, double and f2
VHD float vfma(float a, float b, float c) { return a * b + c; }
VHD double vfma(double a, double b, double c) { return a * b + c; }
template <typename V> VHD V vconst(double c);
template <> VHD float vconst<float>(double c) { return (float)c; }
template <> VHD double vconst<double>(double c) { return c; }
template <> VHD f2 vconst<f2>(double c) { return bc((float)c); }

}  // namespace p2
}  // namespace veils
