// arith.cuh -- the two arithmetic profiles the ray/shade code is written against.
//
// The reference has two observable arithmetics:
//
//  * ArithDev  -- what nvcc makes of the reference's kernel.cu with its own flags
//                 (--use_fast_math, reference GNUmakefile:8): flush-to-zero, FMA
//                 contraction, div.approx / sqrt.approx / lg2.approx / ex2.approx.
//                 Every primitive here is ONE explicit PTX instruction, so the
//                 result does not depend on how this translation unit is compiled.
//                 Running the same PTX ops in the same order as the reference's
//                 PTX is what makes LYAP_MODE_EXACT bit-match the reference kernel
//                 on the same GPU.
//  * ArithHost -- the reference's host build (IEEE single ops, no contraction,
//                 correctly rounded divide/sqrt); used by LYAP_MODE_HOST, which is
//                 checked against the CPU oracle.
//
// `madd(a,b,c)` is "a*b+c as the respective build evaluates it": one fused op on
// the device profile, two rounded ops on the host profile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lyap {

struct ArithDev {
    static __device__ __forceinline__ float mul(float a, float b)
    {
        float r;
        asm("mul.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    }
    static __device__ __forceinline__ float add(float a, float b)
    {
        float r;
        asm("add.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    }
    static __device__ __forceinline__ float sub(float a, float b)
    {
        float r;
        asm("sub.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    }
    static __device__ __forceinline__ float fma(float a, float b, float c)
    {
        float r;
        asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
        return r;
    }
    static __device__ __forceinline__ float madd(float a, float b, float c) { return fma(a, b, c); }
    static __device__ __forceinline__ float div(float a, float b)
    {
        float r;
        asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    }
    static __device__ __forceinline__ float sqrt(float a)
    {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float lg2(float a)
    {
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float ex2(float a)
    {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    // natural log as fast-math lowers std::log: lg2.approx * ln2, one rounded multiply
    static __device__ __forceinline__ float log(float a) { return mul(lg2(a), 0.693147182f); }
    // __powf: ex2(y * lg2(x))
    static __device__ __forceinline__ float pow(float x, float y) { return ex2(mul(y, lg2(x))); }
    static __device__ __forceinline__ double f2d(float a)
    {
        double r;
        asm("cvt.ftz.f64.f32 %0, %1;" : "=d"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float d2f(double a)
    {
        float r;
        asm("cvt.rn.ftz.f32.f64 %0, %1;" : "=f"(r) : "d"(a));
        return r;
    }
    // a.b as the fast-math build contracts x*x' + y*y' + z*z': fma(z, fma(x, mul(y)))
    static __device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
    {
        return fma(az, bz, fma(ax, bx, mul(ay, by)));
    }
    // Color::to_rgba on the device: cvt.rzi.u32.f64 saturates, then the low byte is stored
    static __device__ __forceinline__ uint8_t to_byte(float c)
    {
        unsigned u;
        double d = f2d(c) * 255.0;
        asm("cvt.rzi.u32.f64 %0, %1;" : "=r"(u) : "d"(d));
        return (uint8_t)u;
    }
};

struct ArithHost {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float madd(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ double f2d(float a) { return (double)a; }
    static __device__ __forceinline__ float d2f(double a) { return __double2float_rn(a); }
    static __device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
    {
        return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
    }
    // x86 (unsigned char)(255.0*c): truncating 32-bit conversion, low byte kept;
    // out-of-range and NaN give the "integer indefinite" 0x80000000 -> byte 0.
    static __device__ __forceinline__ uint8_t to_byte(float c)
    {
        double d = (double)c * 255.0;
        int i = (d > -2147483649.0 && d < 2147483648.0) ? (int)d : (int)0x80000000;
        return (uint8_t)i;
    }
};

__device__ __forceinline__ bool is_finite(float a) { return fabsf(a) < __int_as_float(0x7f800000); }
__device__ __forceinline__ float quiet_nan() { return __int_as_float(0x7fc00000); }

} // namespace lyap
