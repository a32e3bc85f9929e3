// arith.cuh -- the two arithmetic policies of enum c3sc_arith.
#pragma once

namespace c3sc {

// Reference order, no contraction: the CPU build is -std=c99 on x86-64 without
// -march (CMakeLists.txt:37) => separate IEEE mul/add/div, never an FMA.
struct Exact {
    static constexpr bool exact = true;
    __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
    __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
    __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
    // a*b + c as two roundings
    __device__ __forceinline__ static double mad(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
};

// Free to contract and reassociate.
struct Fast {
    static constexpr bool exact = false;
    __device__ __forceinline__ static double mul(double a, double b) { return a * b; }
    __device__ __forceinline__ static double add(double a, double b) { return a + b; }
    __device__ __forceinline__ static double sub(double a, double b) { return a - b; }
    __device__ __forceinline__ static double div(double a, double b) { return a / b; }
    __device__ __forceinline__ static double mad(double a, double b, double c) { return fma(a, b, c); }
};

}  // namespace c3sc
