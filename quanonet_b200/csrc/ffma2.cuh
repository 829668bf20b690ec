// Blackwell packed-FP32 helpers: one 64-bit register pair holds a complex amplitude (lo = re, hi = im).
//
// fma2<PAT>(a, z, c) = (a, a) * P(z) + c  and  mul2<PAT>(a, z) = (a, a) * P(z)  as ONE FFMA2 / FMUL2 each:
// sm_100's FFMA2 takes a 32-bit register broadcast to both halves (.F32), a half swap (.LO_HI) and
// per-half negation (-, .NP) as operand modifiers.  PTX's fma.rn.f32x2 has no such modifiers; ptxas
// folds an adjacent unpack / neg / repack into them, but only when the repacked value has a single
// use.  Written in plain CUDA C++ the compiler's CSE shares the repacked operand between the two
// outputs of a gate and ptxas then materialises it (MOV + FADD per use — measured: 1,241 MOV + 786
// FADD in the fwd+grad kernel).  Keeping the unpack/neg/repack INSIDE each asm block gives every
// FFMA2 a private copy that always folds (verified with cuobjdump: no MOV / FADD remains).
//
// Operand patterns P(z) for z = (x, y):
//   0: ( x,  y)    1: ( y,  x)    2: (-y,  x) = i z     3: ( y, -x) = -i z
//   4: (-x,  y)    5: ( x, -y) = conj z    6: (-x, -y)    7: (-y, -x)
#pragma once
#include <cuda_runtime.h>

namespace qon {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float lo2(u64 v) { float a, b; unpack2(v, a, b); return a; }
__device__ __forceinline__ float hi2(u64 v) { float a, b; unpack2(v, a, b); return b; }

#define QON_P2_DECL ".reg .b64 pa, pb; .reg .f32 zx, zy, tx, ty; mov.b64 pa, {%1, %1}; mov.b64 {zx, zy}, %2; "
#define QON_P2_0 "mov.b64 pb, {zx, zy}; "
#define QON_P2_1 "mov.b64 pb, {zy, zx}; "
#define QON_P2_2 "neg.f32 ty, zy; mov.b64 pb, {ty, zx}; "
#define QON_P2_3 "neg.f32 tx, zx; mov.b64 pb, {zy, tx}; "
#define QON_P2_4 "neg.f32 tx, zx; mov.b64 pb, {tx, zy}; "
#define QON_P2_5 "neg.f32 ty, zy; mov.b64 pb, {zx, ty}; "
#define QON_P2_6 "neg.f32 tx, zx; neg.f32 ty, zy; mov.b64 pb, {tx, ty}; "
#define QON_P2_7 "neg.f32 tx, zx; neg.f32 ty, zy; mov.b64 pb, {ty, tx}; "

#define QON_FMA2_CASE(N)                                                                               \
    if constexpr (PAT == N)                                                                            \
        asm("{ " QON_P2_DECL QON_P2_##N "fma.rn.f32x2 %0, pb, pa, %3; }" : "=l"(d) : "f"(a), "l"(z), "l"(c));
#define QON_MUL2_CASE(N)                                                                               \
    if constexpr (PAT == N)                                                                            \
        asm("{ " QON_P2_DECL QON_P2_##N "mul.rn.f32x2 %0, pb, pa; }" : "=l"(d) : "f"(a), "l"(z));

template <int PAT>
__device__ __forceinline__ u64 fma2(float a, u64 z, u64 c) {
    u64 d;
    QON_FMA2_CASE(0) QON_FMA2_CASE(1) QON_FMA2_CASE(2) QON_FMA2_CASE(3)
    QON_FMA2_CASE(4) QON_FMA2_CASE(5) QON_FMA2_CASE(6) QON_FMA2_CASE(7)
    return d;
}

template <int PAT>
__device__ __forceinline__ u64 mul2(float a, u64 z) {
    u64 d;
    QON_MUL2_CASE(0) QON_MUL2_CASE(1) QON_MUL2_CASE(2) QON_MUL2_CASE(3)
    QON_MUL2_CASE(4) QON_MUL2_CASE(5) QON_MUL2_CASE(6) QON_MUL2_CASE(7)
    return d;
}

// elementwise a * b + c on two packed operands (no broadcast)
__device__ __forceinline__ u64 fma2_vv(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// packed a + b (one FADD2)
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

}  // namespace qon
