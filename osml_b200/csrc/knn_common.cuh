// Shared device helpers: dtype traits that reproduce numpy's arithmetic for predict.py:742 bit for bit, the
// (distance, index) ordering key, and numpy's two summation orders.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace osml {

constexpr int kIdxSentinel = 0x7fffffff;             // local index of an empty list slot
constexpr int64_t kIdxSentinel64 = 0x7fffffffffffffffLL;

// ---------------------------------------------------------------------------------------------------------------------
// Elem<T>: how numpy evaluates  sqrt(sum(square(t - q)))  for element type T.
//   floats: every subtract, multiply, add rounds separately in T (the _rn intrinsics are never contracted to FMA);
//   ints:   subtract and square wrap in T, the sum widens to int64 / uint64 and wraps there, sqrt is taken in fp64.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T> struct Elem;

template <> struct Elem<float> {
    using Acc = float; using Dist = float;
    static constexpr bool kIsFloat = true;
    __device__ __forceinline__ static Acc sq_diff(float t, float q) { float d = __fsub_rn(t, q); return __fmul_rn(d, d); }
    __device__ __forceinline__ static Acc add(Acc a, Acc b) { return __fadd_rn(a, b); }
    __device__ __forceinline__ static Dist finish(Acc a) { return __fsqrt_rn(a); }
    // conservative bound in accumulator space: acc > thr  =>  finish(acc) > worst
    __device__ __forceinline__ static Acc threshold(Dist worst) {
        if (!(worst < INFINITY)) return INFINITY;   // NaN or inf: let everything through to the exact comparison
        return __fmul_ru(__fmul_ru(worst, worst), 1.0000005f);
    }
    __device__ __forceinline__ static bool may_beat(Acc a, Acc thr) { return !(a > thr); }
};

template <> struct Elem<double> {
    using Acc = double; using Dist = double;
    static constexpr bool kIsFloat = true;
    __device__ __forceinline__ static Acc sq_diff(double t, double q) { double d = __dsub_rn(t, q); return __dmul_rn(d, d); }
    __device__ __forceinline__ static Acc add(Acc a, Acc b) { return __dadd_rn(a, b); }
    __device__ __forceinline__ static Dist finish(Acc a) { return __dsqrt_rn(a); }
    __device__ __forceinline__ static Acc threshold(Dist worst) {
        if (!(worst < (double)INFINITY)) return (double)INFINITY;
        return __dmul_ru(__dmul_ru(worst, worst), 1.000000000000001);
    }
    __device__ __forceinline__ static bool may_beat(Acc a, Acc thr) { return !(a > thr); }
};

// widest cheap unsigned type the wrap-around arithmetic can be carried out in (low bits are all that survive)
template <int kBytes> struct WrapWord { using type = unsigned int; };
template <> struct WrapWord<8> { using type = unsigned long long; };

template <typename T, typename U> struct ElemSigned {
    using Acc = long long; using Dist = double;
    using W = typename WrapWord<sizeof(T)>::type;
    static constexpr bool kIsFloat = false;
    __device__ __forceinline__ static Acc sq_diff(T t, T q) {
        T d = (T)(U)((W)(U)t - (W)(U)q);
        T p = (T)(U)((W)(U)d * (W)(U)d);
        return (Acc)p;
    }
    __device__ __forceinline__ static Acc add(Acc a, Acc b) { return (Acc)((unsigned long long)a + (unsigned long long)b); }
    __device__ __forceinline__ static Dist finish(Acc a) { return __dsqrt_rn(__ll2double_rn(a)); }
    __device__ __forceinline__ static Acc threshold(Dist worst) {
        if (!(worst < 3.0e9)) return 0x7fffffffffffffffLL;
        double t = __dmul_ru(__dmul_ru(worst, worst), 1.000000000000001) + 1.0;
        return (Acc)t;
    }
    // a negative sum has sqrt = NaN; it can only enter while the list still holds NaN / empty slots, and then the
    // threshold is at its maximum, so the comparison below already lets it through.
    __device__ __forceinline__ static bool may_beat(Acc a, Acc thr) { return a <= thr; }
};

template <typename T> struct ElemUnsigned {
    using Acc = unsigned long long; using Dist = double;
    using W = typename WrapWord<sizeof(T)>::type;
    static constexpr bool kIsFloat = false;
    __device__ __forceinline__ static Acc sq_diff(T t, T q) {
        T d = (T)((W)t - (W)q);
        T p = (T)((W)d * (W)d);
        return (Acc)p;
    }
    __device__ __forceinline__ static Acc add(Acc a, Acc b) { return a + b; }
    __device__ __forceinline__ static Dist finish(Acc a) { return __dsqrt_rn(__ull2double_rn(a)); }
    __device__ __forceinline__ static Acc threshold(Dist worst) {
        if (!(worst < 4.0e9)) return 0xffffffffffffffffULL;
        double t = __dmul_ru(__dmul_ru(worst, worst), 1.000000000000001) + 1.0;
        return (Acc)t;
    }
    __device__ __forceinline__ static bool may_beat(Acc a, Acc thr) { return a <= thr; }
};

template <> struct Elem<int8_t> : ElemSigned<int8_t, uint8_t> {};
template <> struct Elem<int16_t> : ElemSigned<int16_t, uint16_t> {};
template <> struct Elem<int32_t> : ElemSigned<int32_t, uint32_t> {};
template <> struct Elem<int64_t> : ElemSigned<int64_t, uint64_t> {};
template <> struct Elem<uint8_t> : ElemUnsigned<uint8_t> {};
template <> struct Elem<uint16_t> : ElemUnsigned<uint16_t> {};
template <> struct Elem<uint32_t> : ElemUnsigned<uint32_t> {};
template <> struct Elem<uint64_t> : ElemUnsigned<uint64_t> {};

// ---------------------------------------------------------------------------------------------------------------------
// ordering key: (distance, index) ascending, NaN after everything (np.argpartition sorts NaN last)
// ---------------------------------------------------------------------------------------------------------------------
template <typename D, typename I>
__device__ __forceinline__ bool key_less(D d1, I i1, D d2, I i2) {
    const bool n1 = d1 != d1, n2 = d2 != d2;
    if (n1 | n2) return (n1 & n2) ? (i1 < i2) : n2;
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

template <typename D> __device__ __forceinline__ D nan_of();
template <> __device__ __forceinline__ float nan_of<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double nan_of<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---------------------------------------------------------------------------------------------------------------------
// exact distance of one (train row, query) pair in numpy's order; scalar, used by rescoring and the generic kernel
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ typename Elem<T>::Acc sum_sequential(const T* __restrict__ t, const T* __restrict__ q, int d) {
    using E = Elem<T>;
    typename E::Acc acc = E::sq_diff(t[0], q[0]);
    for (int j = 1; j < d; ++j) acc = E::add(acc, E::sq_diff(t[j], q[j]));
    return acc;
}

// numpy pairwise_sum leaf: n <= 128 (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE = 128)
template <typename T>
__device__ __forceinline__ typename Elem<T>::Acc pairwise_leaf(const T* __restrict__ t, const T* __restrict__ q, int n) {
    using E = Elem<T>;
    using Acc = typename E::Acc;
    if (n < 8) {
        Acc res = E::sq_diff(t[0], q[0]);   // numpy starts from 0. ; 0 + x == x for x >= +0
        for (int i = 1; i < n; ++i) res = E::add(res, E::sq_diff(t[i], q[i]));
        return res;
    }
    Acc r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = E::sq_diff(t[j], q[j]);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = E::add(r[j], E::sq_diff(t[i + j], q[i + j]));
    }
    Acc res = E::add(E::add(E::add(r[0], r[1]), E::add(r[2], r[3])), E::add(E::add(r[4], r[5]), E::add(r[6], r[7])));
    for (; i < n; ++i) res = E::add(res, E::sq_diff(t[i], q[i]));
    return res;
}

// full pairwise recursion, iterative with an explicit stack (depth <= 24 covers any int32 d)
template <typename T>
__device__ inline typename Elem<T>::Acc sum_pairwise(const T* __restrict__ t, const T* __restrict__ q, int d) {
    using E = Elem<T>;
    using Acc = typename E::Acc;
    if (d <= 128) return pairwise_leaf<T>(t, q, d);
    // post-order walk: each frame is (lo, n, state); state 0 = descend left, 1 = descend right, 2 = combine
    int lo_stack[24], n_stack[24];
    unsigned char st_stack[24];
    Acc val_stack[24];
    int sp = 0, vp = 0;
    lo_stack[0] = 0; n_stack[0] = d; st_stack[0] = 0;
    while (sp >= 0) {
        const int lo = lo_stack[sp], n = n_stack[sp];
        if (n <= 128) {
            val_stack[vp++] = pairwise_leaf<T>(t + lo, q + lo, n);
            --sp;
            continue;
        }
        int n2 = n / 2; n2 -= n2 % 8;
        if (st_stack[sp] == 0) {
            st_stack[sp] = 1;
            ++sp; lo_stack[sp] = lo; n_stack[sp] = n2; st_stack[sp] = 0;
        } else if (st_stack[sp] == 1) {
            st_stack[sp] = 2;
            ++sp; lo_stack[sp] = lo + n2; n_stack[sp] = n - n2; st_stack[sp] = 0;
        } else {
            const Acc right = val_stack[--vp];
            const Acc left = val_stack[--vp];
            val_stack[vp++] = E::add(left, right);
            --sp;
        }
    }
    return val_stack[0];
}

template <typename T>
__device__ __forceinline__ typename Elem<T>::Dist exact_distance(const T* __restrict__ t, const T* __restrict__ q, int d,
                                                                 int sum_order) {
    using E = Elem<T>;
    if (!E::kIsFloat || sum_order == 1) return E::finish(sum_sequential<T>(t, q, d));   // integer adds are associative
    return E::finish(sum_pairwise<T>(t, q, d));
}

}  // namespace osml
