// Fixed-width multi-limb unsigned integers (little-endian u32 limbs) for host and device.
// Everything the decimal layer (dec.cuh), the fixed-point exp/ln (fxpow.cuh) and the exact
// shortest-path distances (solve) need.  No reference counterpart: the reference gets this from
// CPython's libmpdec (Decimal) and from the third-party fastpathz solver's big integers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#define PB_HDN __host__ __device__
#define PB_HDNI __host__ __device__ __noinline__      /* one copy in the kernel image: keeps the hot loops inside the I-cache */
#else
#define PB_HD inline
#define PB_HDN
#define PB_HDNI __attribute__((noinline))
#endif

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;
typedef uint8_t u8;

#define PB_TABLE(type, name, dims) static const type h_##name dims
#include "tables.inc"
#undef PB_TABLE
#ifdef __CUDACC__
#define PB_TABLE(type, name, dims) static __device__ const type d_##name dims
#include "tables.inc"
#undef PB_TABLE
#endif
#ifdef __CUDA_ARCH__
#define TBL(name) d_##name
#else
#define TBL(name) h_##name
#endif

template <int N>
struct Wide {
    u32 w[N];
};

template <int N>
PB_HD void w_zero(Wide<N>& a) {
#pragma unroll
    for (int i = 0; i < N; i++) a.w[i] = 0;
}
template <int N>
PB_HD Wide<N> w_from_u64(u64 v) {
    Wide<N> a;
    w_zero(a);
    a.w[0] = (u32)v;
    if (N > 1) a.w[1] = (u32)(v >> 32);
    return a;
}
template <int N>
PB_HD bool w_is_zero(const Wide<N>& a) {
    u32 o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= a.w[i];
    return o == 0;
}
// resize (zero-extend or truncate)
template <int M, int N>
PB_HD Wide<M> w_resize(const Wide<N>& a) {
    Wide<M> r;
#pragma unroll
    for (int i = 0; i < M; i++) r.w[i] = (i < N) ? a.w[i] : 0u;
    return r;
}
template <int N>
PB_HD int w_cmp(const Wide<N>& a, const Wide<N>& b) {
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        if (a.w[i] != b.w[i]) return a.w[i] > b.w[i] ? 1 : -1;
    }
    return 0;
}
template <int N>
PB_HD u32 w_add(Wide<N>& a, const Wide<N>& b) {   // a += b, returns carry
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        c += (u64)a.w[i] + b.w[i];
        a.w[i] = (u32)c;
        c >>= 32;
    }
    return (u32)c;
}
template <int N>
PB_HD u32 w_sub(Wide<N>& a, const Wide<N>& b) {   // a -= b, returns borrow
    i64 c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        c += (i64)a.w[i] - (i64)b.w[i];
        a.w[i] = (u32)c;
        c >>= 32;
    }
    return (u32)(c & 1);
}
template <int N>
PB_HD u32 w_add_small(Wide<N>& a, u32 v) {
    u64 c = v;
#pragma unroll
    for (int i = 0; i < N; i++) {
        c += a.w[i];
        a.w[i] = (u32)c;
        c >>= 32;
    }
    return (u32)c;
}
template <int N>
PB_HD u32 w_mul_small(Wide<N>& a, u32 m) {        // a *= m, returns carry-out limb
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        c += (u64)a.w[i] * m;
        a.w[i] = (u32)c;
        c >>= 32;
    }
    return (u32)c;
}
// schoolbook product, column-wise with a 96-bit accumulator
template <int NA, int NB>
PB_HD Wide<NA + NB> w_mul(const Wide<NA>& a, const Wide<NB>& b) {
    Wide<NA + NB> r;
    u64 acc = 0;
    u32 hi = 0;
#pragma unroll
    for (int k = 0; k < NA + NB - 1; k++) {
#pragma unroll
        for (int i = 0; i < NA; i++) {
            int j = k - i;
            if (j >= 0 && j < NB) {
                u64 p = (u64)a.w[i] * b.w[j];
                acc += p;
                hi += (acc < p) ? 1u : 0u;
            }
        }
        r.w[k] = (u32)acc;
        acc = (acc >> 32) | ((u64)hi << 32);
        hi = 0;
    }
    r.w[NA + NB - 1] = (u32)acc;
    return r;
}
template <int N>
PB_HD int w_bitlen(const Wide<N>& a) {
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        if (a.w[i]) {
#ifdef __CUDA_ARCH__
            return 32 * i + (32 - __clz((int)a.w[i]));
#else
            return 32 * i + (32 - __builtin_clz(a.w[i]));
#endif
        }
    }
    return 0;
}
// variable shifts with compile-time limb indices only (keeps the limbs in registers on the GPU)
template <int N>
PB_HD Wide<N> w_shr(const Wide<N>& a, int s) {     // logical right shift, 0 <= s < 32*N
    Wide<N> r = a;
    const int ws = s >> 5, bs = s & 31;
#pragma unroll
    for (int k = 32; k >= 1; k >>= 1) {
        if (k < N && (ws & k)) {
#pragma unroll
            for (int i = 0; i < N; i++) r.w[i] = (i + k < N) ? r.w[(i + k < N) ? i + k : 0] : 0u;
        }
    }
    if (bs) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            u32 hi = (i + 1 < N) ? r.w[(i + 1 < N) ? i + 1 : 0] : 0u;
            r.w[i] = (r.w[i] >> bs) | (hi << (32 - bs));
        }
    }
    return r;
}
template <int N>
PB_HD Wide<N> w_shl(const Wide<N>& a, int s) {     // left shift, bits shifted out are lost
    Wide<N> r = a;
    const int ws = s >> 5, bs = s & 31;
#pragma unroll
    for (int k = 32; k >= 1; k >>= 1) {
        if (k < N && (ws & k)) {
#pragma unroll
            for (int i = N - 1; i >= 0; i--) r.w[i] = (i - k >= 0) ? r.w[(i - k >= 0) ? i - k : 0] : 0u;
        }
    }
    if (bs) {
#pragma unroll
        for (int i = N - 1; i >= 0; i--) {
            u32 lo = (i - 1 >= 0) ? r.w[(i - 1 >= 0) ? i - 1 : 0] : 0u;
            r.w[i] = (r.w[i] << bs) | (lo >> (32 - bs));
        }
    }
    return r;
}
// 64-by-32 division through a precomputed floor(2^64/d): returns x/d, sets rem.  Requires x < d*2^32.
PB_HD u32 div_u64_magic(u64 x, u32 d, u64 magic, u32& rem) {
#ifdef __CUDA_ARCH__
    u64 q = __umul64hi(x, magic);
#else
    u64 q = (u64)(((unsigned __int128)x * magic) >> 64);
#endif
    u64 r = x - q * d;
    while (r >= d) {
        r -= d;
        q++;
    }
    rem = (u32)r;
    return (u32)q;
}
// a /= 10^r (1 <= r <= 9), returns the remainder
template <int N>
PB_HD u32 w_div_p10(Wide<N>& a, int r) {
    const u32 d = TBL(p10_u32)[r];
    const u64 magic = TBL(p10_magic)[r];
    u32 rem = 0;
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        u64 x = ((u64)rem << 32) | a.w[i];
        a.w[i] = div_u64_magic(x, d, magic, rem);
    }
    return rem;
}
template <int N>
PB_HD Wide<N> w_pow10(int k) {                      // 10^k truncated to N limbs (k < PB_NPOW10)
    Wide<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.w[i] = (i < 8) ? TBL(pow10_w8)[k][(i < 8) ? i : 0] : 0u;
    return r;
}
// number of decimal digits (0 for zero); valid for values < 10^77
template <int N>
PB_HD int w_ndigits(const Wide<N>& a) {
    int bl = w_bitlen(a);
    if (bl == 0) return 0;
    int d = (bl * 1233) >> 12;                      // floor(bl*log10(2)) : d or d+1 digits
    Wide<N> p = w_pow10<N>(d);
    return (w_cmp(a, p) >= 0) ? d + 1 : d;
}
// a *= 10^k (k < PB_NPOW10); the caller guarantees the product fits
template <int N>
PB_HD void w_mul_pow10(Wide<N>& a, int k) {
    while (k >= 9) {
        w_mul_small(a, 1000000000u);
        k -= 9;
    }
    if (k > 0) w_mul_small(a, TBL(p10_u32)[k]);
}
// Strip up to `maxz` trailing decimal zeros from a; returns how many were removed.  Chunked (8, 4, 2,
// 1 digits) so that exact quotients like 1/4 = 0.25000... cost a handful of short divisions.
template <int N>
PB_HD int w_strip_zeros(Wide<N>& a, int maxz) {
    int done = 0;
#pragma unroll 1
    for (int step = 8; step >= 1; step >>= 1) {
        while (maxz - done >= step) {
            Wide<N> t = a;
            u32 rem = w_div_p10(t, step);
            if (rem != 0) break;
            a = t;
            done += step;
        }
    }
    return done;
}
// Drop the k lowest decimal digits of a with round-half-even.  *inexact is set when a non-zero
// digit was dropped.  If `sticky_in` is non-zero the dropped part is treated as being followed by
// further non-zero digits (used for division remainders).
template <int N>
PB_HD void w_round_drop(Wide<N>& a, int k, bool sticky_in, bool* inexact) {
    bool sticky = sticky_in;
    if (k <= 0) {
        if (inexact) *inexact = sticky;
        return;
    }
    while (k > 9) {
        u32 rem = w_div_p10(a, 9);
        sticky = sticky || (rem != 0);
        k -= 9;
    }
    u32 rem = w_div_p10(a, k);
    u32 half = 5u * TBL(p10_u32)[k - 1];
    bool up;
    if (rem > half) up = true;
    else if (rem < half) up = false;
    else up = sticky || (a.w[0] & 1u);
    if (inexact) *inexact = sticky || (rem != 0);
    if (up) w_add_small(a, 1u);
}
// Knuth algorithm D: q = floor(u / v), r = u mod v.  v != 0.  NU >= NV.
template <int NU, int NV>
PB_HDNI void w_divmod(const Wide<NU>& u_in, const Wide<NV>& v_in, Wide<NU>& q, Wide<NV>& r) {
    int n = NV;
    while (n > 0 && v_in.w[n - 1] == 0) n--;
    w_zero(q);
    w_zero(r);
    if (n == 1) {
        u32 d = v_in.w[0];
        u64 rem = 0;
        for (int i = NU - 1; i >= 0; i--) {
            u64 x = (rem << 32) | u_in.w[i];
            q.w[i] = (u32)(x / d);
            rem = x % d;
        }
        r.w[0] = (u32)rem;
        return;
    }
#ifdef __CUDA_ARCH__
    int s = __clz((int)v_in.w[n - 1]);
#else
    int s = __builtin_clz(v_in.w[n - 1]);
#endif
    u32 v[NV];
    u32 u[NU + 1];
    for (int i = n - 1; i > 0; i--) v[i] = s ? ((v_in.w[i] << s) | (v_in.w[i - 1] >> (32 - s))) : v_in.w[i];
    v[0] = v_in.w[0] << s;
    u[NU] = s ? (u_in.w[NU - 1] >> (32 - s)) : 0u;
    for (int i = NU - 1; i > 0; i--) u[i] = s ? ((u_in.w[i] << s) | (u_in.w[i - 1] >> (32 - s))) : u_in.w[i];
    u[0] = u_in.w[0] << s;
    for (int j = NU - n; j >= 0; j--) {
        u64 num = ((u64)u[j + n] << 32) | u[j + n - 1];
        u64 qhat = num / v[n - 1];
        u64 rhat = num % v[n - 1];
        while (qhat >= (1ull << 32) || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
            qhat--;
            rhat += v[n - 1];
            if (rhat >= (1ull << 32)) break;
        }
        i64 borrow = 0;
        u64 carry = 0;
        for (int i = 0; i < n; i++) {
            u64 p = qhat * v[i] + carry;
            carry = p >> 32;
            i64 t = (i64)u[i + j] - borrow - (i64)(p & 0xFFFFFFFFull);
            u[i + j] = (u32)t;
            borrow = (t < 0) ? 1 : 0;
        }
        i64 t = (i64)u[j + n] - borrow - (i64)carry;
        u[j + n] = (u32)t;
        if (t < 0) {
            qhat--;
            u64 c = 0;
            for (int i = 0; i < n; i++) {
                c += (u64)u[i + j] + v[i];
                u[i + j] = (u32)c;
                c >>= 32;
            }
            u[j + n] += (u32)c;
        }
        if (j < NU) q.w[j] = (u32)qhat;
    }
    for (int i = 0; i < n; i++) {
        u32 lo = u[i] >> s;
        u32 hi = (s && i + 1 <= NU) ? (u[i + 1] << (32 - s)) : 0u;
        r.w[i] = s ? (lo | hi) : u[i];
    }
}
