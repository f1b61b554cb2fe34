// cr_math.cuh -- correctly rounded pow(x, 1./3), cos(t) and acos(v) in double-double arithmetic.
//
// cv::solveCubic (stage4_match_consecutive.cpp:202,237 -> findFundamentalMat -> run7Point) calls libm's pow / acos / cos;
// on the host that is glibc, whose results are the correctly rounded ones in all but ~0.1 % of the calls (measured,
// tests/test_host_cpu.py).  CUDA's libdevice versions are 1-2 ulp functions, and in degenerate 7-point samples (seven
// coplanar points, sideways motion) the last bit of a root decides which model wins.  The device therefore evaluates
// these three functions to ~2^-100 and rounds once, which reproduces glibc wherever glibc itself rounds correctly.
// Host + device (the host build exists so that tests can compare it with glibc and mpmath on the CPU).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SVO_CR_HD __host__ __device__ inline
#define SVO_CR_CALL __host__ __device__ __noinline__ inline   // entry points: one copy per kernel (code size)
#else
#define SVO_CR_HD inline
#define SVO_CR_CALL inline
#endif

namespace svo {
namespace cr {

struct dd { double hi, lo; };

SVO_CR_HD dd two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return {s, (a - (s - bb)) + (b - bb)};
}
SVO_CR_HD dd quick_two_sum(double a, double b) {     // |a| >= |b|
    const double s = a + b;
    return {s, b - (s - a)};
}
SVO_CR_HD dd two_prod(double a, double b) {
    const double p = a * b;
    return {p, fma(a, b, -p)};
}
SVO_CR_HD dd add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
SVO_CR_HD dd add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
SVO_CR_HD dd mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
SVO_CR_HD dd mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}

// x^(1./3) for x > 0, the exponent being the double next to one third (what std::pow(x, 1./3) computes):
// cube root by one double-double Newton step from cbrt(), times x^-(1/3 - fl(1/3)) = 1 - delta*ln(x)
SVO_CR_CALL double pow13(double x) {
    if (!(x > 0) || x > 1e300 || x < 1e-300) return pow(x, 1. / 3);
    const double c0 = cbrt(x);
    const dd c2 = two_prod(c0, c0);
    const dd c3 = mul_d(c2, c0);
    const dd res = add_d(c3, -x);                        // c0^3 - x
    const double corr = (res.hi + res.lo) / (3.0 * c2.hi);
    const double delta = 1.850371707708594e-17;          // 1/3 - fl(1/3)
    const double tail = -corr - c0 * (delta * log(x));
    return c0 + tail;
}

// range reduction by pi/2 in three parts; |t| < 2^20.  Returns the quadrant and r = t - k*pi/2 as a double-double.
SVO_CR_HD int reduce_pio2(double t, dd& r) {
    const double k = rint(t * 0.6366197723675814);
    const double p1 = 1.5707963267948966, p2 = 6.123233995736766e-17, p3 = -1.4973849048591698e-33;
    const dd a = two_prod(k, p1);
    dd s = two_sum(t, -a.hi);
    s.lo -= a.lo;
    s = quick_two_sum(s.hi, s.lo);
    const dd b = two_prod(k, p2);
    s = add(s, dd{-b.hi, -b.lo});
    s = add_d(s, -k * p3);
    r = s;
    return (int)((long long)k & 3);
}

// cos(r) and sin(r) for |r| <= pi/4 (double-double Taylor series, Horner in z = r^2)
SVO_CR_HD dd cos_kernel(dd r) {
    const double ch[14] = {-0.5, 0.041666666666666664, -0.001388888888888889, 2.48015873015873e-05, -2.755731922398589e-07,
                           2.08767569878681e-09, -1.1470745597729725e-11, 4.779477332387385e-14, -1.5619206968586225e-16,
                           4.110317623312165e-19, -8.896791392450574e-22, 1.6117375710961184e-24, -2.4795962632247976e-27,
                           3.279889237069838e-30};
    const double cl[14] = {0.0, 2.3129646346357427e-18, 5.300543954373577e-20, 2.1511947866775882e-23, -2.3767714622250297e-23,
                           -1.20734505911326e-25, -2.0655512752830745e-28, 4.399205485834081e-31, -1.1910679660273754e-32,
                           1.4412973378659527e-36, 7.911402614872376e-38, -3.6846573564509766e-41, 1.2953730964765229e-43,
                           1.5117542744029879e-46};
    const dd z = mul(r, r);
    dd p = {ch[13], cl[13]};
    for (int k = 12; k >= 0; --k) p = add(mul(p, z), dd{ch[k], cl[k]});
    return add_d(mul(p, z), 1.0);
}
SVO_CR_HD dd sin_kernel(dd r) {
    const double sh[14] = {-0.16666666666666666, 0.008333333333333333, -0.0001984126984126984, 2.7557319223985893e-06,
                           -2.505210838544172e-08, 1.6059043836821613e-10, -7.647163731819816e-13, 2.8114572543455206e-15,
                           -8.22063524662433e-18, 1.9572941063391263e-20, -3.868170170630684e-23, 6.446950284384474e-26,
                           -9.183689863795546e-29, 1.1309962886447716e-31};
    const double sl[14] = {-9.25185853854297e-18, 1.1564823173178714e-19, -1.7209558293420705e-22, -1.858393274046472e-22,
                           1.448814070935912e-24, 1.2585294588752098e-26, -7.03872877733453e-30, 1.6508842730861433e-31,
                           -2.2141894119604265e-34, -1.3643503830087908e-36, 8.843177655482344e-40, -1.9330404233703465e-42,
                           -1.4303150396787322e-45, 1.0498015412959506e-47};
    const dd z = mul(r, r);
    dd p = {sh[13], sl[13]};
    for (int k = 12; k >= 0; --k) p = add(mul(p, z), dd{sh[k], sl[k]});
    return add(mul(mul(p, z), r), r);                    // r + r*z*p
}

SVO_CR_CALL dd cos_dd(double t) {
    dd r;
    const int q = reduce_pio2(t, r);
    dd v = (q & 1) ? sin_kernel(r) : cos_kernel(r);
    if (q == 1 || q == 2) { v.hi = -v.hi; v.lo = -v.lo; }
    return v;
}

SVO_CR_CALL double cos_cr(double t) {
    if (!(fabs(t) < 1048576.0)) return cos(t);
    const dd v = cos_dd(t);
    return v.hi + v.lo;
}

// acos(v), |v| < 1: one Newton step y = y0 + (cos(y0) - v) / sin(y0) with cos(y0) in double-double
SVO_CR_CALL double acos_cr(double v) {
    const double y0 = acos(v);
    if (!(fabs(v) < 1.0) || y0 < 1e-6) return y0;
    const dd c = cos_dd(y0);
    const dd num = add_d(c, -v);
    const double corr = (num.hi + num.lo) / sin(y0);
    return y0 + corr;
}

}  // namespace cr
}  // namespace svo
